"""Host emulation of the device code (tests/emu): the SAME per-thread bodies the CUDA kernels run
(csrc/vlb_bvh.cuh, vlb_shade.cuh) executed serially on the CPU and compared with the oracle. This
is how the LBVH build / traversal / shading logic is checked where no GPU exists; the GPU parity
tests (-m gpu) are the real gate."""
import numpy as np
import pytest

import emu_api
from conftest import rel_l2


def _rays(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    return o, (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("max_leaf", [1, 4, 8])
def test_lbvh_hit_ids_bit_exact_room(oa, scenes, max_leaf):
    sc = scenes.small_room()
    e, o = emu_api.Scene(sc, max_leaf), oa.Scene(sc)
    org, d = _rays(20000, 0.1, 3.9, 1)
    ib, tb = o.trace_rays(org, d, accel=1)
    ie, te, _ = e.trace_rays(org, d)
    assert np.array_equal(ib, ie) and np.array_equal(tb, te)
    a = o.trace_rays(org, d, tmin=0.0, tmax=1.2, accel=1, kind=1)[0] >= 0
    b = e.trace_rays(org, d, tmin=0.0, tmax=1.2, kind=1)[0] >= 0
    assert np.array_equal(a, b)


def test_lbvh_hit_ids_bit_exact_atrium_sample(oa, scenes):
    sc = scenes.atrium(32768, seed=7)
    e, o = emu_api.Scene(sc), oa.Scene(sc)
    assert o.n_triangles == 32768 and e.max_depth < 64
    org, d = _rays(3000, 0.5, 11.5, 2)
    org *= np.array([2.5, 1.0, 1.5], np.float32)
    ib, tb = o.trace_rays(org, d, accel=1)
    ie, te, cnt = e.trace_rays(org, d)
    assert np.array_equal(ib, ie) and np.array_equal(tb, te)
    assert cnt[0] / len(org) < 200


def test_degenerate_and_tiny_scenes(oa, scenes):
    sc = scenes.default_cube()
    one = dict(sc)
    one["indices"] = sc["indices"][:3].copy()
    one["instances"] = sc["instances"].copy()
    one["instances"]["index_count"] = 3
    for scene in (one, sc):
        e, o = emu_api.Scene(scene), oa.Scene(scene)
        org, d = _rays(5000, -0.9, 0.9, 3)
        assert np.array_equal(o.trace_rays(org, d, accel=1)[0], e.trace_rays(org, d)[0])
    # coincident duplicate triangles: ties broken by the smaller flat id everywhere
    dup = dict(sc)
    dup["instances"] = np.concatenate([sc["instances"], sc["instances"]])
    e, o = emu_api.Scene(dup), oa.Scene(dup)
    org, d = _rays(5000, -0.9, 0.9, 4)
    ib = o.trace_rays(org, d, accel=1)[0]
    assert np.array_equal(ib, e.trace_rays(org, d)[0]) and ib.max() < 12


@pytest.mark.parametrize("order", [2, 3])
def test_bake_parity_emu_vs_oracle(oa, vlb, scenes, order):
    sc = scenes.small_room()
    e, o = emu_api.Scene(sc), oa.Scene(sc)
    sky = scenes.hdr_sky(64, 32, seed=4)
    e.set_skybox(sky)
    o.set_skybox(sky)
    s = vlb.default_settings()
    s.probes[:] = (2, 2, 3)
    s.dir_w, s.dir_h = 24, 12
    s.sh_order = order
    s.light_pos[:] = (2.0, 3.5, 2.0)
    vlb.settings_from_bounds(s, o.bounds(True))
    for flags in (vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE,
                  vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE | vlb.QUANTIZE_RGBA8,
                  vlb.SKYBOX_ON_MISS, vlb.SHADOW_RAYS | vlb.SH_WORLD_FRAME):
        s.flags = flags
        assert rel_l2(e.bake(s), o.bake_probes(s)[0]) <= 1e-3

"""Host emulation of the device code (tests/emu): the SAME per-thread bodies the CUDA kernels run
(csrc/vlb_bvh.cuh, vlb_shade.cuh) executed serially on the CPU and compared with the oracle. This
is how the LBVH build / traversal / shading logic is checked where no GPU exists; the GPU parity
tests (-m gpu) are the real gate."""
import os

import numpy as np
import pytest

import emu_api
from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


def test_q8_node_format_is_conservative(scenes, tmp_path):
    """VLB_NODE_Q8=1 (64-byte nodes, 8-bit planes, vlb_bvh.cuh store_node4 / bvh4_step): the quantised boxes must
    contain the fp32 ones, so closest hits (id, t, u, v) are bit-identical to the fp32-node build, at the price
    of a few more node visits. Runs the device code's host twin in a child process per format."""
    import subprocess
    import sys
    code = r'''
import sys, importlib, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import emu_api
q8, so, out = sys.argv[1], sys.argv[2], sys.argv[3]
so = emu_api.build(so=so, defines=["-DVLB_NODE_Q8=" + q8.split()[0]] + q8.split()[1:])
emu_api.build = lambda force=False, so=so, defines=(): so
scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
s = emu_api.Scene(scenes.atrium(32768, seed=7))
rng = np.random.default_rng(5)
n = 30000
o = (rng.uniform(0.03, 0.97, (n, 3)) * np.array(scenes.HALL)).astype(np.float32)
d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
d[:300, 0] = 0.0; d[300:600, 1] = 0.0; d[600:900, 2] = 0.0            # axis-parallel slabs (idir = 1e30)
ids, tuv, cnt = s.trace_rays(o, d)
ids2, _, _ = s.trace_rays(o, d, tmin=0.0, tmax=3.0, kind=1)
np.savez(out, ids=ids, tuv=tuv, cnt=cnt, any_ids=ids2)
''' % (ROOT, os.path.join(ROOT, "tests", "emu"))
    res = []
    for q8 in ("0", "1", "1 -DVLB_NODE_ORDER1D=1", "0 -DVLB_BVH8=1"):
        tag = q8.replace(" ", "").replace("-", "").replace("=", "")
        out = str(tmp_path / ("r%s.npz" % tag))
        subprocess.check_call([sys.executable, "-c", code, q8, str(tmp_path / ("emu_q8_%s.so" % tag)), out], timeout=600)
        res.append(np.load(out))
    a, b, c, w8 = res
    # VLB_BVH8 (8-wide nodes, 8-bit planes, octant-ordered slots, (node, mask) groups on the stack): same hits, fewer steps
    assert np.array_equal(a["ids"], w8["ids"]) and np.array_equal(a["tuv"], w8["tuv"])
    assert np.array_equal(a["any_ids"] >= 0, w8["any_ids"] >= 0)
    assert w8["cnt"][0] <= 0.85 * a["cnt"][0]
    # VLB_NODE_ORDER1D (children visited in slot order along the node's ordering axis, no distance sort): traversal order
    # never changes a result, only the number of nodes visited (measured +9 % on primary rays)
    assert np.array_equal(a["ids"], c["ids"]) and np.array_equal(a["tuv"], c["tuv"])
    assert np.array_equal(a["any_ids"] >= 0, c["any_ids"] >= 0)
    assert c["cnt"][0] <= 1.25 * a["cnt"][0]
    assert (a["ids"] >= 0).mean() > 0.5
    assert np.array_equal(a["ids"], b["ids"]) and np.array_equal(a["tuv"], b["tuv"])
    assert np.array_equal(a["any_ids"] >= 0, b["any_ids"] >= 0)          # any-hit: same occlusion answer
    assert b["cnt"][0] >= a["cnt"][0]                                     # never fewer nodes: a superset is visited
    assert b["cnt"][0] <= 1.05 * a["cnt"][0]                              # and only a few more (measured +1.6 %)

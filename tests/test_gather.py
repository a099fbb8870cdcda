"""Multi-bounce gather (BASELINE configs[3]; include/vlb_bake.h: vlb_bake_gather_device) on the CPU:
the oracle's restatement of shaders/main.rchit:124-163 + shaders/sh.rmiss:20-36 checked against a
closed form and its own invariants, the device code run on the host (tests/emu) against the oracle,
and the sharded multi-pass driver (parallel.bake_multibounce_sharded) over gloo, world_size 2."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

import emu_api
from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _settings(vlb, o, probes=(3, 2, 3), dirs=(24, 12), order=3):
    s = vlb.default_settings()
    s.probes[:] = probes
    s.dir_w, s.dir_h = dirs
    s.sh_order = order
    s.light_pos[:] = (2.0, 3.5, 2.0)
    vlb.settings_from_bounds(s, o.bounds(True))
    s.flags = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS
    return s


def test_zero_source_is_the_direct_pass_bitwise(oa, vlb, scenes):
    o = oa.Scene(scenes.small_room())
    o.set_skybox(scenes.hdr_sky(64, 32, seed=4))
    s = _settings(vlb, o)
    direct, _ = o.bake_probes(s)
    assert np.array_equal(o.bake_gather(s, None)[0], direct)
    assert np.array_equal(o.bake_gather(s, np.zeros_like(direct))[0], direct)


def test_gather_is_linear_in_the_source_without_srgb(oa, vlb, scenes):
    o = oa.Scene(scenes.small_room())
    s = _settings(vlb, o, order=2)
    s.flags = vlb.SHADOW_RAYS                       # linear radiance: no sRGB, no quantisation
    rng = np.random.default_rng(5)
    X = rng.normal(size=(s.n_probes, 16, 3)).astype(np.float32)
    X[:, 9:] = 0
    base = o.bake_gather(s, np.zeros_like(X))[0].astype(np.float64)
    a = o.bake_gather(s, X)[0] - base
    b = o.bake_gather(s, (2.5 * X).astype(np.float32))[0] - base
    assert np.abs(b - 2.5 * a).max() <= 2e-5 * np.abs(a).max()
    s.indirect_gain = 0.25
    c = o.bake_gather(s, X)[0] - base
    assert np.abs(c - 0.25 * a).max() <= 2e-5 * np.abs(a).max()


def test_constant_source_over_a_floor_closed_form(oa, vlb, scenes):
    """One floor quad under the grid, every probe holds the same L0-only SH (constant radiance c): all 8 corner
    probes are visible from any floor point, the weights cancel, so the hit radiance is exactly
    baseColor * gain * 0.282095 * c00 and the pass equals that constant projected over the directions
    that hit the floor."""
    b = scenes._Builder()
    pos = np.array([[-2, 0, -2], [2, 0, -2], [2, 0, 2], [-2, 0, 2]], np.float32)   # under the grid: weights > 0
    nrm = np.tile(np.array([[0, 1, 0]], np.float32), (4, 1))
    m = b.add_mesh(pos, nrm, np.array([0, 2, 1, 0, 3, 2], np.uint32))
    b.add_instance(m, scenes.identity12(), 2)
    sc = b.finish(scenes.make_materials())
    o = oa.Scene(sc)
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 3)
    s.dir_w, s.dir_h = 32, 16
    s.sh_order = 3
    vlb.settings_from_bounds(s, (-2.0, 0.5, -2.0, 2.0, 1.5, 2.0))
    s.flags = 0                                      # no shadow rays, no sky, linear
    s.c_diffuse = s.c_specular = 0.0                 # direct term off: only the gathered term remains
    s.indirect_gain = 1.7
    c00 = np.array([0.9, 0.5, 0.2], np.float32)
    prev = np.zeros((s.n_probes, 16, 3), np.float32)
    prev[:, 0] = c00
    got, _ = o.bake_gather(s, prev)
    t, r, w = oa.probe_dirs(s.dir_w, s.dir_h)
    bc = np.asarray(sc["materials"]["base_color_factor"][2][:3], np.float64)
    rad = bc * 1.7 * (0.282095 * c00.astype(np.float64))
    pp = oa.probe_positions(s)
    for q in (0, 7, s.n_probes - 1):
        rr = r.reshape(-1, 3)
        ids, _ = o.trace_rays(np.tile(pp[q], (len(rr), 1)), rr, accel=1)
        hit = (ids >= 0).reshape(-1)
        basis = oa.sh_basis(t.reshape(-1, 3))[:, :16].astype(np.float64)
        exp = (basis[hit] * w.reshape(-1)[hit, None].astype(np.float64)).sum(0)[:, None] * rad[None, :]
        assert hit.any() and not hit.all()
        assert np.linalg.norm(got[q] - exp) <= 1e-5 * np.linalg.norm(exp)


@pytest.mark.parametrize("order", [2, 3])
def test_gather_parity_emu_vs_oracle(oa, vlb, scenes, order):
    sc = scenes.small_room()
    e, o = emu_api.Scene(sc), oa.Scene(sc)
    sky = scenes.hdr_sky(64, 32, seed=4)
    e.set_skybox(sky)
    o.set_skybox(sky)
    s = _settings(vlb, o, order=order)
    for flags in (vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS, vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE,
                  vlb.SHADOW_RAYS | vlb.SH_WORLD_FRAME):
        s.flags = flags
        prev, _ = o.bake_probes(s)
        ref, _ = o.bake_gather(s, prev)
        assert rel_l2(e.bake(s, prev), ref) <= 1e-3
        assert rel_l2(ref, prev) > 1e-2              # the gathered term is really there


def test_multibounce_converges(oa, vlb, scenes):
    o = oa.Scene(scenes.small_room())
    s = _settings(vlb, o, order=2)
    s.flags = vlb.SHADOW_RAYS
    passes, prev = [], None
    for _ in range(4):
        prev, _ = o.bake_gather(s, prev)
        passes.append(prev.astype(np.float64))
    d = [np.abs(passes[i + 1] - passes[i]).max() for i in range(3)]
    assert d[0] > d[1] > d[2] > 0                    # albedo < 1: every bounce adds less
    s.bounces = 3
    assert np.array_equal(o.bake_multibounce(s), passes[3].astype(np.float32))


# ---------------------------------------------------------------- sharded driver over gloo ----
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cyclic, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    vlb = importlib.import_module("vulkan-light-bakery_b200")
    par = importlib.import_module("vulkan-light-bakery_b200.parallel")
    scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
    from oracle import oracle_api as oa
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oa.set_num_threads(2)
        osc = oa.Scene(scenes.small_room())
        s = vlb.default_settings()
        s.probes[:] = (2, 2, 3)
        s.dir_w, s.dir_h = 16, 8
        s.sh_order = 2
        s.flags = vlb.SHADOW_RAYS
        s.light_pos[:] = (2.0, 3.5, 2.0)
        s.bounces = 2
        vlb.settings_from_bounds(s, osc.bounds(tight=True))

        def bake_pass(ss, prev, out):
            got, _ = osc.bake_gather(ss, None if prev is None else prev.numpy())
            out.copy_(torch.from_numpy(got.reshape(-1, 48)))

        full = par.bake_multibounce_sharded(bake_pass, s, rank, world, device="cpu", cyclic=cyclic)
        if rank == 0:
            ref = osc.bake_multibounce(s).reshape(-1, 48)
            q.put(("ok", bool(np.array_equal(full.numpy(), ref)), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cyclic", [False, True])
def test_multibounce_sharded_world2_equals_single(cyclic):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cyclic, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    tag, equal, shape = q.get(timeout=5)
    assert tag == "ok" and equal and shape == (12, 48)

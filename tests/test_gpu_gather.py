"""GPU parity of the multi-bounce gather (BASELINE configs[3]; vlb_bake_gather_device, vlb_bake_probes
with bounces > 0) against the CPU oracle's restatement of shaders/main.rchit:124-163 +
shaders/sh.rmiss:20-36. Tolerance: BASELINE.json's 1e-3 max relative L2 per probe SH vector."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

PROBE_TOL = 1e-3


@pytest.fixture(scope="module")
def room(scenes, oa):
    sc = scenes.small_room()
    return sc, oa.Scene(sc)


def _settings(vlb, ctx, flags, order=3, probes=(3, 2, 3), dirs=(32, 16)):
    s = vlb.default_settings()
    s.probes[:] = probes
    s.dir_w, s.dir_h = dirs
    s.sh_order = order
    s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = flags
    vlb.settings_from_bounds(s, ctx.scene_bounds(tight=True))
    return s


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("flags", ["linear", "srgb", "world", "noshadow"])
def test_multibounce_room_vs_oracle(ctx, vlb, scenes, room, order, flags):
    sc, osc = room
    ctx.set_scene(sc)
    sky = scenes.hdr_sky(64, 32, seed=4)
    ctx.set_skybox(sky)
    osc.set_skybox(sky)
    base = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS
    f = {"linear": base, "srgb": base | vlb.SRGB_ENCODE, "world": base | vlb.SH_WORLD_FRAME,
         "noshadow": vlb.SKYBOX_ON_MISS}[flags]
    s = _settings(vlb, ctx, f, order)
    s.bounces = 2
    s.indirect_gain = 0.8
    got = ctx.bake_probes(s)
    ref = osc.bake_multibounce(s)
    assert rel_l2(got, ref) <= PROBE_TOL
    direct = s.copy()
    direct.bounces = 0
    assert rel_l2(ctx.bake_probes(direct), ref) > 1e-2      # the bounces really changed the result
    st = ctx.last_bake_stats()
    assert st.n_primary_rays == direct.n_probes * s.dir_w * s.dir_h


def test_gather_pass_device_slabs_are_bitwise_rows_of_the_full_pass(ctx, vlb, scenes, room):
    import torch
    sc, osc = room
    ctx.set_scene(sc)
    sky = scenes.hdr_sky(64, 32, seed=4)
    ctx.set_skybox(sky)
    osc.set_skybox(sky)          # (the oracle scene is shared by the module: do not rely on an earlier test having set it)
    s = _settings(vlb, ctx, vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS, order=2, probes=(3, 2, 4))
    prev = torch.from_numpy(ctx.bake_probes(s).reshape(-1, 48)).cuda()
    full = torch.zeros((s.n_probes, 48), device="cuda")
    ctx.bake_gather_device(s, prev.data_ptr(), full.data_ptr())
    ctx.synchronize()
    ref, _ = osc.bake_gather(s, prev.cpu().numpy())
    assert rel_l2(full.cpu().numpy(), ref) <= PROBE_TOL
    # zero source and NULL source are the direct pass, bit for bit
    z = torch.zeros_like(full)
    ctx.bake_gather_device(s, torch.zeros_like(prev).data_ptr(), z.data_ptr())
    ctx.synchronize()
    assert torch.equal(z, prev)
    # cyclic shares (rank r of 3) of the gather pass == rows of the full pass
    f4 = full.view(4, 6, 48)
    for r in range(3):
        p = s.copy()
        p.slab_k0, p.slab_k1, p.slab_stride = r, 4, 3
        part = torch.zeros((p.n_slab_probes, 48), device="cuda")
        ctx.bake_gather_device(p, prev.data_ptr(), part.data_ptr())
        ctx.synchronize()
        assert torch.equal(part.view(-1, 6, 48), f4[r::3])


def test_gather_atrium_sample_vs_oracle(vlb, oa, scenes):
    import torch
    sc = scenes.atrium(16384, seed=7)
    sky = scenes.hdr_sky(128, 64, seed=1)
    with vlb.Context(0) as c:
        c.set_scene(sc)
        c.build_bvh()
        c.set_skybox(sky)
        s = scenes.atrium_settings(probes=(6, 4, 5), dirs=(32, 32), order=3, bounds=(0, 0, 0) + tuple(scenes.HALL))
        s.flags = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS
        s.indirect_gain = 1.0
        prev = torch.from_numpy(c.bake_probes(s).reshape(-1, 48)).cuda()
        out = torch.zeros((s.n_probes, 48), device="cuda")
        c.bake_gather_device(s, prev.data_ptr(), out.data_ptr())
        c.synchronize()
        got = out.cpu().numpy()
    osc = oa.Scene(sc)
    osc.set_skybox(sky)
    ids = np.linspace(0, s.n_probes - 1, 24).astype(np.int64)
    ref, _ = osc.bake_gather(s, prev.cpu().numpy(), probe_ids=ids)
    assert rel_l2(got[ids], ref) <= PROBE_TOL
    assert rel_l2(got[ids], prev.cpu().numpy()[ids]) > 1e-3


def test_gather_errors(ctx, vlb, scenes, room):
    import torch
    sc, _ = room
    ctx.set_scene(sc)
    ctx.set_skybox(scenes.hdr_sky(64, 32, seed=4))
    s = _settings(vlb, ctx, vlb.SHADOW_RAYS)
    s.bounces = 1
    s.slab_k0, s.slab_k1 = 0, 1
    with pytest.raises(vlb.VlbError) as e:                   # a slab cannot iterate on its own
        ctx.bake_probes(s)
    assert e.value.code == vlb.ERR_INVALID
    s.slab_k0, s.slab_k1 = 0, -1
    s.flags |= vlb.REFERENCE_PROBE_ORDER
    buf = torch.zeros((s.n_probes, 48), device="cuda")
    with pytest.raises(vlb.VlbError) as e:
        ctx.bake_gather_device(s, buf.data_ptr(), buf.data_ptr())
    assert e.value.code == vlb.ERR_INVALID

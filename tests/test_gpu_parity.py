"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.
Tolerances are BASELINE.json's: max relative L2 <= 1e-4 per skybox SH vector, <= 1e-3 per probe SH
vector; BVH hit ids bit-exact against the brute-force CUDA intersector (and the oracle)."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

SKY_TOL = 1e-4
PROBE_TOL = 1e-3


def _rays(n, lo, hi, seed=0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


# ---------------------------------------------------------------- skybox projection ----------
@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("shape", [(32, 64), (256, 512), (100, 314), (7, 13), (1, 1), (33, 130), (130, 66), (1000, 3141)])
def test_skybox_rgba32f_vs_oracle(ctx, oa, scenes, order, shape):
    img = scenes.hdr_sky(shape[1], shape[0], seed=5)
    got = ctx.skybox_project_sh(img, order)
    ref = oa.skybox_project(img, order)
    assert rel_l2(got, ref) <= SKY_TOL
    if order == 2:
        assert np.all(got[9:] == 0)


@pytest.mark.parametrize("order", [2, 3])
def test_skybox_rgba8_vs_oracle(ctx, oa, order):
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (96, 200, 4), dtype=np.uint8)
    assert rel_l2(ctx.skybox_project_sh(img, order), oa.skybox_project(img, order)) <= SKY_TOL


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("shape", [(50, 97), (64, 1), (37, 258)])
def test_skybox_rgba8_unaligned_rows_fallback_kernel(ctx, oa, order, shape):
    # W*4 is not a multiple of 16: the TMA bulk-copy path cannot be used, the LDG kernel runs
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (shape[0], shape[1], 4), dtype=np.uint8)
    assert rel_l2(ctx.skybox_project_sh(img, order), oa.skybox_project(img, order)) <= SKY_TOL


def test_skybox_many_maps_more_tiles_than_sms(ctx, oa, scenes):
    # 40 maps x 8 strips = 320 whole-column tiles > 148 persistent CTAs: several tiles per CTA
    maps = [scenes.hdr_sky(512, 64, seed=200 + i) for i in range(40)]
    got = ctx.skybox_project_sh_batched(maps, 2)
    for i in (0, 7, 39):
        assert rel_l2(got[i], oa.skybox_project(maps[i], 2)) <= SKY_TOL


@pytest.mark.parametrize("order", [2, 3])
def test_envmap_vs_oracle(ctx, oa, scenes, order):
    img = scenes.hdr_sky(314, 100, seed=2)        # sh.comp on a (scaled-down) 3141x1000-shaped map
    assert rel_l2(ctx.envmap_project_sh(img, order), oa.envmap_project(img, order)) <= SKY_TOL


def test_skybox_constant_image_kat(ctx):
    # analytic: constant radiance c projects to c00 = 2*sqrt(pi)*c, everything else ~ 0
    img = np.ones((1024, 2048, 4), np.float32)
    img[..., 1] = 0.5
    got = ctx.skybox_project_sh(img, 3)
    assert np.allclose(got[0], 2 * np.sqrt(np.pi) * np.array([1, 0.5, 1]), rtol=2e-5)
    assert np.abs(got[1:]).max() < 2e-5


def test_skybox_batched_equals_single(ctx, scenes):
    maps = [scenes.hdr_sky(128, 64, seed=100 + i) for i in range(9)]
    got = ctx.skybox_project_sh_batched(maps, 3)
    for i, m in enumerate(maps):
        assert rel_l2(got[i], ctx.skybox_project_sh(m, 3)) <= 1e-6


def test_skybox_full_size_c1(ctx, oa, scenes):
    img = scenes.hdr_sky(2048, 1024, seed=1)      # BASELINE config 1
    assert rel_l2(ctx.skybox_project_sh(img, 2), oa.skybox_project(img, 2)) <= SKY_TOL


def test_skybox_deterministic(ctx, scenes):
    img = scenes.hdr_sky(512, 256, seed=3)
    a = ctx.skybox_project_sh(img, 3)
    for _ in range(3):
        assert np.array_equal(a, ctx.skybox_project_sh(img, 3))


@pytest.mark.parametrize("order,fmt", [(2, "f32"), (3, "f32"), (3, "u8")])
def test_skybox_device_ptrs_lanes_vs_oracle(ctx, vlb, oa, scenes, order, fmt):
    # 11 independent device-resident maps (more than the auxiliary lanes) through the multi-pointer
    # entry point: every map must match the oracle and the single-launch path bit for bit
    import torch
    n, W, H = 11, 320, 96
    if fmt == "f32":
        host = [scenes.hdr_sky(W, H, seed=300 + i) for i in range(n)]
        code = vlb.FMT_RGBA32F
    else:
        rng = np.random.default_rng(7)
        host = [rng.integers(0, 256, (H, W, 4), dtype=np.uint8) for _ in range(n)]
        code = vlb.FMT_RGBA8
    dev = [torch.from_numpy(m).cuda() for m in host]
    out = torch.full((n, 48), -1.0, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.skybox_project_sh_device_ptrs([d.data_ptr() for d in dev], code, W, H, order, out.data_ptr())
    ctx.synchronize()
    got = out.cpu().numpy().reshape(n, 16, 3)
    for i in range(n):
        assert rel_l2(got[i], oa.skybox_project(host[i], order)) <= SKY_TOL
        assert np.array_equal(got[i], ctx.skybox_project_sh(host[i], order))
    if order == 2:
        assert np.all(got[:, 9:] == 0)


def test_skybox_device_ptrs_repeated_call_is_replayed_from_a_graph(vlb, oa, scenes):
    """The same vlb_skybox_project_sh_device_ptrs call issued again and again (direct launches, then the capture, then
    graph replays) gives the same bits every time; new texels in the same buffers are picked up by the replays; a
    different argument list in between does not disturb the cached graph."""
    import torch
    W, H = 320, 96
    with vlb.Context(0) as c:
        host = [scenes.hdr_sky(W, H, seed=900 + i) for i in range(6)]
        dev = [torch.from_numpy(m).cuda() for m in host]
        out = torch.zeros((6, 48), device="cuda")
        other = torch.zeros((2, 48), device="cuda")
        ptrs = [d.data_ptr() for d in dev]
        first = None
        for rep in range(5):
            out.fill_(-1.0); torch.cuda.synchronize()
            c.skybox_project_sh_device_ptrs(ptrs, vlb.FMT_RGBA32F, W, H, 3, out.data_ptr())
            if rep == 2:
                c.skybox_project_sh_device_ptrs(ptrs[:2], vlb.FMT_RGBA32F, W, H, 2, other.data_ptr())
            c.synchronize()
            got = out.cpu().numpy().copy()
            if first is None:
                first = got
                for i in range(6):
                    assert rel_l2(got[i], oa.skybox_project(host[i], 3)) <= SKY_TOL
            assert np.array_equal(got, first), rep
        new = scenes.hdr_sky(W, H, seed=990)
        dev[3].copy_(torch.from_numpy(new)); torch.cuda.synchronize()
        c.skybox_project_sh_device_ptrs(ptrs, vlb.FMT_RGBA32F, W, H, 3, out.data_ptr()); c.synchronize()
        got = out.cpu().numpy()
        assert rel_l2(got[3], oa.skybox_project(new, 3)) <= SKY_TOL and np.array_equal(got[2], first[2])


def test_skybox_device_ptrs_errors(ctx, vlb):
    import torch
    out = torch.zeros((2, 48), device="cuda")
    with pytest.raises(vlb.VlbError):
        ctx.skybox_project_sh_device_ptrs([0, 0], vlb.FMT_RGBA32F, 64, 32, 2, out.data_ptr())
    with pytest.raises(vlb.VlbError):
        ctx.skybox_project_sh_device_ptrs([out.data_ptr()], vlb.FMT_RGBA32F, 64, 32, 5, out.data_ptr())


# ---------------------------------------------------------------- BVH hit ids ----------------
@pytest.fixture(scope="module")
def room(ctx, oa, scenes):
    sc = scenes.small_room()
    ctx.set_scene(sc)
    ctx.build_bvh()
    return sc, oa.Scene(sc)


def test_hit_ids_room_bvh_vs_brute_vs_oracle(ctx, vlb, room):
    sc, osc = room
    ctx.set_scene(sc)
    o, d = _rays(50000, 0.1, 3.9, seed=1)
    ib, tb = ctx.trace_rays(o, d, accel=vlb.TRACE_BRUTE_FORCE)
    iv, tv = ctx.trace_rays(o, d, accel=vlb.TRACE_BVH)
    io, to = osc.trace_rays(o, d, accel=1)
    assert np.array_equal(ib, iv) and np.array_equal(tb, tv)
    assert np.array_equal(ib, io) and np.array_equal(tb, to)
    assert 0.02 < (ib < 0).mean() < 0.9


def _soup(vlb, scenes, n_tris, seed, duplicates=False):
    """n_tris random small triangles in the unit cube as one instance (ragged sizes for the LBVH build: the
    hand-written radix sort works on tiles of 2,048 pairs, the collapse on levels)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.05, 0.95, (n_tris, 1, 3))
    if duplicates:                                   # many identical centroids -> identical Morton keys (ties by index)
        c[: n_tris // 2] = c[0]
    p = (c + rng.normal(size=(n_tris, 3, 3)) * 0.03).reshape(-1, 3).astype(np.float32)
    v = np.zeros(3 * n_tris, vlb.VERTEX_DTYPE)
    v["position"][:, :3] = p
    v["position"][:, 3] = 1.0
    v["normal"][:, 1] = 1.0
    inst = np.zeros(1, vlb.INSTANCE_DTYPE)
    inst["index_count"], inst["vertex_count"] = 3 * n_tris, 3 * n_tris
    inst["transform"] = scenes.identity12()
    return {"vertices": v, "indices": np.arange(3 * n_tris, dtype=np.uint32), "instances": inst, "materials": scenes.make_materials()}


@pytest.mark.parametrize("n_tris,dup", [(1, False), (2, False), (33, False), (2047, False), (2048, False), (2049, False),
                                         (4096, True), (50000, False), (300000, True)])
def test_lbvh_any_size_hit_ids_equal_brute_force(ctx, vlb, scenes, n_tris, dup):
    ctx.set_scene(_soup(vlb, scenes, n_tris, seed=n_tris, duplicates=dup))
    st = ctx.build_bvh()
    assert st.n_triangles == n_tris and st.n_nodes >= 1
    o, d = _rays(4000 if n_tris > 10000 else 20000, 0.0, 1.0, seed=3)
    ib, tb = ctx.trace_rays(o, d, accel=vlb.TRACE_BRUTE_FORCE)
    iv, tv = ctx.trace_rays(o, d, accel=vlb.TRACE_BVH)
    assert np.array_equal(ib, iv) and np.array_equal(tb, tv)
    if n_tris >= 33:
        assert (ib >= 0).any()
    st2 = ctx.build_bvh()                              # deterministic: same tree again
    assert st2.n_nodes == st.n_nodes
    iv2, tv2 = ctx.trace_rays(o, d, accel=vlb.TRACE_BVH)
    assert np.array_equal(iv, iv2) and np.array_equal(tv, tv2)


def test_any_hit_room(ctx, vlb, room):
    sc, osc = room
    ctx.set_scene(sc)
    o, d = _rays(20000, 0.1, 3.9, seed=2)
    a, _ = ctx.trace_rays(o, d, tmin=0.0, tmax=1.5, accel=vlb.TRACE_BVH, kind=vlb.TRACE_ANY)
    b, _ = ctx.trace_rays(o, d, tmin=0.0, tmax=1.5, accel=vlb.TRACE_BRUTE_FORCE, kind=vlb.TRACE_ANY)
    c, _ = osc.trace_rays(o, d, tmin=0.0, tmax=1.5, accel=1, kind=1)
    assert np.array_equal(a >= 0, b >= 0) and np.array_equal(a >= 0, c >= 0)


# ---------------------------------------------------------------- bake -----------------------
def _room_settings(vlb, ctx, flags, order=3, probes=(3, 2, 3), dirs=(32, 16)):
    s = vlb.default_settings()
    s.probes[:] = probes
    s.dir_w, s.dir_h = dirs
    s.sh_order = order
    s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = flags
    vlb.settings_from_bounds(s, ctx.scene_bounds(tight=True))
    return s


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("flags", ["base", "quant", "noshadow", "linear", "nosky", "world"])
def test_bake_room_vs_oracle(ctx, vlb, scenes, room, order, flags):
    sc, osc = room
    ctx.set_scene(sc)
    sky = scenes.hdr_sky(64, 32, seed=4)
    ctx.set_skybox(sky)
    osc.set_skybox(sky)
    base = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    f = {"base": base, "quant": base | vlb.QUANTIZE_RGBA8, "noshadow": base & ~vlb.SHADOW_RAYS,
         "linear": base & ~vlb.SRGB_ENCODE, "nosky": base & ~vlb.SKYBOX_ON_MISS,
         "world": base | vlb.SH_WORLD_FRAME}[flags]
    s = _room_settings(vlb, ctx, f, order)
    got = ctx.bake_probes(s)
    ref, n_shadow = osc.bake_probes(s)
    assert rel_l2(got, ref) <= PROBE_TOL
    st = ctx.last_bake_stats()
    assert st.n_primary_rays == s.n_probes * s.dir_w * s.dir_h
    assert st.n_shadow_rays == n_shadow
    if order == 2:
        assert np.all(got[:, 9:] == 0)


@pytest.mark.parametrize("dirs", [(5, 3), (9, 7), (33, 17), (100, 37), (64, 64)])
def test_bake_odd_direction_grids_vs_oracle(ctx, vlb, scenes, room, dirs):
    """Direction grids that do not fill their 32-texel tiles, one tile column (W <= 8: the kernel's tile / tiles_x takes its
    tiles_x == 1 branch), odd tile counts (the host reciprocal of tile_xy), and a grid of full-size chunks (two-phase schedule)."""
    sc, osc = room
    ctx.set_scene(sc)
    sky = scenes.hdr_sky(64, 32, seed=4)
    ctx.set_skybox(sky)
    osc.set_skybox(sky)
    s = _room_settings(vlb, ctx, vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE, 3, probes=(2, 1, 2), dirs=dirs)
    got = ctx.bake_probes(s)
    ref, n_shadow = osc.bake_probes(s)
    assert rel_l2(got, ref) <= PROBE_TOL
    assert ctx.last_bake_stats().n_shadow_rays == n_shadow


def test_bake_slab_union_is_bitwise_full(ctx, vlb, scenes, room):
    sc, _ = room
    ctx.set_scene(sc)
    ctx.set_skybox(scenes.hdr_sky(64, 32, seed=4))
    s = _room_settings(vlb, ctx, vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE, probes=(3, 2, 4))
    full = ctx.bake_probes(s)
    parts = []
    for k0, k1 in ((0, 1), (1, 3), (3, 4)):
        p = s.copy()
        p.slab_k0, p.slab_k1 = k0, k1
        parts.append(ctx.bake_probes(p))
    assert np.array_equal(full, np.concatenate(parts))


def test_bake_cyclic_slices_union_is_bitwise_full(ctx, vlb, scenes, room):
    # cyclic sharding (rank r of N bakes k = r, r + N, ...): interleaving the shares reproduces the full bake
    sc, _ = room
    ctx.set_scene(sc)
    ctx.set_skybox(scenes.hdr_sky(64, 32, seed=4))
    s = _room_settings(vlb, ctx, vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE, probes=(3, 2, 5))
    full = ctx.bake_probes(s).reshape(5, 6, 16, 3)
    world = 3
    for r in range(world):
        p = s.copy()
        p.slab_k0, p.slab_k1, p.slab_stride = r, 5, world
        assert p.slab_slices == list(range(r, 5, world))
        part = ctx.bake_probes(p).reshape(len(p.slab_slices), 6, 16, 3)
        assert np.array_equal(part, full[r::world])


def test_bake_device_is_asynchronous_and_stats_collect(ctx, vlb, scenes, room):
    import torch
    sc, _ = room
    ctx.set_scene(sc)
    s = _room_settings(vlb, ctx, vlb.SHADOW_RAYS | vlb.SRGB_ENCODE, probes=(3, 2, 3))
    ref = ctx.bake_probes(s)
    out = torch.zeros((s.n_probes, 48), device="cuda")
    ctx.bake_probes_device(s, out.data_ptr())
    ctx.bake_probes_device(s, out.data_ptr())          # a second enqueue first collects the pending statistics
    st = ctx.last_bake_stats()
    assert st.n_probes == s.n_probes and st.n_primary_rays == s.n_probes * s.dir_w * s.dir_h
    assert 0 < st.n_shadow_rays <= st.n_primary_rays and st.kernel_ms > 0
    assert np.array_equal(out.cpu().numpy().reshape(ref.shape), ref)


def test_bake_reference_probe_order_and_accumulate(ctx, vlb, oa, scenes, room):
    sc, osc = room
    ctx.set_scene(sc)
    sky = scenes.hdr_sky(64, 32, seed=4)
    ctx.set_skybox(sky)
    osc.set_skybox(sky)
    base = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    s = _room_settings(vlb, ctx, base, probes=(3, 2, 3))
    x_fast = ctx.bake_probes(s)
    r = s.copy()
    r.flags = base | vlb.REFERENCE_PROBE_ORDER
    ref_order = ctx.bake_probes(r)
    # same probes, permuted: match positions
    px, pr = vlb.probe_positions(s), vlb.probe_positions(r)
    for i in range(len(pr)):
        j = int(np.where((px == pr[i]).all(axis=1))[0][0])
        assert np.array_equal(ref_order[i], x_fast[j])
    a = r.copy()
    a.flags = r.flags | vlb.ACCUMULATE_ACROSS_PROBES
    acc = ctx.bake_probes(a)
    assert rel_l2(acc, osc.bake_probes(a)[0]) <= PROBE_TOL


def test_skybox_set_async_equals_blocking_upload(ctx, vlb, scenes, room):
    """vlb_skybox_set_async: the upload overlaps scene upload + LBVH build; the bake waits for it on the device."""
    sc, _ = room
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 3); s.dir_w, s.dir_h = 32, 16; s.sh_order = 2; s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    vlb.settings_from_bounds(s, (0.3, 0.3, 0.3, 3.7, 3.7, 3.7))
    sky_a, sky_b = scenes.hdr_sky(128, 64, seed=4), scenes.hdr_sky(64, 32, seed=5)
    sky_u8 = (np.clip(sky_b, 0, 1) * 255).astype(np.uint8)
    want = {}
    for name, sky in (("a", sky_a), ("b", sky_b), ("u8", sky_u8)):
        ctx.set_scene(sc); ctx.build_bvh(); ctx.set_skybox(sky)
        want[name] = ctx.bake_probes(s)
    assert not np.array_equal(want["a"], want["b"])
    for name, sky in (("a", sky_a), ("u8", sky_u8), ("b", sky_b), ("a", sky_a)):
        ctx.set_skybox_async(sky)            # first, then the calls it overlaps
        ctx.set_scene(sc)
        ctx.build_bvh()
        assert np.array_equal(ctx.bake_probes(s), want[name]), name
    ctx.set_skybox_async(sky_b)
    ctx.set_skybox(sky_a)                    # a blocking upload right behind an asynchronous one wins
    assert np.array_equal(ctx.bake_probes(s), want["a"])
    ctx.set_skybox_async(sky_b)
    ctx.synchronize()
    assert np.array_equal(ctx.bake_probes(s), want["b"])


def test_bake_errors(vlb, scenes):
    with vlb.Context(0) as c:
        s = vlb.default_settings()
        with pytest.raises(vlb.VlbError) as e:
            c.bake_probes(s)
        assert e.value.code == vlb.ERR_STATE
        c.set_scene(scenes.default_cube())
        with pytest.raises(vlb.VlbError) as e:      # SKYBOX_ON_MISS without a skybox
            c.bake_probes(s)
        assert e.value.code == vlb.ERR_STATE
        s.sh_order = 5
        with pytest.raises(vlb.VlbError) as e:
            c.bake_probes(s)
        assert e.value.code == vlb.ERR_INVALID


def test_bake_empty_scene_sees_only_sky(ctx, vlb, oa, scenes):
    sc = scenes.default_cube()
    empty = {k: v[:0] if k != "materials" else v for k, v in sc.items()}
    with vlb.Context(0) as c:
        c.set_scene(empty)
        sky = scenes.hdr_sky(64, 32, seed=9)
        c.set_skybox(sky)
        s = vlb.default_settings()
        s.probes[:] = (1, 1, 1)
        s.dir_w, s.dir_h = 64, 32
        s.flags = vlb.SKYBOX_ON_MISS
        got = c.bake_probes(s)
        osc = oa.Scene(empty)
        osc.set_skybox(sky)
        assert rel_l2(got, osc.bake_probes(s)[0]) <= PROBE_TOL


def test_bake_cube_reference_defaults_scaled(ctx, vlb, oa, scenes):
    # the reference's own scene (default cube), its 7x7x7 grid and default flags, direction grid
    # scaled down from 3141x1000 so the oracle finishes in seconds
    sc = scenes.default_cube()
    with vlb.Context(0) as c:
        c.set_scene(sc)
        sky = scenes.hdr_sky(64, 32, seed=6)
        c.set_skybox(sky)
        s = vlb.default_settings()
        s.dir_w, s.dir_h = 157, 50
        vlb.settings_from_bounds(s, c.scene_bounds(tight=False))
        assert np.allclose(c.scene_bounds(False), [-1, -1, -1, 1, 1, 1])
        assert np.allclose(list(s.step), [2 / 6] * 3)
        got = c.bake_probes(s)
        osc = oa.Scene(sc)
        osc.set_skybox(sky)
        assert rel_l2(got, osc.bake_probes(s)[0]) <= PROBE_TOL


def test_whole_probe_items_and_per_chunk_items_give_the_same_bits(vlb, scenes, monkeypatch):
    """bake_device's work decomposition (a probe as ONE work item whose warp adds the chunk sums in order, or one item per
    chunk + k_sum_partials adding them in the same order) never shows in the result: all-whole, all-partial and the default
    mix are bitwise equal. The grid is large enough (10,240 probes > 2 per resident warp) for whole-probe items to be used."""
    import torch
    sc = scenes.small_room()
    sky = scenes.hdr_sky(64, 32, seed=4)
    s = vlb.default_settings()
    s.probes[:] = (32, 32, 10); s.dir_w, s.dir_h = 64, 64; s.sh_order = 2; s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    vlb.settings_from_bounds(s, (0.3, 0.3, 0.3, 3.7, 3.7, 3.7))
    with vlb.Context(0) as c:
        c.set_scene(sc); c.build_bvh(); c.set_skybox(sky)
        outs = {}
        for mode in ("default", "0", "-1"):
            if mode == "default":
                monkeypatch.delenv("VLB_BAKE_TAIL_WAVES", raising=False)
            else:
                monkeypatch.setenv("VLB_BAKE_TAIL_WAVES", mode)
            out = torch.full((s.n_probes, 48), -1.0, device="cuda")
            c.bake_probes_device(s, out.data_ptr()); c.synchronize()
            outs[mode] = out.cpu().numpy()
        assert np.abs(outs["default"]).sum() > 0
        assert np.array_equal(outs["0"], outs["default"]) and np.array_equal(outs["-1"], outs["default"])
        # and a few-probe call (all per-chunk items whatever the setting) reproduces its rows of the large one
        monkeypatch.delenv("VLB_BAKE_TAIL_WAVES", raising=False)
        t = s.copy(); t.slab_k0, t.slab_k1 = 3, 4
        part = torch.zeros((32 * 32, 48), device="cuda")
        c.bake_probes_device(t, part.data_ptr()); c.synchronize()
        assert np.array_equal(part.cpu().numpy(), outs["default"][3 * 1024: 4 * 1024])


def test_ray_slot_policies_give_the_same_bits(vlb, scenes, monkeypatch):
    """How the warp schedules its ray slots -- interleaved refills (order 0 / 1), deferred shading in two phases per chunk
    (order 3, the default for long chunks), refill thresholds, per-direction tables or the same values computed in place -- never shows in the result: every policy gives the same coefficients bit for bit and the same shadow-ray count."""
    import torch
    sc = scenes.small_room()
    sky = scenes.hdr_sky(64, 32, seed=4)
    s = vlb.default_settings()
    s.probes[:] = (12, 12, 8); s.dir_w, s.dir_h = 64, 64; s.sh_order = 2; s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    vlb.settings_from_bounds(s, (0.3, 0.3, 0.3, 3.7, 3.7, 3.7))
    knobs = ("VLB_BAKE_REFILL_ORDER", "VLB_BAKE_REFILL_MIN", "VLB_BAKE_NODE_MIN", "VLB_BAKE_DIR_TABLES")
    policies = [{}, {"VLB_BAKE_REFILL_ORDER": "0"}, {"VLB_BAKE_REFILL_ORDER": "1"}, {"VLB_BAKE_REFILL_ORDER": "3"},
                {"VLB_BAKE_REFILL_ORDER": "3", "VLB_BAKE_REFILL_MIN": "1", "VLB_BAKE_NODE_MIN": "1"},
                {"VLB_BAKE_REFILL_ORDER": "1", "VLB_BAKE_REFILL_MIN": "32", "VLB_BAKE_NODE_MIN": "16"},
                {"VLB_BAKE_DIR_TABLES": "0"}, {"VLB_BAKE_DIR_TABLES": "0", "VLB_BAKE_REFILL_ORDER": "1"}]
    with vlb.Context(0) as c:
        c.set_scene(sc); c.build_bvh(); c.set_skybox(sky)
        ref = None
        for pol in policies:
            for k in knobs:
                monkeypatch.delenv(k, raising=False)
            for k, v in pol.items():
                monkeypatch.setenv(k, v)
            out = torch.full((s.n_probes, 48), -1.0, device="cuda")
            c.bake_probes_device(s, out.data_ptr()); c.synchronize()
            got = (out.cpu().numpy(), c.last_bake_stats().n_shadow_rays)
            if ref is None:
                ref = got
                assert np.abs(ref[0]).sum() > 0 and ref[1] > 0
            assert np.array_equal(got[0], ref[0]) and got[1] == ref[1], pol


def test_scene_with_an_index_past_its_vertex_range_is_refused(vlb, scenes):
    """Index validation and the per-instance bounds run on the device (k_instance_checks): a bad index fails the call with
    VLB_ERR_INVALID (never reads outside the vertex array), the context stays usable, and the reference-mode bounds of a
    good scene are the ones the host loop used to compute."""
    good = scenes.small_room()
    with vlb.Context(0) as c:
        bad = {k: np.array(v, copy=True) for k, v in good.items() if k in ("vertices", "indices", "instances", "materials")}
        inst = bad["instances"][0]
        bad["indices"][int(inst["first_index"]) + 1] = int(inst["vertex_count"]) + 7
        with pytest.raises(vlb.VlbError) as e:
            c.set_scene(bad)
        assert e.value.code == vlb.ERR_INVALID and "index out of range" in str(e.value)
        with pytest.raises(vlb.VlbError):
            c.build_bvh()                                   # no scene is set after the failure
        c.set_scene(good)
        c.build_bvh()
        ref = c.scene_bounds(tight=False)
        # host restatement of Scene_t::loadNode's bounds (scene_manager.cpp:497-507)
        lo, hi = np.zeros(3), np.zeros(3)
        for it in good["instances"]:
            v = good["vertices"]["position"][int(it["first_vertex"]): int(it["first_vertex"]) + int(it["vertex_count"]), :3]
            m = np.asarray(it["transform"], np.float32).reshape(3, 4)
            a = m[:, :3] @ v.min(0) + m[:, 3]; b = m[:, :3] @ v.max(0) + m[:, 3]
            lo = np.minimum(lo, a); hi = np.maximum(hi, b)
        assert np.allclose(ref[:3], lo, atol=1e-5) and np.allclose(ref[3:], hi, atol=1e-5)


@pytest.mark.parametrize("radius", [4, 16, 64])
def test_ploc_builder_same_hits_same_bake_same_tree_as_host_twin(vlb, scenes, radius):
    """vlb_bvh_set_builder(PLOC): the agglomerative builder (csrc/vlb_ploc.cuh, all rounds in one cooperative kernel) gives
    another tree, never another result: hit ids / (t, u, v) identical to the LBVH and to brute force, the bake bitwise
    identical; and the tree is the one the serial host twin of the same code builds (tests/emu), wide node for wide node."""
    import emu_api
    sc = scenes.atrium(20000, seed=3)
    sky = scenes.hdr_sky(64, 32, seed=2)
    rng = np.random.default_rng(11)
    n = 20000
    o = (rng.uniform(0.03, 0.97, (n, 3)) * np.array(scenes.HALL)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    s = scenes.atrium_settings(probes=(5, 3, 4), dirs=(32, 32), order=3, bounds=(0, 0, 0) + tuple(scenes.HALL))
    with vlb.Context(0) as c:
        c.set_scene(sc); c.set_skybox(sky)
        lb = c.build_bvh()
        ids0, tuv0 = c.trace_rays(o, d)
        any0, _ = c.trace_rays(o, d, tmin=0.0, tmax=4.0, kind=vlb.TRACE_ANY)
        bake0 = c.bake_probes(s)
        c.set_bvh_builder("ploc", radius)
        pl = c.build_bvh()
        ids1, tuv1 = c.trace_rays(o, d)
        idsb, tuvb = c.trace_rays(o, d, accel=vlb.TRACE_BRUTE_FORCE)
        any1, _ = c.trace_rays(o, d, tmin=0.0, tmax=4.0, kind=vlb.TRACE_ANY)
        assert (ids0 >= 0).mean() > 0.5
        assert np.array_equal(ids1, ids0) and np.array_equal(tuv1, tuv0)
        assert np.array_equal(ids1, idsb) and np.array_equal(tuv1, tuvb)
        assert np.array_equal(any1 >= 0, any0 >= 0)
        assert np.array_equal(c.bake_probes(s), bake0)
        assert pl.n_triangles == lb.n_triangles and pl.n_nodes != lb.n_nodes
        twin = emu_api.Scene(sc, max_leaf=int(pl.max_leaf_size), builder="ploc", ploc_radius=radius)
        assert int(twin.node_stats()[0]) == int(pl.n_nodes)
        c.set_bvh_builder("lbvh")
        assert c.build_bvh().n_nodes == lb.n_nodes
        with pytest.raises(vlb.VlbError):
            c.set_bvh_builder("ploc", 65)


def test_bake_probes_multi_one_process_several_contexts(vlb, scenes, room):
    """vlb_bake_probes_multi: one host process, n contexts (here all on device 0, which exercises the same threads,
    cyclic shares and strided copies as n GPUs) == one context, bit for bit; with and without gather passes."""
    sc, _ = room
    sky = scenes.hdr_sky(64, 32, seed=2)
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 5); s.dir_w, s.dir_h = 32, 16; s.sh_order = 3; s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    vlb.settings_from_bounds(s, (0.3, 0.3, 0.3, 3.7, 3.7, 3.7))
    ctxs = [vlb.Context(0) for _ in range(3)]
    try:
        for c in ctxs:
            c.set_scene(sc); c.build_bvh(); c.set_skybox(sky)
        want = ctxs[0].bake_probes(s)
        for n in (1, 2, 3):
            assert np.array_equal(vlb.bake_probes_multi(ctxs[:n], s), want), n
        s.bounces, s.indirect_gain = 2, 0.7
        want2 = ctxs[0].bake_probes(s)
        assert not np.array_equal(want2, want)
        for n in (2, 3):
            assert np.array_equal(vlb.bake_probes_multi(ctxs[:n], s), want2), n
        s.bounces = 0
        with pytest.raises(vlb.VlbError):
            vlb.bake_probes_multi([ctxs[0], ctxs[0]], s)
        t = s.copy(); t.slab_k0, t.slab_k1 = 0, 2
        with pytest.raises(vlb.VlbError):
            vlb.bake_probes_multi(ctxs[:2], t)
    finally:
        for c in ctxs:
            c.close()


def test_skybox_chained_device_projections_on_own_stream(vlb, oa, scenes):
    """Back-to-back vlb_skybox_project_sh_device calls on the ctx's own stream are chained with programmatic dependent
    launch (the next launch streams its texels while the previous one reduces and retires): every result must still be
    exact, also when other calls (which close the chain) are interleaved and when the shape changes in between."""
    import torch
    with vlb.Context(0) as c:
        W, H = 640, 160
        host = [scenes.hdr_sky(W, H, seed=500 + i) for i in range(6)]
        dev = [torch.from_numpy(m).cuda() for m in host]
        small = scenes.hdr_sky(128, 64, seed=77)
        dsmall = torch.from_numpy(small).cuda()
        out = torch.full((8, 48), -1.0, device="cuda")
        torch.cuda.synchronize()
        want = [oa.skybox_project(m, 3) for m in host]
        for rep in range(3):
            out.fill_(-1.0)
            torch.cuda.synchronize()
            for i in range(6):                      # chain of six launches
                c.skybox_project_sh_device(dev[i].data_ptr(), H * W * 16, 1, vlb.FMT_RGBA32F, W, H, 3, out[i].data_ptr())
            c.skybox_project_sh_device(dsmall.data_ptr(), 64 * 128 * 16, 1, vlb.FMT_RGBA32F, 128, 64, 3, out[6].data_ptr())   # new shape: tables re-uploaded
            c.skybox_project_sh_device(dev[0].data_ptr(), H * W * 16, 1, vlb.FMT_RGBA32F, W, H, 2, out[7].data_ptr())
            c.synchronize()
            got = out.cpu().numpy().reshape(8, 16, 3)
            for i in range(6):
                assert rel_l2(got[i], want[i]) <= SKY_TOL
                assert np.array_equal(got[i], c.skybox_project_sh(host[i], 3))          # the host path (never chained), bit for bit
            assert rel_l2(got[6], oa.skybox_project(small, 3)) <= SKY_TOL
            assert rel_l2(got[7], oa.skybox_project(host[0], 2)) <= SKY_TOL and np.all(got[7, 9:] == 0)
        # the same output buffer written by consecutive launches: the last one wins
        for i in range(6):
            c.skybox_project_sh_device(dev[i].data_ptr(), H * W * 16, 1, vlb.FMT_RGBA32F, W, H, 3, out[0].data_ptr())
        c.synchronize()
        assert rel_l2(out[0].cpu().numpy(), want[5]) <= SKY_TOL

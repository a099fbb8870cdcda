"""GPU parity on the configurations BASELINE.json names (the ones the metric is quoted on), not on toy scenes:

  C2  atrium seed 7, 262,144 triangles, 16x8x16 probes x 1,024 rays (32x32), L2, shadow rays + skybox
  C3  same scene, 64x32x64 probes x 4,096 rays (64x64)
  C4  atrium seed 11, 3,145,728 triangles, 32x16x32 probes x 4,096 rays, L3, gather passes
  C5  one 4096x2048 RGBA32F equirect -> L2 / L3 SH

The CUDA path (through the C ABI) bakes whole z-slices / the whole grid at full size; the CPU oracle bakes a strided
sample of the same probes. Tolerances are BASELINE.json's: <= 1e-3 max relative L2 per probe SH vector, <= 1e-4 per
skybox SH vector, hit ids bit-exact (LBVH == brute-force CUDA == oracle), shadow-ray counts equal.
Reference path under test: shaders/env_map.rgen:18-28, env_map.rchit:51-102, main.rmiss:18-41, sh.comp:25-41,
skybox_sh.comp:25-41, main.rchit:124-163."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

PROBE_TOL = 1e-3
SKY_TOL = 1e-4
N_TRIS = 262144


@pytest.fixture(scope="module")
def atrium(vlb, oa, scenes):
    sc = scenes.atrium(N_TRIS, seed=7)
    sky = scenes.hdr_sky(2048, 1024, seed=1)
    c = vlb.Context(0)
    c.set_scene(sc)
    st = c.build_bvh()
    assert st.n_triangles == N_TRIS
    c.set_skybox(sky)
    osc = oa.Scene(sc)
    osc.set_skybox(sky)
    yield c, osc
    c.close()
    osc.close()


def _settings(scenes, probes, dirs, order=2):
    return scenes.atrium_settings(probes=probes, dirs=dirs, order=order, bounds=(0, 0, 0) + tuple(scenes.HALL))


def test_c2_whole_grid_sample_vs_oracle(atrium, vlb, scenes):
    c, osc = atrium
    s = _settings(scenes, (16, 8, 16), (32, 32))
    assert s.flags & vlb.SHADOW_RAYS and s.flags & vlb.SKYBOX_ON_MISS and s.flags & vlb.SRGB_ENCODE
    got = c.bake_probes(s)
    ids = np.arange(5, s.n_probes, 16, dtype=np.int64)           # 128 strided probes of the 2,048
    assert len(ids) >= 64
    ref, _ = osc.bake_probes(s, probe_ids=ids)
    assert rel_l2(got[ids], ref) <= PROBE_TOL
    assert np.all(got[:, 9:] == 0)


@pytest.mark.parametrize("k", [0, 7, 15])
def test_c2_slice_values_and_shadow_ray_count_equal_oracle(atrium, vlb, scenes, k):
    # one whole z-slice (128 probes x 1,024 rays) on both sides: every value and the exact number of shadow rays
    c, osc = atrium
    s = _settings(scenes, (16, 8, 16), (32, 32))
    p = s.copy()
    p.slab_k0, p.slab_k1 = k, k + 1
    got = c.bake_probes(p)
    st = c.last_bake_stats()
    nxy = 16 * 8
    ids = np.arange(k * nxy, (k + 1) * nxy, dtype=np.int64)
    ref, n_shadow = osc.bake_probes(s, probe_ids=ids)
    assert rel_l2(got, ref) <= PROBE_TOL
    assert st.n_primary_rays == nxy * 1024
    assert st.n_shadow_rays == n_shadow


@pytest.mark.parametrize("k", [3, 40])
def test_c3_slice_sample_vs_oracle(atrium, vlb, scenes, k):
    # C3: one z-slice of the 64x32x64 grid = 2,048 probes x 4,096 rays on the GPU, 96 strided probes on the oracle
    c, osc = atrium
    s = _settings(scenes, (64, 32, 64), (64, 64))
    p = s.copy()
    p.slab_k0, p.slab_k1 = k, k + 1
    got = c.bake_probes(p)
    nxy = 64 * 32
    local = np.arange(3, nxy, 21, dtype=np.int64)[:96]
    ref, _ = osc.bake_probes(s, probe_ids=k * nxy + local)
    assert rel_l2(got[local], ref) <= PROBE_TOL


def test_c3_small_block_shadow_ray_count_equals_oracle(atrium, vlb, scenes):
    # the C3 direction grid (4,096 rays) on a 4x4x2 corner block of the C3 lattice: shadow-ray counts are exact
    c, osc = atrium
    s = _settings(scenes, (64, 32, 64), (64, 64))
    b = s.copy()
    b.probes[:] = (4, 4, 2)
    b.origin[:] = [s.origin[d] + off * s.step[d] for d, off in enumerate((20, 9, 30))]
    got = c.bake_probes(b)
    st = c.last_bake_stats()
    ref, n_shadow = osc.bake_probes(b)
    assert rel_l2(got, ref) <= PROBE_TOL
    assert st.n_shadow_rays == n_shadow and st.n_primary_rays == 32 * 4096


def _c3_rays(vlb, scenes, n_probes, seed):
    """Rays as the C3 bake fires them: origins = probe positions of the 64x32x64 lattice, directions = texel centres of
    the 64x64 equirect grid (env_map.rgen:21), a random (probe, texel) sample."""
    s = _settings(scenes, (64, 32, 64), (64, 64))
    pos = vlb.probe_positions(s)
    rng = np.random.default_rng(seed)
    pid = rng.choice(len(pos), n_probes, replace=False)
    W, H = 64, 64
    x = (np.arange(W) + 0.5) * (2 * np.pi / W)
    y = (np.arange(H) + 0.5) * (np.pi / H)
    st, ct = np.sin(y)[:, None], np.cos(y)[:, None]
    t = np.stack([st * np.cos(x)[None], st * np.sin(x)[None], np.broadcast_to(ct, (H, W))], -1)      # toVector
    d = t[..., [0, 2, 1]].reshape(-1, 3).astype(np.float32)                                            # .xzy
    o = np.repeat(pos[pid], 128, axis=0).astype(np.float32)
    sel = rng.integers(0, W * H, len(o))
    return o, np.ascontiguousarray(d[sel])


def test_c3_probe_rays_hit_ids_bvh_equals_brute_force_equals_oracle(atrium, vlb, scenes):
    c, osc = atrium
    o, d = _c3_rays(vlb, scenes, 1600, seed=5)                  # 204,800 rays
    assert len(o) >= 200000
    iv, tv = c.trace_rays(o, d, tmin=0.001, tmax=10000.0, accel=vlb.TRACE_BVH)
    ib, tb = c.trace_rays(o, d, tmin=0.001, tmax=10000.0, accel=vlb.TRACE_BRUTE_FORCE)
    assert np.array_equal(iv, ib) and np.array_equal(tv, tb)
    assert 0.5 < (iv >= 0).mean() <= 1.0
    # the oracle (own binned-SAH BVH, same intersect routine) on a 30,000-ray subset
    sub = np.arange(0, len(o), 7)[:30000]
    io, to = osc.trace_rays(o[sub], d[sub], tmin=0.001, tmax=10000.0, accel=1)
    assert np.array_equal(iv[sub], io) and np.array_equal(tv[sub], to)
    # shadow rays of those hits towards the C3 light: any-hit occlusion agrees with brute force
    hit = iv >= 0
    P = o[hit] + d[hit] * tv[hit, :1]
    L = np.asarray(scenes.ATRIUM_LIGHT, np.float32)[None] - P
    ln = np.linalg.norm(L, axis=1).astype(np.float32)
    keep = ln > 1e-3
    so = (P - d[hit] * 0.01)[keep][:100000].astype(np.float32)
    sd = (L / ln[:, None])[keep][:100000].astype(np.float32)
    tm = float(ln[keep][:100000].min())
    a, _ = c.trace_rays(so, sd, tmin=0.0, tmax=tm, accel=vlb.TRACE_BVH, kind=vlb.TRACE_ANY)
    b, _ = c.trace_rays(so, sd, tmin=0.0, tmax=tm, accel=vlb.TRACE_BRUTE_FORCE, kind=vlb.TRACE_ANY)
    assert np.array_equal(a >= 0, b >= 0)


def test_c4_scene_direct_and_gather_pass_sample_vs_oracle(vlb, oa, scenes):
    # C4: 3,145,728 triangles (seed 11), the 32x16x32 grid, 64x64 directions, L3: direct pass + one gather pass on the
    # GPU over two z-slices each; the oracle on a strided sample of those probes.
    import torch
    sc = scenes.atrium(3 * (1 << 20), seed=11)
    sky = scenes.hdr_sky(512, 256, seed=1)
    s = _settings(scenes, (32, 16, 32), (64, 64), order=3)
    s.indirect_gain = 1.0
    nxy = 32 * 16
    with vlb.Context(0) as c:
        c.set_scene(sc)
        st = c.build_bvh()
        assert st.n_triangles == 3 * (1 << 20)
        c.set_skybox(sky)
        # the gather source over the WHOLE grid: a cheap direct bake with an 8x8 direction grid
        coarse = s.copy()
        coarse.dir_w, coarse.dir_h = 8, 8
        prev = torch.from_numpy(c.bake_probes(coarse).reshape(-1, 48)).cuda()
        got = {}
        for k in (5, 20):
            p = s.copy()
            p.slab_k0, p.slab_k1 = k, k + 1
            direct = c.bake_probes(p)
            out = torch.zeros((nxy, 48), device="cuda")
            c.bake_gather_device(p, prev.data_ptr(), out.data_ptr())
            c.synchronize()
            got[k] = (direct, out.cpu().numpy())
        prev_h = prev.cpu().numpy()
    osc = oa.Scene(sc)
    osc.set_skybox(sky)
    local = np.arange(1, nxy, 37, dtype=np.int64)               # 14 probes per slice
    for k, (direct, gathered) in got.items():
        ids = k * nxy + local
        ref_d, _ = osc.bake_probes(s, probe_ids=ids)
        ref_g, _ = osc.bake_gather(s, prev_h, probe_ids=ids)
        assert rel_l2(direct[local], ref_d) <= PROBE_TOL
        assert rel_l2(gathered[local], ref_g) <= PROBE_TOL
        assert rel_l2(gathered[local], direct.reshape(-1, 48)[local]) > 1e-3     # the gather term is really there
    osc.close()


@pytest.mark.parametrize("order", [2, 3])
def test_c5_largest_map_4096x2048_vs_oracle(ctx, oa, scenes, order):
    img = scenes.hdr_sky(4096, 2048, seed=100)
    got = ctx.skybox_project_sh(img, order)
    assert rel_l2(got, oa.skybox_project(img, order)) <= SKY_TOL
    if order == 2:
        assert np.all(got[9:] == 0)


@pytest.mark.parametrize("shape", [(256, 512), (512, 1024)])
def test_c5_batched_maps_vs_oracle(ctx, oa, scenes, shape):
    # the smaller C5 sizes as a batch in one launch; a strided sample of the maps against the oracle
    n = 24
    maps = [scenes.hdr_sky(shape[1], shape[0], seed=100 + i) for i in range(n)]
    got = ctx.skybox_project_sh_batched(maps, 3)
    for i in (0, 5, 11, 23):
        assert rel_l2(got[i], oa.skybox_project(maps[i], 3)) <= SKY_TOL

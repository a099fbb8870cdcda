"""Pins the oracle (oracle/vlb_oracle.cpp) to the reference: its SH basis and equirect maps are
compared bit-for-bit with (a) golden vectors generated from the reference's own
shaders/sh_common.h and (b) that header compiled live when /root/reference is present. Plus the
analytic known-answer tests of SURVEY.md §4 and the probe-grid restatement of
src/baker/light_baker.cpp:80-101."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sh_common_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_basis_matches_golden_bit_exact(oa, gold):
    got = oa.sh_basis(gold["dirs"])
    assert np.array_equal(got.view(np.uint32), gold["basis"].view(np.uint32))


def test_equirect_maps_match_golden_bit_exact(oa, gold):
    assert np.float32(3.1415926538) == gold["pi"]
    for i, (w, h) in enumerate(gold["sizes"]):
        phi = np.array([oa.lib().vo_x2phi(x, int(w)) for x in range(min(w, 64))], np.float32)
        th = np.array([oa.lib().vo_y2theta(y, int(h)) for y in range(min(h, 64))], np.float32)
        assert np.array_equal(phi, gold["phi_%d" % i]) and np.array_equal(th, gold["theta_%d" % i])


def test_to_vector_matches_golden(oa, gold):
    # the reference header uses fp32 libm sin/cos; the oracle rounds double sin/cos once (GLSL
    # leaves the precision open) -> agreement to fp32 rounding, not bitwise
    import ctypes
    out = np.zeros(3, np.float32)
    for a, v in zip(gold["angles"], gold["to_vector"]):
        oa.lib().vo_to_vector(float(a[0]), float(a[1]), out.ctypes.data_as(ctypes.c_void_p))
        assert np.abs(out - v).max() <= 2.5e-7


def test_basis_matches_live_reference_header(oa):
    R = oa.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(3)
    d = rng.normal(size=(300, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    got = oa.sh_basis(d)
    for n in range(len(d)):
        i = 0
        for l in range(5):
            for m in range(-l, l + 1):
                assert np.float32(R.ref_SH(l, m, float(d[n, 0]), float(d[n, 1]), float(d[n, 2]))) == got[n, i]
                i += 1
    assert R.ref_SH(5, 0, 0.0, 0.0, 1.0) == 0.0     # sh_common.h:222


def test_constant_image_kat(oa):
    # c00 = 2 sqrt(pi) rgb, all other coefficients ~ 0 (quadrature residual ~3e-6)
    img = np.ones((256, 512, 4), np.float32)
    img[..., 2] = 0.25
    for fn in (oa.skybox_project, oa.envmap_project):
        c = fn(img, 3)
        assert np.allclose(c[0], 2 * np.sqrt(np.pi) * np.array([1, 1, 0.25]), rtol=3e-5)
        assert np.abs(c[1:]).max() < 1e-4   # quadrature residual of the 512x256 grid


def test_gram_matrix_is_identity(oa):
    # projecting the image of basis function j gives the unit vector e_j (sh.comp quadrature)
    W, H = 256, 128
    t, _, _ = oa.probe_dirs(W, H)
    b = oa.sh_basis(t.reshape(-1, 3)).reshape(H, W, 25)
    for j in range(16):
        img = np.zeros((H, W, 4), np.float32)
        img[..., 0] = b[..., j]
        c = oa.envmap_project(img, 3)[:, 0]
        e = np.zeros(16)
        e[j] = 1
        assert np.abs(c - e).max() < 2e-4


def test_project_reconstruct_round_trip(oa):
    # sh_sum.comp reconstruction of a band-limited signal is the signal
    rng = np.random.default_rng(5)
    coeffs = np.zeros((16, 3), np.float32)
    coeffs[:9] = rng.normal(size=(9, 3))
    W, H = 128, 64
    img = np.ones((H, W, 4), np.float32)
    img[..., :3] = oa.sh_reconstruct(coeffs, 2, W, H)
    back = oa.envmap_project(img, 3)
    assert np.abs(back - coeffs).max() < 1e-3


def test_rgba8_is_value_over_255(oa):
    rng = np.random.default_rng(7)
    u8 = rng.integers(0, 256, (32, 64, 4), dtype=np.uint8)
    f = (u8.astype(np.float32) / np.float32(255.0))
    assert np.allclose(oa.skybox_project(u8, 3), oa.skybox_project(f, 3), rtol=0, atol=1e-7)


def test_skybox_frame_differs_from_envmap_frame(oa, scenes):
    # skybox_sh.comp shifts phi by -pi/2 and swizzles .xzy; sh.comp does neither
    img = scenes.hdr_sky(128, 64, seed=2)
    a, b = oa.skybox_project(img, 3), oa.envmap_project(img, 3)
    assert np.allclose(a[0], b[0], rtol=1e-6)          # band 0 is frame independent
    assert np.abs(a[1:4] - b[1:4]).max() > 1e-2


def test_probe_positions_literal_vs_separable(oa, vlb):
    # light_baker.cpp:80-101 literally (vector doubling) == per-axis repeated adds + writer-order map
    for counts, bounds in (((7, 7, 7), (-1, -1, -1, 1, 1, 1)), ((4, 2, 5), (0.1, -3, 2, 30.3, 12.7, 18.9)),
                           ((2, 3, 2), (-5, 0, 0, 5, 1, 7))):
        lit, step = oa.probe_positions_literal(bounds, counts)
        s = vlb.default_settings()
        s.probes[:] = counts
        vlb.settings_from_bounds(s, bounds)
        assert np.array_equal(np.array(list(s.step), np.float32), step)
        s.flags |= vlb.REFERENCE_PROBE_ORDER
        assert np.array_equal(oa.probe_positions(s), lit)
        assert np.array_equal(vlb.probe_positions(s), lit)          # host helper of the C ABI
        s.flags &= ~vlb.REFERENCE_PROBE_ORDER
        x_fast = vlb.probe_positions(s)
        assert np.array_equal(x_fast, oa.probe_positions(s))
        assert sorted(map(tuple, x_fast)) == sorted(map(tuple, lit))
        # x-fastest: index = i + j*Nx + k*Nx*Ny (consumer: shaders/sh.rmiss:28)
        assert np.all(np.diff(x_fast[: counts[0], 0]) > 0) and x_fast[0, 1] == x_fast[counts[0] - 1, 1]


def test_cube_bounds_and_grid_step(oa, scenes):
    # default cube: bounds [(-1,-1,-1),(1,1,1)] => gridStep 2/6 for the 7^3 grid
    o = oa.Scene(scenes.default_cube())
    assert o.n_triangles == 12
    assert np.array_equal(o.bounds(False), np.array([-1, -1, -1, 1, 1, 1], np.float32))
    assert np.array_equal(o.bounds(True), np.array([-1, -1, -1, 1, 1, 1], np.float32))


def test_reference_bounds_quirk_includes_origin(oa, scenes):
    # scene_manager.cpp:497-507: bounds start at {0,0,0} -> a scene away from the origin still
    # reports the origin inside its bounds; the tight AABB does not
    sc = scenes.default_cube()
    sc["instances"]["transform"][0] = scenes.trs12((10, 10, 10))
    o = oa.Scene(sc)
    assert np.array_equal(o.bounds(False), np.array([0, 0, 0, 11, 11, 11], np.float32))
    assert np.array_equal(o.bounds(True), np.array([9, 9, 9, 11, 11, 11], np.float32))


def test_oracle_bvh_equals_brute_force(oa, scenes):
    sc = scenes.small_room()
    o = oa.Scene(sc)
    rng = np.random.default_rng(1)
    org = rng.uniform(0.1, 3.9, (20000, 3)).astype(np.float32)
    d = rng.normal(size=(20000, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    ib, tb = o.trace_rays(org, d, accel=1)
    iv, tv = o.trace_rays(org, d, accel=0)
    assert np.array_equal(ib, iv) and np.array_equal(tb, tv)
    a = o.trace_rays(org, d, tmin=0.0, tmax=1.0, accel=1, kind=1)[0] >= 0
    b = o.trace_rays(org, d, tmin=0.0, tmax=1.0, accel=0, kind=1)[0] >= 0
    assert np.array_equal(a, b)


def test_oracle_bake_bvh_equals_brute_force(oa, vlb, scenes):
    sc = scenes.small_room()
    o = oa.Scene(sc)
    o.set_skybox(scenes.hdr_sky(32, 16, seed=1))
    s = vlb.default_settings()
    s.probes[:] = (2, 2, 2)
    s.dir_w, s.dir_h = 16, 8
    s.light_pos[:] = (2.0, 3.5, 2.0)
    vlb.settings_from_bounds(s, o.bounds(True))
    a, na = o.bake_probes(s, brute=False)
    b, nb = o.bake_probes(s, brute=True)
    assert np.array_equal(a, b) and na == nb


def test_probe_envmap_shading_kat(oa, vlb, scenes):
    # a probe at the centre of the closed default cube with a light inside it. The cube's normals
    # point outwards, i.e. away from an interior light: sDotN == 0 -> black (env_map.rchit:79-90).
    # Flipping the normals leaves the positions alone and lights every texel: radiance is the Phong
    # term of env_map.rchit, sRGB-encoded, within (0, 1].
    sc = scenes.default_cube()
    sc["materials"]["base_color_factor"][0] = (1, 1, 1, 1)
    s = vlb.default_settings()
    s.dir_w, s.dir_h = 32, 16
    s.light_pos[:] = (0.0, 0.5, 0.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SRGB_ENCODE
    img, _ = oa.Scene(sc).probe_envmap(s, (0.0, 0.0, 0.0))
    assert img.max() == 0.0
    sc["vertices"]["normal"] *= -1
    img2, sh = oa.Scene(sc).probe_envmap(s, (0.0, 0.0, 0.0))
    assert img2.min() > 0.0 and img2.max() <= 1.0 + 1e-6
    assert sh[0].min() > 0.0

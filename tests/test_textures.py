"""Textured baseColor: getBaseColor's texture branch (shaders/env_map.rchit:36-49, uv0
interpolation :65) over textures set as Scene_t::loadTextures / loadSamplers create them
(src/scene_manager.cpp:941-973, 650-690). The driver's filter arithmetic is "parity unpinned", so the oracle
(oracle/vlb_oracle.cpp tex_sample) is the specification: unnormalised coordinate u*W-0.5, address modes of
VkSamplerAddressMode, fp32 lerps. CPU: oracle KATs + the device sampler's host twin (tests/emu) against the
oracle. GPU (-m gpu): bake of the textured room through the C ABI against the oracle."""
import importlib

import numpy as np
import pytest

import emu_api
from conftest import rel_l2


def _one_texture_scene(scenes, vlb, texels, **sampler):
    sc = scenes.small_room()
    sc["textures"] = [dict(texels=texels, **sampler)]
    return sc


def test_oracle_bilinear_kats(vlb, oa, scenes):
    t = np.zeros((2, 2, 4), np.uint8)
    t[0, 0, :3] = (0, 0, 0); t[0, 1, :3] = (255, 0, 0); t[1, 0, :3] = (0, 255, 0); t[1, 1, :3] = (255, 255, 255)
    o = oa.Scene(_one_texture_scene(scenes, vlb, t, wrap_u=vlb.WRAP_CLAMP_TO_EDGE, wrap_v=vlb.WRAP_CLAMP_TO_EDGE))
    got = o.tex_sample(0, [(0.5, 0.5), (0.25, 0.25), (0.75, 0.25), (-3.0, 0.25), (0.5, 0.25), (9.0, 9.0)])
    want = [(0.5, 0.5, 0.25), (0, 0, 0), (1, 0, 0), (0, 0, 0), (0.5, 0, 0), (1, 1, 1)]
    assert np.allclose(got, want, atol=1e-6)
    # repeat: the left neighbour of texel 0 is texel W-1; mirror: it is texel 0 again
    o = oa.Scene(_one_texture_scene(scenes, vlb, t, wrap_u=vlb.WRAP_REPEAT, wrap_v=vlb.WRAP_REPEAT))
    assert np.allclose(o.tex_sample(0, [(0.0, 0.25)]), [(0.5, 0, 0)], atol=1e-6)
    o = oa.Scene(_one_texture_scene(scenes, vlb, t, wrap_u=vlb.WRAP_MIRRORED_REPEAT, wrap_v=vlb.WRAP_REPEAT))
    assert np.allclose(o.tex_sample(0, [(0.0, 0.25), (1.25, 0.25)]), [(0, 0, 0), (1, 0, 0)], atol=1e-6)
    # nearest: texel = floor(u * W)
    o = oa.Scene(_one_texture_scene(scenes, vlb, t, filter=vlb.FILTER_NEAREST))
    assert np.allclose(o.tex_sample(0, [(0.49, 0.1), (0.51, 0.1), (0.1, 0.9), (1.6, -0.4)]),
                       [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 1)], atol=1e-6)


def test_device_sampler_host_twin_equals_oracle(vlb, oa, scenes):
    sc = scenes.small_room_textured()
    o, e = oa.Scene(sc), emu_api.Scene(sc)
    rng = np.random.default_rng(2)
    uv = rng.uniform(-2.5, 3.5, (20000, 2)).astype(np.float32)
    uv[:64] = np.round(uv[:64] * 4) / 4                      # exact texel edges
    for k in range(len(sc["textures"])):
        assert np.array_equal(o.tex_sample(k, uv), e.tex_sample(k, uv)), k


def _settings(vlb, scenes, order=3):
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 3)
    s.dir_w, s.dir_h = 32, 16
    s.sh_order = order
    s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SRGB_ENCODE
    vlb.settings_from_bounds(s, (0.3, 0.3, 0.3, 3.7, 3.7, 3.7))
    return s


def test_textured_bake_host_twin_vs_oracle(vlb, oa, scenes):
    sc = scenes.small_room_textured()
    s = _settings(vlb, scenes)
    ref, _ = oa.Scene(sc).bake_probes(s)
    got = emu_api.Scene(sc).bake(s)
    assert rel_l2(got, ref) <= 1e-5
    # and the textures matter: the factor-only room differs
    plain, _ = oa.Scene(scenes.small_room()).bake_probes(s)
    assert rel_l2(plain, ref) > 1e-2


def test_constant_texture_equals_factor(vlb, oa, scenes):
    """A 1x1 texture of colour c gives what baseColorFactor = c gives (both branches of getBaseColor)."""
    sc = scenes.small_room()
    c = np.array([51, 102, 204, 255], np.uint8)
    tex = scenes.small_room()
    tex["materials"]["textures"][:, 2, 0] = 0
    tex["textures"] = [{"texels": c.reshape(1, 1, 4)}]
    sc["materials"]["base_color_factor"][:] = c / np.float32(255.0)
    s = _settings(vlb, scenes, order=2)
    a, _ = oa.Scene(sc).bake_probes(s)
    b, _ = oa.Scene(tex).bake_probes(s)
    assert rel_l2(b, a) <= 1e-6


# ------------------------------------------------------------------------------- GPU ------------
@pytest.mark.gpu
@pytest.mark.parametrize("order", [2, 3])
def test_textured_bake_vs_oracle(ctx, vlb, oa, scenes, order):
    sc = scenes.small_room_textured()
    sky = scenes.hdr_sky(64, 32)
    s = _settings(vlb, scenes, order)
    s.flags |= vlb.SKYBOX_ON_MISS
    ctx.set_scene(sc)
    ctx.build_bvh()
    ctx.set_skybox(sky)
    got = ctx.bake_probes(s)
    o = oa.Scene(sc)
    o.set_skybox(sky)
    ref, _ = o.bake_probes(s)
    assert rel_l2(got, ref) <= 1e-3                     # BASELINE tolerance per probe SH vector
    assert rel_l2(got, ref) <= 1e-5                     # what the exact-weights sampler actually achieves
    # the textures can be dropped and set again without re-uploading the geometry
    ctx.set_textures(sc["textures"])
    assert np.array_equal(ctx.bake_probes(s), got)


@pytest.mark.gpu
def test_texture_errors(ctx, vlb, scenes):
    sc = scenes.small_room_textured()
    s = _settings(vlb, scenes)
    ctx.set_scene(sc)
    ctx.build_bvh()
    ctx.set_textures(sc["textures"][:2])                # materials name textures 2 and 3 as well
    with pytest.raises(vlb.VlbError) as e:
        ctx.bake_probes(s)
    assert e.value.code == vlb.ERR_STATE and "texture" in str(e.value)
    bad = [{"texels": np.zeros((2, 2, 4), np.uint8), "wrap_u": 7}]
    with pytest.raises(vlb.VlbError):
        ctx.set_textures(bad)
    ctx.set_textures(sc["textures"])                    # the ctx stays usable
    assert np.isfinite(ctx.bake_probes(s)).all()

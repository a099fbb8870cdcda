"""Generates tests/golden/ref_shaders_golden.npz from the REFERENCE's own shader sources compiled as C++
(oracle/make_ref_shaders.py -> oracle/_ref/libvlb_refshaders.so): sRGB, dir2SkyboxUV, getBaseColor, the whole
dispatches of skybox_sh.comp / sh.comp, and whole probes through the reference's bake pipeline
(env_map.rgen -> env_map.rchit / main.rmiss / shadow.rmiss -> sh.comp; intersection by the oracle's BVH).
Run in the container where /root/reference is mounted; the committed .npz carries the pin to machines without it.

    python tests/golden/make_ref_shaders_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_oracle, oracle_api as oa  # noqa: E402


def room_case(vlb, scenes):
    """The small room as the parity tests bake it: 3x2x3 probes, 32x16 directions, its own light."""
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 3)
    s.dir_w, s.dir_h = 32, 16
    s.sh_order = 3
    s.light_pos[:] = (2.0, 3.5, 2.0)
    return scenes.small_room(), scenes.hdr_sky(64, 32, seed=4), s


def cube_case(vlb, scenes):
    """The reference's own scene and constants: default cube, light (1, 10, 1), a 3x3x3 lattice over its bounds."""
    s = vlb.default_settings()
    s.probes[:] = (3, 3, 3)
    s.dir_w, s.dir_h = 40, 20
    vlb.settings_from_bounds(s, (-1.9, -1.7, -1.8, 1.9, 1.7, 1.8))
    return scenes.default_cube(), scenes.hdr_sky(32, 16, seed=6), s


def viewer_gather_case(vlb, scenes):
    """The conditions under which the oracle's gather (its documented deviations from the literal shader, oracle/vlb_oracle.cpp:
    gather_indirect) and the viewer's main.rchit + sh.rmiss compute the same thing: a 7x7x7 grid with origin 0 (the shader
    hard-codes both) that contains the room with a cell to spare on every side (the room is moved by (1, 1, 1) into positive
    coordinates: the shader's floor(hitPosition / gridStep) indexes the buffer unclamped), every probe holding the SAME
    coefficients (the shader reads the cell's base probe for all eight corners), the SH argument in world frame (the shader
    passes hitNormal), no ambient term of the bake's own (the viewer has none), indirect_gain = ambient * 1250 (main.rchit:166).
    -> (scene, settings, the viewer's `ambient` push constant, prev [343,16,3], probe origins [4,3])"""
    s = vlb.default_settings()
    s.probes[:] = (7, 7, 7)
    s.origin[:] = (0.0, 0.0, 0.0)
    s.step[:] = (1.0, 1.0, 1.0)
    s.dir_w, s.dir_h = 48, 24
    s.sh_order = 3
    s.light_pos[:] = (3.0, 4.5, 3.0)
    s.ambient = 0.0
    amb = np.float32(0.0008)
    s.indirect_gain = float(amb * np.float32(1250.0))
    s.flags = vlb.SHADOW_RAYS | vlb.SRGB_ENCODE | vlb.SH_WORLD_FRAME
    rng = np.random.default_rng(21)
    one = (rng.uniform(0.05, 0.6, (16, 3)) * np.array([1.0] + [0.3] * 15)[:, None]).astype(np.float32)   # a positive DC, smaller bands
    prev = np.broadcast_to(one, (343, 16, 3)).copy()
    sc = scenes.small_room()
    sc = dict(sc, instances=sc["instances"].copy())
    sc["instances"]["transform"][:, [3, 7, 11]] += 1.0                 # the room, one cell into the grid
    origins = np.random.default_rng(22).uniform(1.5, 4.5, (4, 3)).astype(np.float32)
    return sc, s, float(amb), prev, origins


def viewer_gather_images(vlb, scenes, P, case):
    """One environment image per origin from env_map.rgen + the VIEWER's main.rchit + sh.rmiss (reference code)."""
    sc, s, amb, prev, origins = case
    sh25 = np.zeros((343, 25, 3), np.float32)
    sh25[:, :16] = prev
    imgs = []
    for o in origins:
        _sh, img, _n = P.bake_probe_viewer_hit(o, s.dir_w, s.dir_h, s.flags & ~vlb.SH_WORLD_FRAME, tuple(s.light_pos), tuple(s.step), 4,
                                               s.shadow_bias, amb, s.c_diffuse, s.c_specular, s.gloss, sh25)
        imgs.append(img[..., :3].copy())
    return np.stack(imgs)


def main():
    build_oracle.build_ref(force=True)
    vlb = importlib.import_module("vulkan-light-bakery_b200")
    scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
    R = oa.RefShaders()
    rng = np.random.default_rng(20261018)
    g = {}
    x = np.concatenate([rng.uniform(0, 2, 2000), rng.uniform(0, 0.01, 1000), [0.0, 0.0031308, 0.0031307, 0.0031309, 1.0, 0.5, 0.25, 4.0]])
    g["srgb_in"] = x.astype(np.float32).reshape(-1, 4)
    g["srgb_out"] = R.srgb_rchit(g["srgb_in"])
    assert np.array_equal(g["srgb_out"], R.srgb_rmiss(g["srgb_in"]))           # the two copies in the reference agree
    d = rng.normal(size=(2000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:6] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1]]
    g["uv_dirs"] = d.astype(np.float32)
    g["uv_out"] = R.dir2uv(g["uv_dirs"])
    g["default_light"] = R.default_light()
    for i, (W, H) in enumerate([(64, 32), (100, 50), (33, 17)]):
        sky = scenes.hdr_sky(W, H, seed=30 + i)
        g["proj_map_%d" % i] = sky
        g["proj_skybox_%d" % i] = R.skybox_sh(sky)
        g["proj_envmap_%d" % i] = R.envmap_sh(sky)
    u8 = rng.integers(0, 256, (24, 48, 4), dtype=np.uint8)
    g["proj_u8"] = u8
    g["proj_u8_skybox"] = R.skybox_sh(u8.astype(np.float32) / np.float32(255.0))
    base = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    for name, case in (("room", room_case), ("cube", cube_case)):
        sc, sky, s = case(vlb, scenes)
        osc = oa.Scene(sc)
        osc.set_skybox(sky)
        if name == "room":
            vlb.settings_from_bounds(s, osc.bounds(tight=True))
        P = oa.RefPipeline(sc, osc, sky)
        pos = vlb.probe_positions(s)
        for tag, flags in (("q", base | vlb.QUANTIZE_RGBA8), ("f", base), ("nosky", vlb.SHADOW_RAYS | vlb.SRGB_ENCODE)):
            coeffs = np.zeros((len(pos), 16, 3))
            shadow = 0
            for i, p in enumerate(pos):
                c, img, n = P.bake_probe(p, s.dir_w, s.dir_h, flags, tuple(s.light_pos))
                coeffs[i] = c
                shadow += n
                if i == 4:
                    g["bake_%s_%s_image4" % (name, tag)] = img[..., :3].copy()
            g["bake_%s_%s" % (name, tag)] = coeffs
            g["bake_%s_%s_shadow" % (name, tag)] = np.int64(shadow)
        P.close()
        osc.close()
    # the viewer's gather operator (main.rchit:124-167 + sh.rmiss), which the multi-bounce passes iterate
    case = viewer_gather_case(vlb, scenes)
    osc = oa.Scene(case[0])
    P = oa.RefPipeline(case[0], osc)
    g["gather_viewer_images"] = viewer_gather_images(vlb, scenes, P, case)       # NaN where the shader divides 0 / 0 (no corner visible)
    P.close()
    osc.close()
    out = os.path.join(ROOT, "tests", "golden", "ref_shaders_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

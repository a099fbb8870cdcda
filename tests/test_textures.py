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


# ---------------------------------------------------------------------- glTF texture ingest -----
@pytest.mark.parametrize("container,embed", [("gltf", True), ("gltf", False), ("glb", True)])
def test_gltf_textures_decode_to_the_source_texels(vlb, scenes, tmp_path, container, embed):
    """Scene_t::loadTextures / loadSamplers through the from-scratch PNG reader (csrc/png_decode.cpp): RGB and
    RGBA PNGs from data URIs, files and bufferViews come back as the RGBA8 texels and sampler state written."""
    sc = scenes.small_room_textured()
    if container == "glb":
        p = scenes.write_glb(sc, str(tmp_path / "room.glb"))
    else:
        p = scenes.write_gltf(sc, str(tmp_path / "room.gltf"), embed=embed)
    for k, want in enumerate(sc["textures"]):
        got = vlb.gltf_texture(p, k)
        assert got["used"]
        assert np.array_equal(got["texels"], want["texels"]), k
        assert (got["wrap_u"], got["wrap_v"], got["filter"]) == (want["wrap_u"], want["wrap_v"], want["filter"])
    with pytest.raises(vlb.VlbError):
        vlb.gltf_texture(p, len(sc["textures"]))


def test_png_reader_filters_palette_grey_and_errors(vlb, scenes, tmp_path):
    import io
    import json
    import base64
    from PIL import Image
    rng = np.random.default_rng(9)
    sc = scenes.small_room()
    sc["materials"]["textures"][0, 2, 0] = 0
    sc["vertices"]["uv0"][:] = 0.5
    base = scenes.write_gltf(sc, str(tmp_path / "base.gltf"))
    doc = json.load(open(base))

    def with_image(data, name):
        d = dict(doc)
        d["images"] = [{"uri": "data:image/png;base64," + base64.b64encode(data).decode()}]
        d["textures"] = [{"source": 0}]
        p = str(tmp_path / name)
        json.dump(d, open(p, "w"))
        return p

    def png(img, **kw):
        b = io.BytesIO()
        img.save(b, format="PNG", **kw)
        return b.getvalue()

    # smooth + noisy content so that the encoder picks all five row filters
    y, x = np.mgrid[0:37, 0:53]
    rgb = np.stack([(x * 5) % 256, (y * 7) % 256, (x * y) % 256], -1).astype(np.uint8)
    rgb[10:20] = rng.integers(0, 256, (10, 53, 3), dtype=np.uint8)
    got = vlb.gltf_texture(with_image(png(Image.fromarray(rgb), optimize=True), "rgb.gltf"), 0)
    assert np.array_equal(got["texels"][..., :3], rgb) and (got["texels"][..., 3] == 255).all()
    assert (got["wrap_u"], got["wrap_v"], got["filter"]) == (vlb.WRAP_REPEAT, vlb.WRAP_REPEAT, vlb.FILTER_LINEAR)   # no sampler: Application::Sampler{}
    grey = rgb[..., 0]
    got = vlb.gltf_texture(with_image(png(Image.fromarray(np.ascontiguousarray(grey))), "grey.gltf"), 0)
    assert np.array_equal(got["texels"][..., 0], grey) and np.array_equal(got["texels"][..., 2], grey)
    pal = Image.fromarray(rgb).quantize(colors=16)
    got = vlb.gltf_texture(with_image(png(pal), "pal.gltf"), 0)
    assert np.array_equal(got["texels"][..., :3], np.asarray(pal.convert("RGB")))
    rgba = np.concatenate([rgb, rng.integers(0, 256, (37, 53, 1), dtype=np.uint8)], -1)
    got = vlb.gltf_texture(with_image(png(Image.fromarray(rgba)), "rgba.gltf"), 0)
    assert np.array_equal(got["texels"], rgba)
    # unsupported / broken inputs fail loudly with the right code
    b = io.BytesIO(); Image.fromarray(rgb).save(b, format="BMP")
    with pytest.raises(vlb.VlbError) as e:
        vlb.gltf_texture(with_image(b.getvalue(), "bmp.gltf"), 0)
    assert e.value.code == vlb.ERR_UNSUPPORTED
    wide = Image.fromarray(np.ascontiguousarray((rgb.astype(np.uint16) * 257)[..., 0]))
    with pytest.raises(vlb.VlbError) as e:
        vlb.gltf_texture(with_image(png(wide), "p16.gltf"), 0)
    assert e.value.code == vlb.ERR_UNSUPPORTED
    good = png(Image.fromarray(rgb))
    with pytest.raises(vlb.VlbError) as e:
        vlb.gltf_texture(with_image(good[:len(good) // 2], "cut.gltf"), 0)
    assert e.value.code == vlb.ERR_IO
    # a texture nobody uses as baseColor is not decoded at all (placeholder, like the reference's dummy texture)
    d = dict(doc)
    d["images"] = [{"uri": "missing.png"}]
    d["textures"] = [{"source": 0}]
    d["materials"] = [dict(m) for m in doc["materials"]]
    d["materials"][0] = {"pbrMetallicRoughness": {"baseColorFactor": [1, 1, 1, 1]}, "normalTexture": {"index": 0}}
    json.dump(d, open(str(tmp_path / "unused.gltf"), "w"))
    got = vlb.gltf_texture(str(tmp_path / "unused.gltf"), 0)
    assert not got["used"] and got["texels"].shape == (1, 1, 4)


def test_jpeg_reader_vs_libjpeg(vlb, scenes, tmp_path):
    """csrc/jpeg_decode.cpp (baseline and progressive Huffman JPEG) against libjpeg through PIL. JPEG decoding is not bit-normative
    (IDCT / chroma upsampling arithmetic differ between decoders: "parity unpinned"), so the band is a few levels."""
    import io
    import json
    import base64
    from PIL import Image
    rng = np.random.default_rng(1)
    sc = scenes.small_room()
    sc["materials"]["textures"][0, 2, 0] = 0
    doc = json.load(open(scenes.write_gltf(sc, str(tmp_path / "base.gltf"))))

    def with_image(data, name):
        d = dict(doc)
        d["images"] = [{"uri": "data:image/jpeg;base64," + base64.b64encode(data).decode()}]
        d["textures"] = [{"source": 0}]
        p = str(tmp_path / name)
        json.dump(d, open(p, "w"))
        return p

    y, x = np.mgrid[0:97, 0:131]                                   # not a multiple of the MCU size
    rgb = np.stack([(x * 2) % 256, (y * 3) % 256, (x + y) % 256], -1).astype(np.uint8)
    rgb[20:50, 30:90] = rng.integers(0, 256, (30, 60, 3), dtype=np.uint8)
    cases = {"444": dict(subsampling=0), "422": dict(subsampling=1), "420": dict(subsampling=2), "q50": dict(quality=50),
             "huffopt": dict(optimize=True, quality=90), "restart": dict(quality=85, restart_marker_blocks=3),
             "prog444": dict(progressive=True, subsampling=0), "prog420": dict(progressive=True, subsampling=2, quality=80),
             "prog_restart": dict(progressive=True, optimize=True, restart_marker_blocks=4, subsampling=1)}
    for name, kw in cases.items():
        b = io.BytesIO()
        Image.fromarray(rgb).save(b, format="JPEG", **({"quality": 92} | kw))
        ref = np.asarray(Image.open(io.BytesIO(b.getvalue())).convert("RGB")).astype(int)
        got = vlb.gltf_texture(with_image(b.getvalue(), name + ".gltf"), 0)["texels"]
        assert got.shape == (97, 131, 4) and (got[..., 3] == 255).all()
        diff = np.abs(got[..., :3].astype(int) - ref)
        assert diff.max() <= 6 and diff.mean() <= 0.6, (name, diff.max(), diff.mean())
    g = io.BytesIO()
    Image.fromarray(np.ascontiguousarray(rgb[..., 0])).save(g, format="JPEG", quality=90)
    ref = np.asarray(Image.open(io.BytesIO(g.getvalue()))).astype(int)
    got = vlb.gltf_texture(with_image(g.getvalue(), "grey.gltf"), 0)["texels"]
    assert np.abs(got[..., 0].astype(int) - ref).max() <= 2 and np.array_equal(got[..., 0], got[..., 2])
    pg = io.BytesIO()
    Image.fromarray(np.ascontiguousarray(rgb[..., 1])).save(pg, format="JPEG", quality=90, progressive=True)
    ref = np.asarray(Image.open(io.BytesIO(pg.getvalue()))).astype(int)
    assert np.abs(vlb.gltf_texture(with_image(pg.getvalue(), "prog_grey.gltf"), 0)["texels"][..., 0].astype(int) - ref).max() <= 2
    cm = io.BytesIO()
    Image.fromarray(np.concatenate([rgb, rgb[..., :1]], -1), "CMYK").save(cm, format="JPEG")     # 4 components
    with pytest.raises(vlb.VlbError) as e:
        vlb.gltf_texture(with_image(cm.getvalue(), "cmyk.gltf"), 0)
    assert e.value.code == vlb.ERR_UNSUPPORTED
    with pytest.raises(vlb.VlbError) as e:
        vlb.gltf_texture(with_image(b.getvalue()[: len(b.getvalue()) // 2], "cut.gltf"), 0)
    assert e.value.code == vlb.ERR_IO


@pytest.mark.gpu
@pytest.mark.parametrize("container", ["gltf", "glb"])
def test_loaded_textured_scene_bakes_like_the_array_scene(ctx, vlb, scenes, tmp_path, container):
    sc = scenes.small_room_textured()
    p = (scenes.write_gltf if container == "gltf" else scenes.write_glb)(sc, str(tmp_path / ("room." + container)))
    s = _settings(vlb, scenes)
    ctx.set_scene(sc)
    ctx.build_bvh()
    a = ctx.bake_probes(s)
    ctx.load_gltf(p)
    ctx.build_bvh()
    b = ctx.bake_probes(s)
    assert rel_l2(b, a) <= 1e-5          # normals are re-normalised by the loader: last-ulp differences only


def test_image_file_loader_for_skyboxes(vlb, tmp_path):
    """vlb_image_load_rgba8 = the reference's stbi_load(file, ..., STBI_rgb_alpha) for skyboxes (skybox_manager.cpp:14-20)."""
    from PIL import Image
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (32, 64, 3), dtype=np.uint8)
    Image.fromarray(rgb).save(str(tmp_path / "sky.png"))
    px = vlb.image_load_rgba8(str(tmp_path / "sky.png"))
    assert px.shape == (32, 64, 4) and np.array_equal(px[..., :3], rgb) and (px[..., 3] == 255).all()
    smooth = np.zeros((32, 64, 3), np.uint8)
    smooth[..., 0] = np.linspace(0, 255, 64)[None, :]
    smooth[..., 1] = np.linspace(0, 255, 32)[:, None]
    Image.fromarray(smooth).save(str(tmp_path / "sky.jpg"), quality=95)
    pj = vlb.image_load_rgba8(str(tmp_path / "sky.jpg"))
    assert pj.shape == (32, 64, 4) and np.abs(pj[..., :3].astype(int) - smooth).max() <= 8
    with pytest.raises(vlb.VlbError) as e:
        vlb.image_load_rgba8(str(tmp_path / "missing.png"))
    assert e.value.code == vlb.ERR_IO
    open(str(tmp_path / "junk.png"), "wb").write(b"not an image")
    with pytest.raises(vlb.VlbError) as e:
        vlb.image_load_rgba8(str(tmp_path / "junk.png"))
    assert e.value.code == vlb.ERR_UNSUPPORTED


@pytest.mark.gpu
def test_cli_with_skybox_image_and_textures(ctx, vlb, scenes, tmp_path):
    """vlb_baker on a textured glTF with --skybox: equals the ABI calls with the same decoded inputs."""
    import os
    import subprocess
    from PIL import Image
    sc = scenes.small_room_textured()
    p = scenes.write_gltf(sc, str(tmp_path / "room.gltf"))
    sky8 = (np.clip(scenes.hdr_sky(64, 32, seed=2), 0, 1) * 255).astype(np.uint8)
    sky8[..., 3] = 255
    Image.fromarray(sky8).save(str(tmp_path / "sky.png"))
    exe = os.path.join(os.path.dirname(vlb.LIB_PATH), "vlb_baker")
    r = subprocess.run([exe, p, "--probes", "3x2x3", "--dirs", "32x16", "--light", "2,3.5,2", "--skybox", str(tmp_path / "sky.png")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    coeffs, _ = vlb.deserialize_gltf(str(tmp_path / "baked_room.gltf"))
    ctx.load_gltf(p)
    ctx.build_bvh()
    ctx.set_skybox(vlb.image_load_rgba8(str(tmp_path / "sky.png")))
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 3); s.dir_w, s.dir_h = 32, 16; s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags |= vlb.SKYBOX_ON_MISS
    vlb.settings_from_bounds(s, ctx.scene_bounds(tight=False))
    want = ctx.bake_probes(s)
    assert np.array_equal(np.asarray(coeffs).reshape(want.shape), want)


def test_radiance_hdr_loader(vlb, scenes, tmp_path):
    """vlb_image_load_rgba32f: Radiance RGBE -> linear RGBA32F (csrc/hdr_decode.cpp), flat and run-length coded
    scanlines; checked against the source within RGBE's 8-bit mantissa and, when OpenCV is there, exactly against it."""
    sky = scenes.hdr_sky(128, 64, seed=3)[..., :3]

    def to_rgbe(rgb):
        m = rgb.max(-1)
        e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0).astype(int)      # m < 2^e
        scale = np.where(m > 1e-32, 256.0 / np.exp2(e), 0.0)
        out = np.zeros(rgb.shape[:2] + (4,), np.uint8)
        out[..., :3] = np.clip(np.floor(rgb * scale[..., None]), 0, 255).astype(np.uint8)
        out[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
        return out

    rgbe = to_rgbe(sky)
    flat = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 64 +X 128\n" + rgbe.tobytes()
    open(str(tmp_path / "flat.hdr"), "wb").write(flat)
    px = vlb.image_load_rgba32f(str(tmp_path / "flat.hdr"))
    assert px.shape == (64, 128, 4) and (px[..., 3] == 1).all()
    want = rgbe[..., :3].astype(np.float32) * np.exp2(rgbe[..., 3:4].astype(np.float32) - 136)
    assert np.array_equal(px[..., :3], want)
    assert np.abs(px[..., :3] - sky).max() <= sky.max() / 128
    # new-style RLE: every scanline = 2 2 hi lo, then the 4 channels as runs / literals
    body = bytearray()
    for y in range(64):
        body += bytes([2, 2, 0, 128])
        for c in range(4):
            ch = rgbe[y, :, c]
            x = 0
            while x < 128:
                run = 1
                while x + run < 128 and run < 127 and ch[x + run] == ch[x]:
                    run += 1
                if run >= 3:
                    body += bytes([128 + run, int(ch[x])]); x += run
                else:
                    n = min(128 - x, 5)
                    body += bytes([n]) + ch[x:x + n].tobytes(); x += n
    open(str(tmp_path / "rle.hdr"), "wb").write(b"#?RADIANCE\n# made by the test\nFORMAT=32-bit_rle_rgbe\n\n-Y 64 +X 128\n" + bytes(body))
    assert np.array_equal(vlb.image_load_rgba32f(str(tmp_path / "rle.hdr")), px)
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None and cv2.imwrite(str(tmp_path / "cv.hdr"), np.ascontiguousarray(sky[..., ::-1])):
        ref = cv2.imread(str(tmp_path / "cv.hdr"), cv2.IMREAD_UNCHANGED)[..., ::-1]
        assert np.array_equal(vlb.image_load_rgba32f(str(tmp_path / "cv.hdr"))[..., :3], ref)
    open(str(tmp_path / "cut.hdr"), "wb").write(flat[: len(flat) // 2])
    with pytest.raises(vlb.VlbError) as e:
        vlb.image_load_rgba32f(str(tmp_path / "cut.hdr"))
    assert e.value.code == vlb.ERR_IO
    open(str(tmp_path / "xy.hdr"), "wb").write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n+Y 64 +X 128\n" + rgbe.tobytes())
    with pytest.raises(vlb.VlbError) as e:
        vlb.image_load_rgba32f(str(tmp_path / "xy.hdr"))
    assert e.value.code == vlb.ERR_UNSUPPORTED

"""vlb_baker — the reference's `baker` executable (src/baker/main.cpp:8-44) on top of the C ABI.
CPU: argument handling, error text / exit codes of main.cpp:13-16,26-43, --dry-run (host-only glTF parse).
GPU: `vlb_baker scene.gltf` writes baked_scene.gltf whose coefficient buffer is bit-identical to
vlb_bake_probes called with the same settings (LightBaker::bake + serialize, light_baker.cpp:287-402)."""
import os
import subprocess

import numpy as np
import pytest


def _cli(vlb):
    exe = os.path.join(os.path.dirname(vlb.LIB_PATH), "vlb_baker")
    if not os.path.exists(exe):
        import importlib
        importlib.import_module("vulkan-light-bakery_b200.build").build()
    return exe


def _run(vlb, *args, cwd=None):
    return subprocess.run([_cli(vlb)] + list(args), capture_output=True, text=True, cwd=cwd, timeout=300)


def test_no_argument_is_the_reference_error(vlb):
    r = _run(vlb)
    assert r.returncode == 1
    assert "std::exception: Select scene to bake." in r.stderr          # main.cpp:15,33


def test_missing_file_and_bad_options(vlb, tmp_path):
    r = _run(vlb, str(tmp_path / "nope.gltf"), "--dry-run")
    assert r.returncode == 1 and "std::exception:" in r.stderr
    r = _run(vlb, "a.gltf", "--probes", "7x7")
    assert r.returncode == 1 and "--probes" in r.stderr
    r = _run(vlb, "sky.png")
    assert r.returncode == 1 and "image input" in r.stderr
    r = _run(vlb, "a.gltf", "--builder", "sah")
    assert r.returncode == 1 and "--builder" in r.stderr


def test_dry_run_reports_reference_defaults(vlb, scenes, tmp_path):
    p = scenes.write_gltf(scenes.small_room(), str(tmp_path / "room.gltf"))
    r = _run(vlb, p, "--dry-run")
    assert r.returncode == 0, r.stderr
    assert "7x7x7 probes x 3141x1000 rays, order 3" in r.stdout           # light_baker.cpp:38,65,294
    assert r.stdout.strip().endswith("exiting...")                          # main.cpp:42
    assert str(tmp_path / "baked_room.gltf") in r.stdout                    # light_baker.cpp:399
    r = _run(vlb, "room.gltf", "--dry-run", cwd=str(tmp_path))
    assert "-> baked_room.gltf" in r.stdout


@pytest.mark.gpu
def test_cli_bake_equals_abi_bake(ctx, vlb, scenes, tmp_path):
    p = scenes.write_gltf(scenes.small_room(), str(tmp_path / "room.gltf"))
    r = _run(vlb, p, "--probes", "3x2x3", "--dirs", "64x32", "--light", "2,3.5,2")
    assert r.returncode == 0, r.stderr + r.stdout
    out = str(tmp_path / "baked_room.gltf")
    assert os.path.exists(out)
    coeffs, step = vlb.deserialize_gltf(out)
    ctx.load_gltf(p)
    ctx.build_bvh()
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 3)
    s.dir_w, s.dir_h = 64, 32
    s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags &= ~vlb.SKYBOX_ON_MISS                 # the CLI has no skybox input (reference bake: App. B-5)
    vlb.settings_from_bounds(s, ctx.scene_bounds(tight=False))
    want = ctx.bake_probes(s)
    assert np.array_equal(np.asarray(coeffs).reshape(want.shape), want)
    assert np.allclose(step, list(s.step))


@pytest.mark.gpu
def test_cli_builder_option_never_changes_the_file(vlb, scenes, tmp_path):
    """--builder lbvh / ploc / auto: another tree, the same coefficients bit for bit."""
    p = scenes.write_gltf(scenes.atrium(8192, seed=3), str(tmp_path / "atrium.gltf"), index_dtype=np.uint32)
    args = ["--tight-bounds", "--probes", "4x3x4", "--dirs", "64x32", "--light", "15,11,9"]
    outs = []
    for b in ("lbvh", "ploc", "auto"):
        o = str(tmp_path / ("%s.gltf" % b))
        r = _run(vlb, p, *args, "--builder", b, "--out", o)
        assert r.returncode == 0, r.stderr + r.stdout
        assert ("BVH (PLOC)" in r.stdout) == (b == "ploc")         # auto: 98 k rays through 8 k triangles -> LBVH
        outs.append(vlb.deserialize_gltf(o)[0])
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.gpu
def test_cli_devices_equals_single_device(vlb, scenes, tmp_path):
    """`--devices 0,0`: two contexts in one process (vlb_bake_probes_multi) write the same file as one."""
    p = scenes.write_gltf(scenes.small_room(), str(tmp_path / "room.gltf"))
    args = ["--probes", "3x2x4", "--dirs", "32x16", "--light", "2,3.5,2"]
    r1 = _run(vlb, p, *args, "--out", str(tmp_path / "one.gltf"))
    r2 = _run(vlb, p, *args, "--devices", "0,0", "--out", str(tmp_path / "two.gltf"))
    assert r1.returncode == 0 and r2.returncode == 0, r1.stderr + r2.stderr
    a, _ = vlb.deserialize_gltf(str(tmp_path / "one.gltf"))
    b, _ = vlb.deserialize_gltf(str(tmp_path / "two.gltf"))
    assert np.array_equal(a, b) and "2 GPU(s)" in r2.stdout

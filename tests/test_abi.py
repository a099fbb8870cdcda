"""The C-ABI library loads and exports every symbol include/vlb_bake.h declares; host-only helpers
behave; and without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "vlb_bake.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(vlb):
    lib = ctypes.CDLL(vlb.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libvlb_bake.so does not export %s" % name
    assert sorted(vlb.ABI_SYMBOLS) == declared
    assert vlb.load_library().vlb_abi_version() == 1


def test_struct_layouts_match_reference(vlb):
    # shaders/structures.h:13-71 scalar layout
    assert vlb.VERTEX_DTYPE.itemsize == 44 and vlb.MATERIAL_DTYPE.itemsize == 144
    assert vlb.VERTEX_DTYPE.fields["normal"][1] == 16 and vlb.VERTEX_DTYPE.fields["uv0"][1] == 28
    assert vlb.MATERIAL_DTYPE.fields["base_color_factor"][1] == 64
    assert ctypes.sizeof(vlb.BakeSettings) == 4 * (3 + 3 + 3 + 3 + 3 + 5 + 2 + 1 + 2 + 4)


def test_default_settings_are_the_reference_constants(vlb):
    s = vlb.default_settings()
    assert list(s.probes) == [7, 7, 7]                      # light_baker.cpp:38
    assert (s.dir_w, s.dir_h) == (3141, 1000)               # light_baker.cpp:65 / env_map_generator.cpp:24-28
    assert int(np.float32(3.1415926538) * np.float32(500) * np.float32(2)) == 3141
    assert s.sh_order == 3
    assert list(s.light_pos) == [1.0, 10.0, 1.0]            # env_map.rchit:25
    assert (s.shadow_bias, s.c_diffuse, s.c_specular, s.gloss, s.ambient) == (np.float32(0.005), 0.5, 0.5, 16.0, 0.0)
    assert (s.tmin, s.tmax) == (np.float32(0.001), 10000.0)  # env_map.rgen:22-23
    assert s.flags == vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE | vlb.QUANTIZE_RGBA8
    assert s.slab == (0, 7)


def test_no_gpu_means_loud_failure(vlb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(vlb.VlbError) as e:
        vlb.Context(0)
    assert e.value.code == vlb.ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_comm_host_side(vlb):
    """The multi-GPU section of the ABI without a GPU: the NCCL id is 128 bytes (run-time libnccl), bad arguments are
    refused, and a communicator cannot be made without a context."""
    lib = vlb.load_library()
    assert vlb.COMM_ID_BYTES == 128
    small = ctypes.create_string_buffer(16)
    assert lib.vlb_comm_get_unique_id(ctypes.cast(small, ctypes.c_void_p), 16) == vlb.ERR_INVALID or \
        lib.vlb_comm_get_unique_id(ctypes.cast(small, ctypes.c_void_p), 16) == vlb.ERR_UNSUPPORTED
    try:
        a, b = vlb.comm_unique_id(), vlb.comm_unique_id()
    except vlb.VlbError as e:                       # a machine without libnccl.so.2: loud, not silent
        assert e.code == vlb.ERR_UNSUPPORTED and "NCCL" in str(e)
        return
    assert len(a) == 128 and len(b) == 128 and a != b
    assert lib.vlb_comm_init_rank(None, a, 0, 1) == vlb.ERR_INVALID
    assert lib.vlb_comm_destroy(None) == vlb.ERR_INVALID
    assert lib.vlb_comm_sharded_uploads(None, 1) == vlb.ERR_INVALID
    assert lib.vlb_bake_probes_sharded(None, None, None) == vlb.ERR_INVALID


def test_builder_recommendation(vlb):
    """vlb_bvh_recommend_builder (host-only): PLOC when the trace is long enough to pay for its build."""
    assert vlb.recommend_builder(262144, 64 * 32 * 64 * 4096) == "ploc"            # C3 on one GPU
    assert vlb.recommend_builder(262144, 64 * 32 * 64 * 4096 // 8) == "ploc"       # C3, one of 8 ranks
    assert vlb.recommend_builder(262144, 16 * 8 * 16 * 1024) == "lbvh"             # C2: a 0.9 ms bake
    assert vlb.recommend_builder(3 << 20, 32 * 16 * 32 * 4096 * 4) == "lbvh"       # C4: no trace gain measured at 3 M triangles
    assert vlb.recommend_builder(12, 343 * 3141 * 1000) == "lbvh"                  # the default cube
    assert vlb.recommend_builder(0, 10 ** 12) == "lbvh"


def test_slab_partition(vlb):
    import importlib
    par = importlib.import_module("vulkan-light-bakery_b200.parallel")
    for nz in (1, 7, 16, 64, 65):
        for world in (1, 2, 3, 4, 8):
            ranges = [par.slab_range(nz, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/vlb_bake.h must compile as C99 (no C++ / CUDA / torch types) and a C caller
    must link against the library."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "vlb_bake.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "caller.c"
    src.write_text('#include "vlb_bake.h"\n#include <stdio.h>\n'
                   'int main(void) { vlb_bake_settings s; vlb_bake_settings_default(&s);\n'
                   '  printf("%d %d %d %d\\n", vlb_abi_version(), s.probes[0], s.dir_w, (int)sizeof(vlb_material)); return 0; }\n')
    lib_dir = os.path.join(ROOT, "vulkan-light-bakery_b200")
    exe = tmp_path / "caller"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib_dir,
                           "-lvlb_bake", "-Wl,-rpath," + lib_dir])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert out == ["1", "7", "3141", "144"]


def test_tile_reciprocal_is_exact():
    """bake.cu divides a tile index by tiles_x with umul64hi(tile, rcp), rcp = floor(2^64 / tiles_x) + 1 as bake_device computes
    it in 64-bit arithmetic. Exact for every 32-bit tile and every tiles_x >= 2 (tiles_x == 1 takes a branch)."""
    rng = np.random.default_rng(5)
    M = (1 << 64) - 1
    ds = [2, 3, 5, 7, 16, 20, 393, 1024, 4097, 65535, 65536, (1 << 31) - 1, 1 << 31, (1 << 32) - 1] + [int(x) for x in rng.integers(2, 1 << 32, 200)]
    for d in ds:
        rcp = M // d + (2 if (M % d) + 1 == d else 1)          # the host expression of bake_device
        assert rcp == (1 << 64) // d + 1 and rcp <= M
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 32) - 1, ((1 << 32) - 1) // d * d, ((1 << 32) - 1) // d * d - 1] + [int(x) for x in rng.integers(0, 1 << 32, 50)]
        for n in ns:
            if 0 <= n < (1 << 32):
                assert (n * rcp) >> 64 == n // d, (n, d)

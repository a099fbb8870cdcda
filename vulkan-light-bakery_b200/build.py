"""Build recipe of libvlb_bake.so (hand-written sm_100a CUDA behind the C ABI of include/vlb_bake.h).

Built IN-TREE (vulkan-light-bakery_b200/libvlb_bake.so) so that the binary travels to the GPU box
with the repository snapshot. nvcc cross-compiles for sm_100a without a GPU present.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# VLB_BUILD_TAG=x builds a variant library libvlb_bake_x.so (objects in build_x/) next to the product
# one, typically with VLB_NVCC_EXTRA=-D...; python loads it with VLB_LIB=<path> for A/B runs.
TAG = os.environ.get("VLB_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "libvlb_bake%s.so" % ("_" + TAG if TAG else ""))
CLI = os.path.join(HERE, "vlb_baker")          # the reference's `baker` executable on top of the C ABI
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SOURCES = ["context.cu", "bvh_build.cu", "bake.cu", "bake_gather.cu", "diag.cu", "skybox_sh.cu", "comm.cu", "host_tables.cpp", "gltf_io.cpp", "gltf_scene.cpp", "png_decode.cpp", "jpeg_decode.cpp", "hdr_decode.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr"]
NVCC_FLAGS += os.environ.get("VLB_NVCC_EXTRA", "").split()   # e.g. -DVLB_PROJ_TIMING for instrumented debug builds


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(CSRC, "bake.cu"))      # bake_gather.cu includes it
    hdrs.append(os.path.join(HERE, "..", "include", "vlb_bake.h"))
    return hdrs


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _deps()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs, jobs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=4) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lz", "-lpthread", "-ldl"]
        subprocess.check_call(cmd)
    cli_src = os.path.join(CSRC, "vlb_baker_main.cpp")
    if not TAG and (force or _stale(CLI, [cli_src, LIB] + hdrs)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(HERE, "..", "include"), cli_src, "-o", CLI,
                               "-L", HERE, "-lvlb_bake", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

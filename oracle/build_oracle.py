"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE — see oracle/vlb_oracle.cpp header).

  build_oracle()  g++ oracle/vlb_oracle.cpp            -> oracle/libvlb_oracle.so
  build_ref()     g++ oracle/ref_shim.cpp + the reference's own shaders/sh_common.h, compiled
                  from where it lies under /root/reference -> oracle/_ref/libvlb_refsh.so, and
                  oracle/make_ref_shaders.py: the reference's shader sources (env_map.rgen/.rchit, main.rmiss,
                  shadow.rmiss, sh.comp, skybox_sh.comp) behind oracle/glsl_shim.h -> oracle/_ref/libvlb_refshaders.so
                  (only when /root/reference exists; the GPU box uses the prebuilt files).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("VLB_REFERENCE_ROOT", "/root/reference")
ORACLE_SO = os.path.join(HERE, "libvlb_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libvlb_refsh.so")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False):
    srcs = [os.path.join(HERE, "vlb_oracle.cpp"), os.path.join(HERE, "..", "include", "vlb_bake.h")]
    if not force and _newer(ORACLE_SO, srcs):
        return ORACLE_SO
    cmd = ["g++", "-std=c++17", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-shared",
           "-fPIC", "-o", ORACLE_SO, srcs[0]]
    subprocess.check_call(cmd)
    return ORACLE_SO


def build_ref(force=False):
    try:                                   # works as `oracle.build_oracle` and as a script
        from . import make_ref_shaders
    except ImportError:
        import make_ref_shaders
    make_ref_shaders.build(force)          # oracle/_ref/libvlb_refshaders.so: the reference's shaders behind oracle/glsl_shim.h
    header = os.path.join(REFERENCE, "shaders", "sh_common.h")
    if not os.path.exists(header):
        return REF_SO if os.path.exists(REF_SO) else None
    srcs = [os.path.join(HERE, "ref_shim.cpp"), header]
    if not force and _newer(REF_SO, srcs):
        return REF_SO
    os.makedirs(REF_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-fsingle-precision-constant", "-ffp-contract=off", "-shared", "-fPIC",
           "-I", os.path.join(REFERENCE, "shaders"), "-o", REF_SO, srcs[0]]
    subprocess.check_call(cmd)
    return REF_SO


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))

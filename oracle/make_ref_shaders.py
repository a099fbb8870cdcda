"""Compiles the REFERENCE's own shader sources as C++ (TEST INFRASTRUCTURE; outputs only into oracle/_ref/).

For each shader of the bake / skybox path -- and the viewer's main.rchit + sh.rmiss, whose gather the multi-bounce passes
iterate -- the file is read from where it lies under /root/reference/shaders and
only its INTERFACE syntax is touched:
  * `#version` / `#extension` lines are dropped,
  * `layout(...) ... ;` declarations and `hitAttributeEXT ... ;` are dropped; the first one is replaced by
    `#include GLUE_DECLS` (the C++ stand-ins for those resources, oracle/ref_glue/<shader>_decls.inc),
  * `#include "structures.h"` gets `using namespace shader;` appended (the header's C++ side wraps its structs in
    a namespace),
  * `main` is renamed `shader_main`,
  * a GLSL array constructor `T[](a, b, ...)` becomes the C++ aggregate `{a, b, ...}` (main.rchit's gridVertices).
Every function BODY -- sRGB, getBaseColor, dir2SkyboxUV, all main()s, sh_common.h, structures.h -- is compiled from
the reference's text, unmodified, behind oracle/glsl_shim.h. The filtered text only ever exists in a temporary
directory; no reference source is written into the repository.

    python oracle/make_ref_shaders.py          ->  oracle/_ref/libvlb_refshaders.so
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("VLB_REFERENCE_ROOT", "/root/reference")
SHADER_DIR = os.path.join(REFERENCE, "shaders")
REF_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(REF_DIR, "libvlb_refshaders.so")
GLUE = os.path.join(HERE, "ref_glue")

# shader file -> glue translation unit
SHADERS = {
    "env_map.rgen": "env_map_rgen.cpp",
    "env_map.rchit": "env_map_rchit.cpp",
    "main.rmiss": "main_rmiss.cpp",
    "shadow.rmiss": "shadow_rmiss.cpp",
    "sh.comp": "sh_comp.cpp",
    "skybox_sh.comp": "skybox_sh_comp.cpp",
    # the viewer's gather operator, which the multi-bounce passes iterate (main.rchit:124-167 + sh.rmiss:20-36)
    "main.rchit": "main_rchit.cpp",
    "sh.rmiss": "sh_rmiss.cpp",
}

_DECL = re.compile(r"^[ \t]*(?:layout\s*\([^)]*\)|hitAttributeEXT)[^;{]*(?:\{[^}]*\})?[^;]*;[ \t]*\n", re.M | re.S)


def filter_shader(text):
    text = re.sub(r"^[ \t]*#(?:version|extension)[^\n]*\n", "", text, flags=re.M)
    first = [True]

    def repl(_m):
        if first[0]:
            first[0] = False
            return "#include GLUE_DECLS\n"
        return ""

    text = _DECL.sub(repl, text)
    if first[0]:
        raise RuntimeError("no interface declaration found")
    text = re.sub(r'(#include\s+"structures\.h"[^\n]*\n)', r"\1using namespace shader;\n", text)
    # GLSL array constructor `T name[N] = T[]( ... );` (main.rchit:128-137) -> C++ aggregate `T name[N] = { ... };`: syntax only
    text = re.sub(r"=\s*(\w+)\[\]\s*\(((?:[^()]|\([^()]*\))*)\)\s*;", r"= {\2};", text, flags=re.S)
    text, n = re.subn(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", text)
    if n != 1:
        raise RuntimeError("expected exactly one main()")
    return text


def available():
    return all(os.path.exists(os.path.join(SHADER_DIR, s)) for s in SHADERS)


def sources():
    return ([os.path.join(SHADER_DIR, s) for s in SHADERS] + [os.path.join(SHADER_DIR, "sh_common.h"), os.path.join(SHADER_DIR, "structures.h")] +
            [os.path.join(GLUE, f) for f in sorted(os.listdir(GLUE)) if os.path.isfile(os.path.join(GLUE, f))] +
            [os.path.join(HERE, "glsl_shim.h"), os.path.join(HERE, "ref_pipeline.cpp"), os.path.abspath(__file__)])


def build(force=False):
    if not available():
        return OUT_SO if os.path.exists(OUT_SO) else None
    if not force and os.path.exists(OUT_SO) and all(os.path.getmtime(s) <= os.path.getmtime(OUT_SO) for s in sources()):
        return OUT_SO
    os.makedirs(REF_DIR, exist_ok=True)
    flags = ["-std=c++17", "-O2", "-fsingle-precision-constant", "-ffp-contract=off", "-fPIC", "-w"]
    flags += os.environ.get("VLB_REF_CFLAGS", "").split()          # e.g. "-g -fsanitize=address" when debugging the glue
    with tempfile.TemporaryDirectory(prefix="vlb_refshaders_") as tmp:
        objs = []
        for shader, glue in SHADERS.items():
            with open(os.path.join(SHADER_DIR, shader)) as f:
                filtered = filter_shader(f.read())
            with open(os.path.join(tmp, shader + ".inc"), "w") as f:
                f.write(filtered)
            obj = os.path.join(tmp, glue + ".o")
            # include order: the filtered text (tmp), the glue + glm stand-in, then the reference's own headers
            subprocess.check_call(["g++"] + flags + ["-I", tmp, "-I", GLUE, "-I", HERE, "-I", SHADER_DIR, "-c", os.path.join(GLUE, glue), "-o", obj])
            objs.append(obj)
        obj = os.path.join(tmp, "ref_pipeline.o")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-I", HERE, "-I", os.path.join(HERE, "..", "include"),
                               "-c", os.path.join(HERE, "ref_pipeline.cpp"), "-o", obj])
        objs.append(obj)
        subprocess.check_call(["g++", "-shared", "-o", OUT_SO] + objs + os.environ.get("VLB_REF_LDFLAGS", "").split())
    return OUT_SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

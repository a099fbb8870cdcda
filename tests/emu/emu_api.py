"""ctypes binding of tests/emu/libvlb_emu.so — TEST-ONLY host emulation of the device code
(see tests/emu/emu_driver.cpp). Lets the GPU-less `-m "not gpu"` suite exercise the LBVH build,
traversal and shading logic that the CUDA kernels execute."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
SO = os.path.join(_HERE, "libvlb_emu.so")
_vp, _u64, _i32, _f32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_float
_lib = None


def build(force=False, so=None, defines=()):
    """Builds the emulation library; `defines` (e.g. ["-DVLB_NODE_Q8=1"]) + `so` build a variant beside it."""
    so = so or SO
    srcs = [os.path.join(_HERE, "emu_driver.cpp"),
            os.path.join(_ROOT, "vulkan-light-bakery_b200", "csrc", "host_tables.cpp")]
    csrc = os.path.join(_ROOT, "vulkan-light-bakery_b200", "csrc")
    deps = srcs + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not force and os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", cuda_inc,
                           "-o", so] + list(defines) + srcs)
    return so


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.emu_scene_create.restype = _vp
        L.emu_scene_create.argtypes = [_vp, _vp, _vp, ctypes.c_uint32, _vp, ctypes.c_uint32, _i32]
        L.emu_scene_destroy.argtypes = [_vp]
        L.emu_set_builder.argtypes = [_i32, _i32]
        L.emu_scene_max_depth.restype = _i32
        L.emu_scene_max_depth.argtypes = [_vp]
        L.emu_scene_set_skybox.argtypes = [_vp, _vp, _i32, _i32]
        L.emu_scene_node_stats.argtypes = [_vp, _vp]
        L.emu_scene_set_textures.argtypes = [_vp, _vp, ctypes.c_uint32]
        L.emu_tex_sample.argtypes = [_vp, _i32, _vp, _u64, _vp]
        L.emu_trace_rays.argtypes = [_vp, _vp, _vp, _u64, _f32, _f32, _i32, _vp, _vp, _vp]
        L.emu_bake.argtypes = [_vp, _vp, _vp]
        L.emu_bake_gather.argtypes = [_vp, _vp, _vp, _vp]
        _lib = L
    return _lib


class Scene:
    def __init__(self, scene, max_leaf=4, builder="lbvh", ploc_radius=16):
        lib().emu_set_builder({"lbvh": 0, "ploc": 1}[builder], ploc_radius)
        self._keep = [np.ascontiguousarray(scene[k]) for k in ("vertices", "indices", "instances", "materials")]
        v, i, inst, m = self._keep
        self._h = lib().emu_scene_create(_p(v), _p(i), _p(inst), inst.size, _p(m), m.size, max_leaf)
        if scene.get("textures"):
            self.set_textures(scene["textures"])

    def set_textures(self, textures):
        import importlib
        arr, keep = importlib.import_module("vulkan-light-bakery_b200").pack_textures(textures)
        lib().emu_scene_set_textures(self._h, ctypes.cast(arr, _vp), len(textures))

    def tex_sample(self, tex, uv):
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), np.float32)
        lib().emu_tex_sample(self._h, int(tex), _p(uv), uv.shape[0], _p(out))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().emu_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def max_depth(self):
        return int(lib().emu_scene_max_depth(self._h))

    def node_stats(self):
        """(wide nodes, child slots in use, leaf children) of the emitted 4-wide tree."""
        out = np.zeros(3, np.uint64)
        lib().emu_scene_node_stats(self._h, _p(out))
        return tuple(int(x) for x in out)

    def set_skybox(self, rgba32f):
        t = np.ascontiguousarray(rgba32f, np.float32)
        lib().emu_scene_set_skybox(self._h, _p(t), t.shape[1], t.shape[0])

    def trace_rays(self, origins, dirs, tmin=0.001, tmax=10000.0, kind=0):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        ids = np.zeros(o.shape[0], np.int32)
        tuv = np.zeros((o.shape[0], 3), np.float32)
        cnt = np.zeros(2, np.uint64)
        lib().emu_trace_rays(self._h, _p(o), _p(d), o.shape[0], tmin, tmax, kind, _p(ids), _p(tuv), _p(cnt))
        return ids, tuv, cnt

    def bake(self, settings, prev_full=None):
        k1 = settings.probes[2] if settings.slab_k1 < 0 else settings.slab_k1
        k0 = 0 if settings.slab_k1 < 0 else settings.slab_k0
        n = settings.probes[0] * settings.probes[1] * (k1 - k0)
        out = np.zeros((n, 16, 3), np.float32)
        prev = None if prev_full is None else np.ascontiguousarray(prev_full, np.float32).reshape(-1, 48)
        lib().emu_bake_gather(self._h, ctypes.byref(settings), _p(prev), _p(out))
        return out

"""ctypes binding of the CPU oracle (oracle/libvlb_oracle.so) and of the reference-header shim
(oracle/_ref/libvlb_refsh.so). TEST INFRASTRUCTURE: imported only by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / `--impl reference` legs."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libvlb_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libvlb_refsh.so")
REF_SHADERS_SO = os.path.join(_HERE, "_ref", "libvlb_refshaders.so")

_vp, _u64, _i32, _f32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_float
_lib = None
_ref = None


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            from . import build_oracle  # noqa: F401  (package-relative when imported as oracle.oracle_api)
            build_oracle.build_oracle()
        L = ctypes.CDLL(ORACLE_SO)
        L.vo_num_threads.restype = _i32
        L.vo_set_num_threads.argtypes = [_i32]
        L.vo_sh_basis.argtypes = [_vp, _u64, _vp]
        L.vo_x2phi.restype = _f32
        L.vo_x2phi.argtypes = [_i32, _i32]
        L.vo_y2theta.restype = _f32
        L.vo_y2theta.argtypes = [_i32, _i32]
        L.vo_to_vector.argtypes = [_f32, _f32, _vp]
        L.vo_srgb.restype = _f32
        L.vo_srgb.argtypes = [_f32]
        L.vo_dir2uv.argtypes = [_vp, _vp]
        L.vo_sky_lookup.argtypes = [_vp, _vp, _vp]
        L.vo_base_color.argtypes = [_vp, ctypes.c_uint32, _f32, _f32, _vp]
        L.vo_probe_dirs.argtypes = [_i32, _i32, _vp, _vp, _vp]
        L.vo_skybox_project.argtypes = [_vp, _i32, _i32, _i32, _i32, _vp]
        L.vo_envmap_project.argtypes = [_vp, _i32, _i32, _i32, _i32, _vp]
        L.vo_sh_reconstruct.argtypes = [_vp, _i32, _i32, _i32, _vp]
        L.vo_probe_positions.argtypes = [_vp, _vp]
        L.vo_probe_positions_literal.restype = _u64
        L.vo_probe_positions_literal.argtypes = [_vp, _vp, _vp, _vp]
        L.vo_scene_create.restype = _vp
        L.vo_scene_create.argtypes = [_vp, _u64, _vp, _u64, _vp, ctypes.c_uint32, _vp, ctypes.c_uint32]
        L.vo_scene_destroy.argtypes = [_vp]
        L.vo_scene_num_triangles.restype = _u64
        L.vo_scene_num_triangles.argtypes = [_vp]
        L.vo_scene_bounds.argtypes = [_vp, _i32, _vp]
        L.vo_scene_set_skybox.argtypes = [_vp, _vp, _i32, _i32, _i32]
        L.vo_scene_set_textures.argtypes = [_vp, _vp, ctypes.c_uint32]
        L.vo_tex_sample.argtypes = [_vp, ctypes.c_uint32, _vp, _u64, _vp]
        L.vo_scene_triangles.argtypes = [_vp, _vp]
        L.vo_trace_rays.argtypes = [_vp, _vp, _vp, _u64, _f32, _f32, _i32, _i32, _vp, _vp]
        L.vo_bake_probes.restype = _u64
        L.vo_bake_probes.argtypes = [_vp, _vp, _vp, _u64, _i32, _vp]
        L.vo_bake_gather.restype = _u64
        L.vo_bake_gather.argtypes = [_vp, _vp, _vp, _vp, _u64, _i32, _vp]
        L.vo_probe_envmap.argtypes = [_vp, _vp, _vp, _i32, _vp, _vp]
        L.vo_probe_envmap_gather.argtypes = [_vp, _vp, _vp, _vp, _i32, _vp, _vp]
        _lib = L
    return _lib


def ref_lib():
    """The reference's own sh_common.h compiled by oracle/build_oracle.py; None if unavailable."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        R = ctypes.CDLL(REF_SO)
        R.ref_SH.restype = _f32
        R.ref_SH.argtypes = [_i32, _i32, _f32, _f32, _f32]
        R.ref_x2phi.restype = _f32
        R.ref_x2phi.argtypes = [_i32, _i32]
        R.ref_y2theta.restype = _f32
        R.ref_y2theta.argtypes = [_i32, _i32]
        R.ref_toVector.argtypes = [_f32, _f32, _vp]
        R.ref_PI.restype = _f32
        _ref = R
    return _ref


_refsh = None


def ref_shaders_lib():
    """The reference's own shaders (env_map.rgen/.rchit, main.rmiss, shadow.rmiss, sh.comp, skybox_sh.comp, main.rchit, sh.rmiss) compiled
    as C++ by oracle/make_ref_shaders.py into oracle/_ref/libvlb_refshaders.so; None if unavailable."""
    global _refsh
    if _refsh is None and os.path.exists(REF_SHADERS_SO):
        R = ctypes.CDLL(REF_SHADERS_SO)
        for name in ("ref_rchit_srgb", "ref_rmiss_srgb", "ref_rmiss_dir2uv"):
            getattr(R, name).argtypes = [_vp, _vp]
        R.ref_rchit_get_base_color.argtypes = [_vp, _vp, _vp, _vp]
        R.ref_rchit_get_light.argtypes = [_vp]
        R.ref_rmiss_run.argtypes = [_vp, _vp, _i32, _i32, _vp]
        R.ref_sh_comp_dispatch.argtypes = [_vp, _i32, _i32, _vp]
        R.ref_skybox_sh_comp_dispatch.argtypes = [_vp, _i32, _i32, _vp]
        R.rp_create.restype = _vp
        R.rp_create.argtypes = [_vp, _vp, _vp, ctypes.c_uint32, _vp, _vp, _vp]
        R.rp_destroy.argtypes = [_vp]
        R.rp_set_skybox.argtypes = [_vp, _vp, _i32, _i32]
        R.rp_set_textures.argtypes = [_vp, _vp, _vp, ctypes.c_uint32]
        R.rp_bake_probe.restype = _u64
        R.rp_bake_probe.argtypes = [_vp, _vp, _i32, _i32, ctypes.c_uint32, _vp, _vp, _vp]
        if hasattr(R, "rp_bake_probe_viewer_hit"):
            f32 = ctypes.c_float
            R.rp_bake_probe_viewer_hit.restype = _u64
            R.rp_bake_probe_viewer_hit.argtypes = [_vp, _vp, _i32, _i32, ctypes.c_uint32, _vp, _vp, ctypes.c_uint32, f32, f32, f32, f32, f32,
                                                   _vp, _vp, _vp]
        _refsh = R
    return _refsh


class RefShaders:
    """Thin numpy front end of oracle/_ref/libvlb_refshaders.so: the reference's shader functions, one call each."""

    def __init__(self):
        self.R = ref_shaders_lib()
        if self.R is None:
            raise RuntimeError("oracle/_ref/libvlb_refshaders.so not built (needs /root/reference: python oracle/make_ref_shaders.py)")

    def _map4(self, fn, rgba):
        a = np.ascontiguousarray(rgba, np.float32).reshape(-1, 4)
        out = np.zeros_like(a)
        for i in range(a.shape[0]):
            fn(_p(a[i:i + 1]), _p(out[i:i + 1]))
        return out

    def srgb_rchit(self, rgba):            # shaders/env_map.rchit:27-34
        return self._map4(self.R.ref_rchit_srgb, rgba)

    def srgb_rmiss(self, rgba):            # shaders/main.rmiss:9-16
        return self._map4(self.R.ref_rmiss_srgb, rgba)

    def dir2uv(self, dirs):                # shaders/main.rmiss:18-35
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        out = np.zeros((d.shape[0], 2), np.float32)
        for i in range(d.shape[0]):
            self.R.ref_rmiss_dir2uv(_p(d[i:i + 1]), _p(out[i:i + 1]))
        return out

    def miss(self, dirs, sky):             # shaders/main.rmiss:37-41 (texture lookup: oracle/glsl_shim.h)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(sky, np.float32)
        out = np.zeros((d.shape[0], 3), np.float32)
        for i in range(d.shape[0]):
            self.R.ref_rmiss_run(_p(d[i:i + 1]), _p(t), t.shape[1], t.shape[0], _p(out[i:i + 1]))
        return out

    def base_color(self, material, uv, textures_f32=()):   # shaders/env_map.rchit:36-49
        """material: one element of a MATERIAL_DTYPE array; textures_f32: list of float32 [H, W, 4] arrays."""
        m = np.ascontiguousarray(material).reshape(1)
        tex = [np.ascontiguousarray(t, np.float32) for t in textures_f32]
        samplers = np.zeros(max(len(tex), 1), np.dtype([("texels", "<u8"), ("w", "<i4"), ("h", "<i4")]))
        for i, t in enumerate(tex):
            samplers[i] = (t.ctypes.data, t.shape[1], t.shape[0])
        u = np.ascontiguousarray(uv, np.float32).reshape(2)
        out = np.zeros(4, np.float32)
        self.R.ref_rchit_get_base_color(_p(m), _p(u), _p(samplers), _p(out))
        return out

    def default_light(self):               # shaders/env_map.rchit:25
        out = np.zeros(3, np.float32)
        self.R.ref_rchit_get_light(_p(out))
        return out

    def skybox_sh(self, texels_f32):       # the whole dispatch of shaders/skybox_sh.comp, coeffs in double
        t = np.ascontiguousarray(texels_f32, np.float32)
        out = np.zeros((16, 3), np.float64)
        self.R.ref_skybox_sh_comp_dispatch(_p(t), t.shape[1], t.shape[0], _p(out))
        return out

    def envmap_sh(self, texels_f32):       # the whole dispatch of shaders/sh.comp, coeffs in double
        t = np.array(texels_f32, np.float32, order="C")
        out = np.zeros((16, 3), np.float64)
        self.R.ref_sh_comp_dispatch(_p(t), t.shape[1], t.shape[0], _p(out))
        return out


class RefPipeline:
    """The reference's bake pipeline (env_map.rgen -> env_map.rchit / main.rmiss / shadow.rmiss -> sh.comp) run from
    its own shader code, one probe at a time; ray / triangle intersection is the oracle scene's (oracle/ref_pipeline.cpp)."""

    def __init__(self, scene, oracle_scene, sky=None, textures_f32=()):
        self.R = ref_shaders_lib()
        if self.R is None:
            raise RuntimeError("oracle/_ref/libvlb_refshaders.so not built")
        self._keep = [np.ascontiguousarray(scene[k]) for k in ("vertices", "indices", "instances", "materials")]
        v, i, inst, m = self._keep
        self._osc = oracle_scene
        trace = ctypes.cast(lib().vo_trace_rays, _vp)
        self._h = self.R.rp_create(_p(v), _p(i), _p(inst), inst.size, _p(m), trace, oracle_scene._h)
        if sky is not None:
            self._sky = np.ascontiguousarray(sky, np.float32)
            self.R.rp_set_skybox(self._h, _p(self._sky), self._sky.shape[1], self._sky.shape[0])
        if textures_f32:
            self._tex = [np.ascontiguousarray(t, np.float32) for t in textures_f32]
            ptrs = (ctypes.c_void_p * len(self._tex))(*[t.ctypes.data for t in self._tex])
            wh = np.array([[t.shape[1], t.shape[0]] for t in self._tex], np.int32)
            self._wh = wh
            self.R.rp_set_textures(self._h, ctypes.cast(ptrs, _vp), _p(wh), len(self._tex))
            self._ptrs = ptrs

    def bake_probe(self, origin, W, H, flags, light):
        """-> (coeffs [16,3] float64, image [H,W,4] float32, shadow rays)"""
        o = np.ascontiguousarray(origin, np.float32).reshape(3)
        l = np.ascontiguousarray(light, np.float32).reshape(3)
        img = np.zeros((H, W, 4), np.float32)
        out = np.zeros((16, 3), np.float64)
        n = self.R.rp_bake_probe(self._h, _p(o), W, H, int(flags), _p(l), _p(img), _p(out))
        return out, img, int(n)

    def bake_probe_viewer_hit(self, origin, W, H, flags, light, grid_step, lmax, shadow_bias, ambient, c_diffuse, c_specular, c_gloss,
                              sh_coeffs):
        """The same probe with the VIEWER's closest-hit shader (shaders/main.rchit: direct light + the gather over the baked
        probes, probe-visibility rays answered by shaders/sh.rmiss). sh_coeffs: [343, (lmax+1)^2, 3] float32, the SHCoeffs buffer
        of a 7x7x7 grid. -> (coeffs [16,3] float64, image [H,W,4] float32, flagged rays)"""
        o = np.ascontiguousarray(origin, np.float32).reshape(3)
        l = np.ascontiguousarray(light, np.float32).reshape(3)
        g = np.ascontiguousarray(grid_step, np.float32).reshape(3)
        sh = np.ascontiguousarray(sh_coeffs, np.float32)
        assert sh.shape == (343, (lmax + 1) ** 2, 3)
        # The shader indexes the buffer with floor(hitPosition / gridStep) unclamped (main.rchit:126, sh.rmiss:25): a hit a hair
        # outside the grid reads one probe layer before / after it. A Vulkan buffer tolerates that; here the buffer gets a margin
        # of 64 probes on either side (a copy of the first / last probe) and the shader sees the interior.
        pad = 64
        buf = np.concatenate([np.repeat(sh[:1], pad, 0), sh, np.repeat(sh[-1:], pad, 0)], 0)
        inner = ctypes.c_void_p(buf.ctypes.data + pad * sh.shape[1] * 3 * 4)
        img = np.zeros((H, W, 4), np.float32)
        out = np.zeros((16, 3), np.float64)
        n = self.R.rp_bake_probe_viewer_hit(self._h, _p(o), W, H, int(flags), _p(l), _p(g), int(lmax), float(shadow_bias), float(ambient),
                                            float(c_diffuse), float(c_specular), float(c_gloss), inner, _p(img), _p(out))
        return out, img, int(n)

    def close(self):
        if self._h:
            self.R.rp_destroy(self._h)
            self._h = None


def num_threads():
    return int(lib().vo_num_threads())


def set_num_threads(n):
    lib().vo_set_num_threads(int(n))


def sh_basis(dirs):
    d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros((d.shape[0], 25), np.float32)
    lib().vo_sh_basis(_p(d), d.shape[0], _p(out))
    return out


def probe_dirs(W, H):
    t = np.zeros((H, W, 3), np.float32)
    r = np.zeros((H, W, 3), np.float32)
    w = np.zeros((H, W), np.float32)
    lib().vo_probe_dirs(W, H, _p(t), _p(r), _p(w))
    return t, r, w


def _texels(texels):
    t = np.ascontiguousarray(texels)
    assert t.ndim == 3 and t.shape[2] == 4 and t.dtype in (np.uint8, np.float32)
    return t, (0 if t.dtype == np.uint8 else 1)


def skybox_project(texels, order=3):
    t, fmt = _texels(texels)
    out = np.zeros((16, 3), np.float32)
    lib().vo_skybox_project(_p(t), fmt, t.shape[1], t.shape[0], order, _p(out))
    return out


def envmap_project(texels, order=3):
    t, fmt = _texels(texels)
    out = np.zeros((16, 3), np.float32)
    lib().vo_envmap_project(_p(t), fmt, t.shape[1], t.shape[0], order, _p(out))
    return out


def sh_reconstruct(coeffs, order, W, H):
    c = np.ascontiguousarray(coeffs, np.float32).reshape(48)
    out = np.zeros((H, W, 3), np.float32)
    lib().vo_sh_reconstruct(_p(c), order, W, H, _p(out))
    return out


def probe_positions(settings):
    n = settings.probes[0] * settings.probes[1] * settings.probes[2]
    out = np.zeros((n, 3), np.float32)
    lib().vo_probe_positions(ctypes.byref(settings), _p(out))
    return out


def probe_positions_literal(bounds, counts):
    b = np.ascontiguousarray(bounds, np.float32).reshape(6)
    c = np.ascontiguousarray(counts, np.int32).reshape(3)
    out = np.zeros((int(np.prod(c)), 3), np.float32)
    step = np.zeros(3, np.float32)
    n = lib().vo_probe_positions_literal(_p(b), _p(c), _p(out), _p(step))
    assert n == out.shape[0]
    return out, step


class Scene:
    def __init__(self, scene):
        self._keep = [np.ascontiguousarray(scene[k]) for k in ("vertices", "indices", "instances", "materials")]
        v, i, inst, m = self._keep
        self._h = lib().vo_scene_create(_p(v), v.size, _p(i), i.size, _p(inst), inst.size, _p(m), m.size)
        if scene.get("textures"):
            self.set_textures(scene["textures"])

    def set_textures(self, textures):
        import importlib
        arr, keep = importlib.import_module("vulkan-light-bakery_b200").pack_textures(textures)
        lib().vo_scene_set_textures(self._h, ctypes.cast(arr, ctypes.c_void_p), len(textures))

    def tex_sample(self, tex, uv):
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((uv.shape[0], 3), np.float32)
        lib().vo_tex_sample(self._h, int(tex), _p(uv), uv.shape[0], _p(out))
        return out

    def close(self):
        if self._h:
            lib().vo_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_triangles(self):
        return int(lib().vo_scene_num_triangles(self._h))

    def bounds(self, tight=False):
        out = np.zeros(6, np.float32)
        lib().vo_scene_bounds(self._h, int(bool(tight)), _p(out))
        return out

    def triangles(self):
        out = np.zeros((self.n_triangles, 9), np.float32)
        lib().vo_scene_triangles(self._h, _p(out))
        return out

    def set_skybox(self, texels):
        t, fmt = _texels(texels)
        lib().vo_scene_set_skybox(self._h, _p(t), fmt, t.shape[1], t.shape[0])

    def trace_rays(self, origins, dirs, tmin=0.001, tmax=10000.0, accel=0, kind=0):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        ids = np.zeros(o.shape[0], np.int32)
        tuv = np.zeros((o.shape[0], 3), np.float32)
        lib().vo_trace_rays(self._h, _p(o), _p(d), o.shape[0], tmin, tmax, accel, kind, _p(ids), _p(tuv))
        return ids, tuv

    def bake_probes(self, settings, probe_ids=None, brute=False):
        """Whole slab (probe_ids None) or the listed x-fastest grid indices. Returns (coeffs, n_shadow_rays)."""
        if probe_ids is None:
            n = settings.n_slab_probes
            ids = None
        else:
            ids = np.ascontiguousarray(probe_ids, np.int64)
            n = ids.size
        out = np.zeros((n, 16, 3), np.float32)
        sr = lib().vo_bake_probes(self._h, ctypes.byref(settings), _p(ids), n, int(bool(brute)), _p(out))
        return out, int(sr)

    def bake_gather(self, settings, prev_full, probe_ids=None, brute=False):
        """One gather pass over prev_full ([n_probes,16,3], whole grid, x-fastest) -- None = direct pass."""
        if probe_ids is None:
            n, ids = settings.n_slab_probes, None
        else:
            ids = np.ascontiguousarray(probe_ids, np.int64)
            n = ids.size
        prev = None if prev_full is None else np.ascontiguousarray(prev_full, np.float32).reshape(settings.n_probes, 48)
        out = np.zeros((n, 16, 3), np.float32)
        sr = lib().vo_bake_gather(self._h, ctypes.byref(settings), _p(prev), _p(ids), n, int(bool(brute)), _p(out))
        return out, int(sr)

    def bake_multibounce(self, settings, brute=False):
        """1 + settings.bounces passes over the whole grid (include/vlb_bake.h: vlb_bake_probes with bounces > 0)."""
        prev = None
        for _ in range(1 + max(0, settings.bounces)):
            prev, _sr = self.bake_gather(settings, prev, brute=brute)
        return prev

    def probe_envmap_gather(self, settings, pos, prev_full, brute=False):
        """Environment image of one probe position in a gather pass over prev_full ([n_probes,16,3], whole grid)."""
        img = np.zeros((settings.dir_h, settings.dir_w, 3), np.float32)
        sh = np.zeros((16, 3), np.float32)
        p = np.ascontiguousarray(pos, np.float32).reshape(3)
        prev = np.ascontiguousarray(prev_full, np.float32).reshape(-1, 48)
        assert prev.shape[0] == settings.n_probes
        lib().vo_probe_envmap_gather(self._h, ctypes.byref(settings), _p(p), _p(prev), int(bool(brute)), _p(img), _p(sh))
        return img, sh

    def probe_envmap(self, settings, pos, brute=False):
        img = np.zeros((settings.dir_h, settings.dir_w, 3), np.float32)
        sh = np.zeros((16, 3), np.float32)
        p = np.ascontiguousarray(pos, np.float32).reshape(3)
        lib().vo_probe_envmap(self._h, ctypes.byref(settings), _p(p), int(bool(brute)), _p(img), _p(sh))
        return img, sh

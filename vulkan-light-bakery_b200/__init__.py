"""vulkan-light-bakery_b200 — host-side Python mirror of the reference's bake interface over
the C ABI of libvlb_bake.so (include/vlb_bake.h).

The reference's host code is C++ (src/baker, src/scene_manager.cpp, src/skybox_manager.cpp); its C++
counterpart here is inside the library (csrc/context.cu, gltf_scene.cpp, ...) and the `vlb_baker` executable
(csrc/vlb_baker_main.cpp). This module is the thin ctypes layer the parity tests
and bench.py use: same entry points, same argument meaning, same error behaviour (a failing call
raises VlbError carrying vlb_last_error(), as the reference throws std::runtime_error).

There is NO CPU fallback: if libvlb_bake.so is missing or no CUDA device is usable, every compute
entry point raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvlb_bake.so")
LIB_PATH = os.environ.get("VLB_LIB", LIB_PATH)   # A/B runs of instrumented / variant builds

# ---- layouts shared with shaders/structures.h (scalar layout) ------------------------------
VERTEX_DTYPE = np.dtype([("position", "<f4", (4,)), ("normal", "<f4", (3,)), ("uv0", "<f4", (2,)),
                         ("uv1", "<f4", (2,))])
INSTANCE_DTYPE = np.dtype([("first_index", "<u4"), ("index_count", "<u4"), ("first_vertex", "<u4"),
                           ("vertex_count", "<u4"), ("material_index", "<u4"), ("transform", "<f4", (12,))])
MATERIAL_DTYPE = np.dtype([("textures", "<i4", (8, 2)), ("base_color_factor", "<f4", (4,)),
                           ("emissive_factor", "<f4", (4,)), ("diffuse_factor", "<f4", (4,)),
                           ("specular_factor", "<f4", (3,)), ("metallic", "<f4"), ("roughness", "<f4"),
                           ("alpha_cutoff", "<f4"), ("pad", "<f4", (2,))])
assert VERTEX_DTYPE.itemsize == 44 and INSTANCE_DTYPE.itemsize == 68 and MATERIAL_DTYPE.itemsize == 144

FMT_RGBA8, FMT_RGBA32F = 0, 1
WRAP_REPEAT, WRAP_CLAMP_TO_EDGE, WRAP_MIRRORED_REPEAT = 0, 1, 2
FILTER_LINEAR, FILTER_NEAREST = 0, 1
SH_STRIDE = 48

SHADOW_RAYS = 1 << 0
SKYBOX_ON_MISS = 1 << 1
SRGB_ENCODE = 1 << 2
QUANTIZE_RGBA8 = 1 << 3
REFERENCE_PROBE_ORDER = 1 << 4
ACCUMULATE_ACROSS_PROBES = 1 << 5
SH_WORLD_FRAME = 1 << 6

STREAM_OWN = (1 << 64) - 1
TRACE_BVH, TRACE_BRUTE_FORCE = 0, 1
TRACE_CLOSEST, TRACE_ANY = 0, 1

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE, ERR_IO, ERR_UNSUPPORTED, ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6, -7


class BakeSettings(ctypes.Structure):
    """vlb_bake_settings (include/vlb_bake.h); defaults = the reference's hard-coded constants."""
    _fields_ = [("probes", ctypes.c_int32 * 3), ("origin", ctypes.c_float * 3), ("step", ctypes.c_float * 3),
                ("dir_w", ctypes.c_int32), ("dir_h", ctypes.c_int32), ("sh_order", ctypes.c_int32),
                ("light_pos", ctypes.c_float * 3), ("shadow_bias", ctypes.c_float), ("c_diffuse", ctypes.c_float),
                ("c_specular", ctypes.c_float), ("gloss", ctypes.c_float), ("ambient", ctypes.c_float),
                ("tmin", ctypes.c_float), ("tmax", ctypes.c_float), ("flags", ctypes.c_uint32),
                ("slab_k0", ctypes.c_int32), ("slab_k1", ctypes.c_int32), ("slab_stride", ctypes.c_int32),
                ("bounces", ctypes.c_int32), ("indirect_gain", ctypes.c_float), ("reserved", ctypes.c_int32)]

    def copy(self):
        c = BakeSettings()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(self), ctypes.sizeof(self))
        return c

    @property
    def n_probes(self):
        return self.probes[0] * self.probes[1] * self.probes[2]

    @property
    def slab(self):
        k1 = self.probes[2] if self.slab_k1 < 0 else self.slab_k1
        k0 = 0 if self.slab_k1 < 0 else self.slab_k0
        return k0, k1

    @property
    def slab_slices(self):
        """The z-slices this settings object bakes, in output order."""
        k0, k1 = self.slab
        return list(range(k0, k1, max(1, self.slab_stride)))

    @property
    def n_slab_probes(self):
        return self.probes[0] * self.probes[1] * len(self.slab_slices)


class Texture(ctypes.Structure):
    """vlb_texture (include/vlb_bake.h): one glTF texture = RGBA8 image + sampler state."""
    _fields_ = [("texels", ctypes.c_void_p), ("width", ctypes.c_int32), ("height", ctypes.c_int32),
                ("wrap_u", ctypes.c_int32), ("wrap_v", ctypes.c_int32), ("filter", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


def pack_textures(textures):
    """list of {"texels": uint8 [H, W, 4], "wrap_u", "wrap_v", "filter"} -> (ctypes array of Texture, keep-alive list)."""
    keep = []
    arr = (Texture * max(len(textures), 1))()
    for i, t in enumerate(textures):
        px = np.ascontiguousarray(t["texels"], np.uint8)
        if px.ndim != 3 or px.shape[2] != 4:
            raise ValueError("texture %d: texels must be uint8 [H, W, 4]" % i)
        keep.append(px)
        arr[i] = Texture(px.ctypes.data, px.shape[1], px.shape[0], int(t.get("wrap_u", WRAP_REPEAT)),
                         int(t.get("wrap_v", WRAP_REPEAT)), int(t.get("filter", FILTER_LINEAR)), 0)
    return arr, keep


class BvhStats(ctypes.Structure):
    _fields_ = [("n_triangles", ctypes.c_uint64), ("n_nodes", ctypes.c_uint64), ("max_leaf_size", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32), ("bounds", ctypes.c_float * 6), ("build_ms", ctypes.c_float),
                ("sort_ms", ctypes.c_float)]


class BakeStats(ctypes.Structure):
    _fields_ = [("n_probes", ctypes.c_uint64), ("n_primary_rays", ctypes.c_uint64), ("n_shadow_rays", ctypes.c_uint64),
                ("n_nodes_visited", ctypes.c_uint64), ("n_tris_tested", ctypes.c_uint64), ("kernel_ms", ctypes.c_float),
                ("total_ms", ctypes.c_float)]


class CachePeaks(ctypes.Structure):      # vlb_cache_peaks
    _fields_ = [("l2_read_gbs", ctypes.c_double), ("l1_read_gbs", ctypes.c_double), ("l1_scatter_lines_per_request", ctypes.c_double),
                ("l1_scatter_requests_per_s", ctypes.c_double), ("l1_scatter_wavefronts_per_s", ctypes.c_double),
                ("l1_scatter_gbs", ctypes.c_double)]


class VlbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("vlb error %d: %s" % (code, message))
        self.code = code


# every symbol include/vlb_bake.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "vlb_abi_version", "vlb_ctx_create", "vlb_ctx_destroy", "vlb_ctx_set_stream", "vlb_ctx_stream", "vlb_ctx_synchronize",
    "vlb_last_error", "vlb_ctx_launch_count", "vlb_scene_set_triangles", "vlb_scene_load_gltf", "vlb_gltf_probe", "vlb_scene_bounds", "vlb_bvh_build", "vlb_bvh_set_builder", "vlb_bvh_recommend_builder",
    "vlb_scene_set_textures", "vlb_gltf_texture", "vlb_image_load_rgba8", "vlb_image_load_rgba32f", "vlb_bake_probes_multi", "vlb_skybox_set", "vlb_skybox_set_async", "vlb_skybox_project_sh", "vlb_skybox_project_sh_batched", "vlb_skybox_project_sh_device",
    "vlb_skybox_project_sh_device_ptrs", "vlb_envmap_project_sh", "vlb_bake_settings_default", "vlb_bake_settings_from_bounds", "vlb_probe_positions",
    "vlb_bake_probes", "vlb_bake_probes_device", "vlb_bake_gather_device", "vlb_bake_last_stats", "vlb_trace_rays",
    "vlb_bake_serialize_gltf", "vlb_bake_deserialize_gltf", "vlb_diag_cache_peaks",
    "vlb_comm_get_unique_id", "vlb_comm_init_rank", "vlb_comm_init_all", "vlb_comm_destroy", "vlb_comm_info",
    "vlb_comm_sharded_uploads", "vlb_bake_probes_sharded_device", "vlb_bake_probes_sharded", "vlb_bake_probes_sharded_rows",
]

_lib = None


def load_library():
    """dlopen libvlb_bake.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python __graft_entry__.py` (build()) first; there is no CPU fallback"
                          % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, u64, u32, i32, f32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_float
    S = ctypes.POINTER(BakeSettings)
    sig = {
        "vlb_abi_version": (i32, []),
        "vlb_ctx_create": (i32, [i32, ctypes.POINTER(vp)]),
        "vlb_ctx_destroy": (None, [vp]),
        "vlb_ctx_set_stream": (i32, [vp, u64]),
        "vlb_ctx_synchronize": (i32, [vp]),
        "vlb_last_error": (ctypes.c_char_p, [vp]),
        "vlb_ctx_launch_count": (u64, [vp]),
        "vlb_ctx_stream": (u64, [vp]),
        "vlb_scene_set_triangles": (i32, [vp, vp, u64, vp, u64, vp, u32, vp, u32]),
        "vlb_scene_load_gltf": (i32, [vp, ctypes.c_char_p]),
        "vlb_gltf_probe": (i32, [ctypes.c_char_p, vp, vp]),
        "vlb_scene_bounds": (i32, [vp, i32, vp]),
        "vlb_bvh_build": (i32, [vp, ctypes.POINTER(BvhStats)]),
        "vlb_bvh_set_builder": (i32, [vp, i32, i32]),
        "vlb_bvh_recommend_builder": (i32, [u64, u64]),
        "vlb_scene_set_textures": (i32, [vp, vp, u32]),
        "vlb_gltf_texture": (i32, [ctypes.c_char_p, u32, vp, u64, vp]),
        "vlb_image_load_rgba8": (i32, [ctypes.c_char_p, vp, u64, vp]),
        "vlb_image_load_rgba32f": (i32, [ctypes.c_char_p, vp, u64, vp]),
        "vlb_bake_probes_multi": (i32, [vp, u32, S, vp]),
        "vlb_skybox_set": (i32, [vp, vp, i32, i32, i32]),
        "vlb_skybox_set_async": (i32, [vp, vp, i32, i32, i32]),
        "vlb_skybox_project_sh": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "vlb_skybox_project_sh_batched": (i32, [vp, vp, u32, i32, i32, i32, i32, vp]),
        "vlb_skybox_project_sh_device": (i32, [vp, vp, u64, u32, i32, i32, i32, i32, vp]),
        "vlb_skybox_project_sh_device_ptrs": (i32, [vp, vp, u32, i32, i32, i32, i32, vp]),
        "vlb_envmap_project_sh": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "vlb_bake_settings_default": (None, [S]),
        "vlb_bake_settings_from_bounds": (i32, [S, vp]),
        "vlb_probe_positions": (i32, [S, vp]),
        "vlb_bake_probes": (i32, [vp, S, vp]),
        "vlb_bake_probes_device": (i32, [vp, S, vp]),
        "vlb_bake_gather_device": (i32, [vp, S, vp, vp]),
        "vlb_bake_last_stats": (i32, [vp, ctypes.POINTER(BakeStats)]),
        "vlb_trace_rays": (i32, [vp, vp, vp, u64, f32, f32, i32, i32, vp, vp]),
        "vlb_bake_serialize_gltf": (i32, [ctypes.c_char_p, ctypes.c_char_p, vp, u64, S]),
        "vlb_bake_deserialize_gltf": (i32, [ctypes.c_char_p, vp, u64, ctypes.POINTER(u64), vp]),
        "vlb_diag_cache_peaks": (i32, [vp, ctypes.POINTER(CachePeaks)]),
        "vlb_comm_get_unique_id": (i32, [vp, u64]),
        "vlb_comm_init_rank": (i32, [vp, vp, i32, i32]),
        "vlb_comm_init_all": (i32, [vp, u32]),
        "vlb_comm_destroy": (i32, [vp]),
        "vlb_comm_info": (i32, [vp, vp, vp, vp]),
        "vlb_comm_sharded_uploads": (i32, [vp, i32]),
        "vlb_bake_probes_sharded_device": (i32, [vp, S, vp, vp]),
        "vlb_bake_probes_sharded": (i32, [vp, S, vp]),
        "vlb_bake_probes_sharded_rows": (i32, [vp, S, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def default_settings():
    s = BakeSettings()
    load_library().vlb_bake_settings_default(ctypes.byref(s))
    return s


def settings_from_bounds(s, bounds):
    b = np.ascontiguousarray(bounds, np.float32).reshape(6)
    r = load_library().vlb_bake_settings_from_bounds(ctypes.byref(s), _ptr(b))
    if r:
        raise VlbError(r, "vlb_bake_settings_from_bounds: bad probe counts")
    return s


def probe_positions(s):
    out = np.zeros((s.n_probes, 3), np.float32)
    r = load_library().vlb_probe_positions(ctypes.byref(s), _ptr(out))
    if r:
        raise VlbError(r, "vlb_probe_positions: bad settings")
    return out


class Context:
    """One baker context bound to one CUDA device (mirrors vlb::LightBaker's device ownership,
    src/baker/light_baker.cpp:20-74)."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = ctypes.c_void_p()
        r = self._lib.vlb_ctx_create(int(device), ctypes.byref(h))
        if r:
            raise VlbError(r, self._lib.vlb_last_error(None).decode())
        self._h = h
        self.stream_handle = None      # None: the ctx's own stream (set_stream)
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vlb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, r):
        if r:
            raise VlbError(r, self._lib.vlb_last_error(self._h).decode())

    # -- plumbing
    def set_stream(self, handle):
        """handle: cudaStream_t as int (0 = CUDA's legacy default stream); None = the ctx's own stream."""
        self._check(self._lib.vlb_ctx_set_stream(self._h, STREAM_OWN if handle is None else int(handle)))
        self.stream_handle = None if handle is None else int(handle)     # parallel.order_after_bake compares it with torch's stream

    @property
    def stream(self):
        """cudaStream_t (int) the ctx enqueues on; wrap it with torch.cuda.ExternalStream to record events on it."""
        return int(self._lib.vlb_ctx_stream(self._h))

    def synchronize(self):
        self._check(self._lib.vlb_ctx_synchronize(self._h))

    @property
    def launch_count(self):
        return int(self._lib.vlb_ctx_launch_count(self._h))

    # -- scene (SceneManager::pushScene output)
    def set_scene(self, scene):
        v = np.ascontiguousarray(scene["vertices"], VERTEX_DTYPE)
        i = np.ascontiguousarray(scene["indices"], np.uint32)
        inst = np.ascontiguousarray(scene["instances"], INSTANCE_DTYPE)
        m = np.ascontiguousarray(scene["materials"], MATERIAL_DTYPE)
        self._check(self._lib.vlb_scene_set_triangles(self._h, _ptr(v), v.size, _ptr(i), i.size, _ptr(inst), inst.size,
                                                      _ptr(m), m.size))
        self.set_textures(scene.get("textures", []))

    def set_textures(self, textures):
        """Scene_t::loadTextures: the textures vlb_material.base_color.index refers to (see pack_textures)."""
        arr, keep = pack_textures(textures)
        self._check(self._lib.vlb_scene_set_textures(self._h, ctypes.cast(arr, ctypes.c_void_p), len(textures)))

    def load_gltf(self, path):
        """SceneManager::pushScene: ingest a .gltf / .glb file."""
        self._check(self._lib.vlb_scene_load_gltf(self._h, os.fsencode(path)))

    def scene_bounds(self, tight=False):
        out = np.zeros(6, np.float32)
        self._check(self._lib.vlb_scene_bounds(self._h, int(bool(tight)), _ptr(out)))
        return out

    def set_bvh_builder(self, builder="lbvh", ploc_radius=0):
        """vlb_bvh_set_builder: "lbvh" (Karras, default) or "ploc" (agglomerative, radius 1..64, 0 = default 16)."""
        self._check(self._lib.vlb_bvh_set_builder(self._h, {"lbvh": 0, "ploc": 1}[builder], int(ploc_radius)))

    def build_bvh(self):
        st = BvhStats()
        self._check(self._lib.vlb_bvh_build(self._h, ctypes.byref(st)))
        return st

    # -- skybox (Skybox_t)
    @staticmethod
    def _texels(texels):
        t = np.ascontiguousarray(texels)
        if t.ndim != 3 or t.shape[2] != 4 or t.dtype not in (np.uint8, np.float32):
            raise ValueError("texels must be HxWx4 uint8 or float32")
        return t, (FMT_RGBA8 if t.dtype == np.uint8 else FMT_RGBA32F)

    def set_skybox(self, texels):
        t, fmt = self._texels(texels)
        self._check(self._lib.vlb_skybox_set(self._h, _ptr(t), fmt, t.shape[1], t.shape[0]))

    def set_skybox_async(self, texels):
        """vlb_skybox_set_async: returns at once; `texels` (ideally pinned) must stay alive until the next bake /
        synchronize. The array is kept referenced by the Context until then."""
        t, fmt = self._texels(texels)
        self._sky_keep = t
        self._check(self._lib.vlb_skybox_set_async(self._h, _ptr(t), fmt, t.shape[1], t.shape[0]))

    def skybox_project_sh(self, texels, order=3):
        t, fmt = self._texels(texels)
        out = np.zeros((16, 3), np.float32)
        self._check(self._lib.vlb_skybox_project_sh(self._h, _ptr(t), fmt, t.shape[1], t.shape[0], order, _ptr(out)))
        return out

    def envmap_project_sh(self, texels, order=3):
        t, fmt = self._texels(texels)
        out = np.zeros((16, 3), np.float32)
        self._check(self._lib.vlb_envmap_project_sh(self._h, _ptr(t), fmt, t.shape[1], t.shape[0], order, _ptr(out)))
        return out

    def skybox_project_sh_batched(self, maps, order=3):
        ts = [self._texels(m) for m in maps]
        fmt, shape = ts[0][1], ts[0][0].shape
        if any(t.shape != shape or f != fmt for t, f in ts):
            raise ValueError("all maps must have the same size and format")
        arr = (ctypes.c_void_p * len(ts))(*[t.ctypes.data for t, _ in ts])
        out = np.zeros((len(ts), 16, 3), np.float32)
        self._check(self._lib.vlb_skybox_project_sh_batched(self._h, arr, len(ts), fmt, shape[1], shape[0], order, _ptr(out)))
        return out

    def skybox_project_sh_device(self, d_texels, map_stride, n_maps, fmt, width, height, order, d_out):
        self._check(self._lib.vlb_skybox_project_sh_device(self._h, int(d_texels), int(map_stride), int(n_maps), fmt,
                                                           width, height, order, int(d_out)))

    def skybox_project_sh_device_ptrs(self, d_maps, fmt, width, height, order, d_out):
        """d_maps: device addresses of n independent maps; d_out: device address of n x 48 floats."""
        arr = d_maps if isinstance(d_maps, ctypes.Array) else (ctypes.c_void_p * len(d_maps))(*[int(x) for x in d_maps])
        self._check(self._lib.vlb_skybox_project_sh_device_ptrs(self._h, arr, len(arr), fmt, width, height, order, int(d_out)))

    # -- bake (LightBaker::bake)
    def bake_probes(self, s):
        n = s.n_slab_probes
        out = np.zeros((n, 16, 3), np.float32)
        self._check(self._lib.vlb_bake_probes(self._h, ctypes.byref(s), _ptr(out)))
        return out

    def bake_probes_device(self, s, d_out):
        self._check(self._lib.vlb_bake_probes_device(self._h, ctypes.byref(s), int(d_out)))

    def bake_gather_device(self, s, d_prev_full, d_out):
        """One gather pass: d_prev_full = previous pass over the WHOLE grid ([n_probes, 48] device floats) or 0."""
        self._check(self._lib.vlb_bake_gather_device(self._h, ctypes.byref(s), int(d_prev_full) if d_prev_full else None, int(d_out)))

    def cache_peaks(self):
        """Measured L2 / L1 read ceilings of this device (vlb_diag_cache_peaks), as a dict."""
        pk = CachePeaks()
        self._check(self._lib.vlb_diag_cache_peaks(self._h, ctypes.byref(pk)))
        return {k: getattr(pk, k) for k, _ in CachePeaks._fields_}

    def last_bake_stats(self):
        st = BakeStats()
        self._check(self._lib.vlb_bake_last_stats(self._h, ctypes.byref(st)))
        return st

    # -- multi-GPU: one NCCL rank per Context (include/vlb_bake.h, "multi-GPU")
    def comm_init_rank(self, unique_id, rank, world):
        """unique_id: the COMM_ID_BYTES bytes rank 0 got from comm_unique_id(), carried here by the host program."""
        buf = ctypes.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._check(self._lib.vlb_comm_init_rank(self._h, ctypes.cast(buf, ctypes.c_void_p), int(rank), int(world)))

    def comm_destroy(self):
        self._check(self._lib.vlb_comm_destroy(self._h))

    def comm_info(self):
        """(rank, world, NCCL version code) -- (0, 1, v) without a communicator."""
        r, w, v = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        self._check(self._lib.vlb_comm_info(self._h, ctypes.byref(r), ctypes.byref(w), ctypes.byref(v)))
        return r.value, w.value, v.value

    def comm_sharded_uploads(self, enable=True):
        self._check(self._lib.vlb_comm_sharded_uploads(self._h, int(bool(enable))))

    def bake_probes_sharded_device(self, s, d_prev_full, d_full_out):
        """This rank's cyclic z-slices + all-gather: d_full_out ([n_probes, 48] device floats) holds the whole grid on every rank."""
        self._check(self._lib.vlb_bake_probes_sharded_device(self._h, ctypes.byref(s), int(d_prev_full) if d_prev_full else None, int(d_full_out)))

    def bake_probes_sharded(self, s, want_output=True):
        """1 + s.bounces passes over the communicator; returns the whole grid [n_probes, 16, 3] (None if not wanted)."""
        out = np.zeros((s.n_probes, 16, 3), np.float32) if want_output else None
        self._check(self._lib.vlb_bake_probes_sharded(self._h, ctypes.byref(s), _ptr(out)))
        return out

    def bake_probes_sharded_rows(self, s, host_grid_ptr):
        """vlb_bake_probes_sharded_rows: this rank's z-slices land in the host grid at `host_grid_ptr` (address of a
        [n_probes, 48] float32 buffer every rank of the host shares); synchronous."""
        self._check(self._lib.vlb_bake_probes_sharded_rows(self._h, ctypes.byref(s), int(host_grid_ptr)))

    # -- validation
    def trace_rays(self, origins, dirs, tmin=0.001, tmax=10000.0, accel=TRACE_BVH, kind=TRACE_CLOSEST):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        ids = np.full(o.shape[0], -2, np.int32)
        tuv = np.zeros((o.shape[0], 3), np.float32)
        self._check(self._lib.vlb_trace_rays(self._h, _ptr(o), _ptr(d), o.shape[0], tmin, tmax, accel, kind, _ptr(ids),
                                             _ptr(tuv)))
        return ids, tuv


def gltf_probe(path):
    """Host-only parse: ({vertices, indices, instances, materials, triangles}, reference-mode bounds)."""
    lib = load_library()
    counts = (ctypes.c_uint64 * 5)()
    bounds = np.zeros(6, np.float32)
    r = lib.vlb_gltf_probe(os.fsencode(path), counts, _ptr(bounds))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    keys = ("vertices", "indices", "instances", "materials", "triangles")
    return dict(zip(keys, (int(c) for c in counts))), bounds


COMM_ID_BYTES = 128


def recommend_builder(n_triangles, n_primary_rays):
    """vlb_bvh_recommend_builder: "lbvh" or "ploc" for a bake of that many primary rays (per GPU) through that many triangles."""
    return ("lbvh", "ploc")[load_library().vlb_bvh_recommend_builder(int(n_triangles), int(n_primary_rays))]


def comm_unique_id():
    """vlb_comm_get_unique_id (rank 0): bytes to hand to every rank's Context.comm_init_rank."""
    lib = load_library()
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    r = lib.vlb_comm_get_unique_id(ctypes.cast(buf, ctypes.c_void_p), COMM_ID_BYTES)
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    return buf.raw


def comm_init_all(contexts):
    """vlb_comm_init_all: the contexts of ONE process (distinct devices) become ranks 0..n-1."""
    lib = load_library()
    arr = (ctypes.c_void_p * len(contexts))(*[c._h for c in contexts])
    r = lib.vlb_comm_init_all(ctypes.cast(arr, ctypes.c_void_p), len(contexts))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(contexts[0]._h).decode())


def bake_probes_multi(contexts, s):
    """vlb_bake_probes_multi: one host process, several contexts (one per GPU) holding the same scene; returns the
    whole grid [n_probes, 16, 3]."""
    lib = load_library()
    arr = (ctypes.c_void_p * len(contexts))(*[c._h for c in contexts])
    out = np.zeros((s.n_probes, 16, 3), np.float32)
    r = lib.vlb_bake_probes_multi(ctypes.cast(arr, ctypes.c_void_p), len(contexts), ctypes.byref(s), _ptr(out))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(contexts[0]._h).decode())
    return out


def image_load_rgba8(path):
    """stbi_load(path, ..., 4) of the reference's skybox / texture ingest: PNG or baseline JPEG -> uint8 [H, W, 4]."""
    lib = load_library()
    size = np.zeros(2, np.int32)
    r = lib.vlb_image_load_rgba8(os.fsencode(path), None, 0, _ptr(size))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    px = np.zeros((int(size[1]), int(size[0]), 4), np.uint8)
    r = lib.vlb_image_load_rgba8(os.fsencode(path), _ptr(px), px.nbytes, _ptr(size))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    return px


def image_load_rgba32f(path):
    """Radiance RGBE (.hdr) -> linear float32 [H, W, 4] (alpha 1): RGBA32F equirect skyboxes from a file."""
    lib = load_library()
    size = np.zeros(2, np.int32)
    r = lib.vlb_image_load_rgba32f(os.fsencode(path), None, 0, _ptr(size))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    px = np.zeros((int(size[1]), int(size[0]), 4), np.float32)
    r = lib.vlb_image_load_rgba32f(os.fsencode(path), _ptr(px), px.nbytes, _ptr(size))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    return px


def gltf_texture(path, index):
    """Texture `index` of a glTF as the loader decodes it: dict like pack_textures takes, plus "used"."""
    lib = load_library()
    info = np.zeros(6, np.int32)
    r = lib.vlb_gltf_texture(os.fsencode(path), index, None, 0, _ptr(info))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    px = np.zeros((int(info[1]), int(info[0]), 4), np.uint8)
    r = lib.vlb_gltf_texture(os.fsencode(path), index, _ptr(px), px.nbytes, _ptr(info))
    if r != 0:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    return {"texels": px, "wrap_u": int(info[2]), "wrap_v": int(info[3]), "filter": int(info[4]), "used": bool(info[5])}


def serialize_gltf(in_path, out_path, coeffs, settings):
    c = np.ascontiguousarray(coeffs, np.float32).reshape(-1, SH_STRIDE)
    lib = load_library()
    r = lib.vlb_bake_serialize_gltf(os.fsencode(in_path), os.fsencode(out_path), _ptr(c), c.shape[0], ctypes.byref(settings))
    if r:
        raise VlbError(r, lib.vlb_last_error(None).decode())


def deserialize_gltf(path):
    lib = load_library()
    n = ctypes.c_uint64()
    step = np.zeros(3, np.float32)
    r = lib.vlb_bake_deserialize_gltf(os.fsencode(path), None, 0, ctypes.byref(n), _ptr(step))
    if r:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    out = np.zeros(n.value, np.float32)
    r = lib.vlb_bake_deserialize_gltf(os.fsencode(path), _ptr(out), out.size, ctypes.byref(n), _ptr(step))
    if r:
        raise VlbError(r, lib.vlb_last_error(None).decode())
    return out.reshape(-1, 16, 3), step

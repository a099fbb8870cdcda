"""Synthetic inputs of the named BASELINE.json shapes (there is no network for assets): the
procedural "atrium" scene, HDR equirect skyboxes, small KAT scenes, and the reference's own
default cube fixture (src/vendor/default_blender_cube.gltf geometry restated as arrays).

Everything is produced in the layouts the C ABI takes (include/vlb_bake.h): shader::Vertex (44 B),
u32 indices, instances (primitive + world transform) and shader::Material (144 B).
"""
import math

import numpy as np

from . import INSTANCE_DTYPE, MATERIAL_DTYPE, VERTEX_DTYPE

PALETTE = np.array([
    [0.80, 0.80, 0.80], [0.75, 0.25, 0.20], [0.20, 0.55, 0.30], [0.25, 0.35, 0.75],
    [0.85, 0.75, 0.35], [0.60, 0.40, 0.25], [0.90, 0.90, 0.85], [0.35, 0.35, 0.40],
    [0.70, 0.55, 0.65], [0.45, 0.65, 0.70], [0.95, 0.60, 0.30], [0.30, 0.30, 0.30],
    [0.65, 0.70, 0.45], [0.55, 0.25, 0.45], [0.85, 0.85, 0.95], [0.50, 0.50, 0.50]], np.float32)


def make_materials(colors=PALETTE):
    m = np.zeros(len(colors), MATERIAL_DTYPE)
    m["textures"][:, :, 0] = -1                     # Texture::index = -1 (structures.h:48)
    m["base_color_factor"][:, :3] = colors
    m["base_color_factor"][:, 3] = 1.0
    m["metallic"] = 1.0
    m["roughness"] = 1.0
    return m


def identity12():
    return np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def trs12(translate=(0, 0, 0), scale=(1, 1, 1), rot_y=0.0):
    c, s = math.cos(rot_y), math.sin(rot_y)
    r = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64) * np.asarray(scale, np.float64)[None, :]
    m = np.zeros((3, 4), np.float64)
    m[:, :3] = r
    m[:, 3] = translate
    return m.astype(np.float32).reshape(12)


class _Builder:
    def __init__(self):
        self.v, self.i, self.inst = [], [], []
        self.nv = 0
        self.ni = 0

    def add_mesh(self, pos, nrm, idx):
        """Appends a mesh, returns its (first_index, index_count, first_vertex, vertex_count)."""
        v = np.zeros(len(pos), VERTEX_DTYPE)
        v["position"][:, :3] = pos
        v["position"][:, 3] = 1.0
        n = np.asarray(nrm, np.float32)
        ln = np.linalg.norm(n, axis=1, keepdims=True)
        v["normal"] = n / np.maximum(ln, 1e-20)
        idx = np.asarray(idx, np.uint32).reshape(-1)
        rec = (self.ni, idx.size, self.nv, len(pos))
        self.v.append(v)
        self.i.append(idx)
        self.nv += len(pos)
        self.ni += idx.size
        return rec

    def add_instance(self, mesh, transform, material):
        r = np.zeros(1, INSTANCE_DTYPE)
        r["first_index"], r["index_count"], r["first_vertex"], r["vertex_count"] = mesh
        r["material_index"] = material
        r["transform"] = transform
        self.inst.append(r)

    def n_tris(self):
        return sum(int(r["index_count"][0]) // 3 for r in self.inst)

    def finish(self, materials):
        return {"vertices": np.concatenate(self.v) if self.v else np.zeros(0, VERTEX_DTYPE),
                "indices": np.concatenate(self.i) if self.i else np.zeros(0, np.uint32),
                "instances": np.concatenate(self.inst) if self.inst else np.zeros(0, INSTANCE_DTYPE),
                "materials": materials}


def _grid(nu, nv):
    """(nu+1)x(nv+1) vertex lattice -> 2*nu*nv triangles (indices into the lattice)."""
    a = (np.arange(nv)[:, None] * (nu + 1) + np.arange(nu)[None, :]).reshape(-1)
    b, c, d = a + 1, a + nu + 1, a + nu + 2
    return np.stack([a, b, d, a, d, c], 1).reshape(-1, 3)


def _plane(origin, du, dv, nu, nv, normal, height=None):
    u = np.linspace(0, 1, nu + 1)
    v = np.linspace(0, 1, nv + 1)
    uu, vv = np.meshgrid(u, v)
    o, du, dv, normal = (np.asarray(a, np.float64) for a in (origin, du, dv, normal))
    p = o[None, None, :] + uu[..., None] * du + vv[..., None] * dv
    n = np.broadcast_to(normal, p.shape).copy()
    if height is not None:
        p = p + height[..., None] * normal
        # finite-difference normals of the displaced sheet
        gu = np.gradient(height, axis=1) * nu / max(np.linalg.norm(du), 1e-9)
        gv = np.gradient(height, axis=0) * nv / max(np.linalg.norm(dv), 1e-9)
        n = normal[None, None, :] - gu[..., None] * du / np.linalg.norm(du) - gv[..., None] * dv / np.linalg.norm(dv)
    idx = _grid(nu, nv)
    # orient the triangles so that the geometric normal agrees with `normal`
    if np.dot(np.cross(du, dv), normal) < 0:
        idx = idx[:, ::-1]
    return p.reshape(-1, 3).astype(np.float32), n.reshape(-1, 3).astype(np.float32), idx


def _cylinder(seg, rings):
    """Unit cylinder: radius 1 around +y, y in [0,1], open ends, smooth normals."""
    th = np.linspace(0, 2 * np.pi, seg + 1)
    y = np.linspace(0, 1, rings + 1)
    tt, yy = np.meshgrid(th, y)
    p = np.stack([np.cos(tt), yy, np.sin(tt)], -1)
    n = np.stack([np.cos(tt), np.zeros_like(tt), np.sin(tt)], -1)
    return p.reshape(-1, 3).astype(np.float32), n.reshape(-1, 3).astype(np.float32), _grid(seg, rings)[:, ::-1]


def _arch(arc_seg, tube_seg, radius, tube):
    """Half torus in the xy plane (arch spanning x in [-radius, radius], apex at +y)."""
    a = np.linspace(0, np.pi, arc_seg + 1)
    b = np.linspace(0, 2 * np.pi, tube_seg + 1)
    aa, bb = np.meshgrid(a, b)
    cx, cy = np.cos(aa), np.sin(aa)
    p = np.stack([(radius + tube * np.cos(bb)) * cx, (radius + tube * np.cos(bb)) * cy, tube * np.sin(bb)], -1)
    n = np.stack([np.cos(bb) * cx, np.cos(bb) * cy, np.sin(bb)], -1)
    return p.reshape(-1, 3).astype(np.float32), n.reshape(-1, 3).astype(np.float32), _grid(arc_seg, tube_seg)


HALL = (30.0, 12.0, 18.0)
ATRIUM_LIGHT = (10.0, 10.0, 6.4)     # the reference's (1,10,1) point light moved into the hall


def atrium(n_tris=262144, seed=7):
    """Procedural atrium with EXACTLY n_tris triangles (BASELINE.json configs 2-4):
    open-top box hall 30x12x18 m, two tiers of tessellated columns joined by arches, gallery
    slabs, a displaced-grid floor, factor-only materials from a 16-entry palette. Columns and
    arches are instanced (shared meshes + per-instance transforms with non-uniform scale)."""
    rng = np.random.default_rng(seed)
    t = math.sqrt(n_tris / 262144.0)
    q = lambda x: max(2, int(round(x * t)))
    b = _Builder()
    X, Y, Z = HALL
    # walls (inward normals), each its own mesh with slight relief
    walls = [((0, 0, 0), (X, 0, 0), (0, Y, 0), (0, 0, 1), q(96), q(40)),
             ((0, 0, Z), (X, 0, 0), (0, Y, 0), (0, 0, -1), q(96), q(40)),
             ((0, 0, 0), (0, 0, Z), (0, Y, 0), (1, 0, 0), q(56), q(40)),
             ((X, 0, 0), (0, 0, Z), (0, Y, 0), (-1, 0, 0), q(56), q(40))]
    for w, (o, du, dv, nrm, nu, nv) in enumerate(walls):
        h = 0.03 * rng.standard_normal((nv + 1, nu + 1))
        h[0, :] = h[-1, :] = 0.0
        h[:, 0] = h[:, -1] = 0.0
        p, n, idx = _plane(o, du, dv, nu, nv, nrm, h)
        b.add_instance(b.add_mesh(p, n, idx), identity12(), 6 + (w & 1) * 8)
    # columns: two rows x 8, two tiers (instanced unit cylinder)
    cyl = b.add_mesh(*_cylinder(q(32), q(48)))
    xs = np.linspace(3.0, X - 3.0, 8)
    for tier, (y0, hgt, rad) in enumerate([(0.0, 5.5, 0.42), (6.0, 5.0, 0.30)]):
        for row, z in enumerate((4.0, Z - 4.0)):
            for c, x in enumerate(xs):
                b.add_instance(cyl, trs12((x, y0, z), (rad, hgt, rad * (1.0 + 0.15 * ((c + row) & 1))),
                                          rot_y=0.37 * c + row), 1 + ((c + tier + row) % 5))
    # arches between neighbouring columns (instanced half torus)
    gap = float(xs[1] - xs[0])
    arch = b.add_mesh(*_arch(q(32), q(16), 0.5, 0.12))
    for tier, y in enumerate((5.5 - 0.5 * gap * 0.0, 11.0)):
        for z in (4.0, Z - 4.0):
            for c in range(7):
                b.add_instance(arch, trs12((0.5 * (xs[c] + xs[c + 1]), y - 0.0, z), (gap, 1.2, 1.0)), 8 + (c % 4))
    # gallery slabs between the long walls and the column rows (top and bottom sheets)
    for z0, z1 in ((0.0, 4.0), (Z - 4.0, Z)):
        for y, nrm in ((5.75, (0, 1, 0)), (5.55, (0, -1, 0))):
            p, n, idx = _plane((0, y, z0), (X, 0, 0), (0, 0, z1 - z0), q(60), q(8), nrm)
            b.add_instance(b.add_mesh(p, n, idx), identity12(), 12)
    # floor: displaced grid sized to land on the exact triangle budget
    rest = n_tris - b.n_tris()
    if rest < 8:
        raise ValueError("triangle budget too small for the atrium (%d)" % n_tris)
    nu = max(2, int(round(math.sqrt(rest / 2.0 * X / Z))))
    nv = max(1, rest // (2 * nu))
    fx, fz = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1))
    h = 0.05 * np.sin(17 * fx) * np.cos(11 * fz) + 0.02 * rng.standard_normal(fx.shape)
    p, n, idx = _plane((0, 0, 0), (X, 0, 0), (0, 0, Z), nu, nv, (0, 1, 0), h)
    b.add_instance(b.add_mesh(p, n, idx), identity12(), 5)
    # remainder: small rubble triangles resting on the floor
    rest = n_tris - b.n_tris()
    if rest > 0:
        c = np.stack([rng.uniform(1, X - 1, rest), np.full(rest, 0.12), rng.uniform(1, Z - 1, rest)], 1)
        d = rng.normal(size=(rest, 3, 3)) * 0.08
        p = (c[:, None, :] + d).reshape(-1, 3).astype(np.float32)
        fn = np.cross(d[:, 1] - d[:, 0], d[:, 2] - d[:, 0])
        n = np.repeat(fn, 3, axis=0).astype(np.float32)
        b.add_instance(b.add_mesh(p, n, np.arange(rest * 3).reshape(-1, 3)), identity12(), 11)
    scene = b.finish(make_materials())
    assert b.n_tris() == n_tris, (b.n_tris(), n_tris)
    return scene


def default_cube():
    """Geometry of the reference's only in-repo scene, src/vendor/default_blender_cube.gltf:
    24 vertices / 12 triangles, POSITION min/max +-1, one material with baseColorFactor 0.8
    (SURVEY §4). Vertex order follows the fixture's accessors (4 vertices per face)."""
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)),
             ((0, 1, 0), (0, 0, 1), (1, 0, 0)), ((0, -1, 0), (1, 0, 0), (0, 0, 1)),
             ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    pos, nrm, idx = [], [], []
    for f, (n, u, v) in enumerate(faces):
        n, u, v = (np.array(a, np.float32) for a in (n, u, v))
        for su, sv in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            pos.append(n + su * u + sv * v)
            nrm.append(n)
        idx += [[4 * f, 4 * f + 1, 4 * f + 2], [4 * f, 4 * f + 2, 4 * f + 3]]
    b = _Builder()
    b.add_instance(b.add_mesh(np.array(pos), np.array(nrm), np.array(idx)), identity12(), 0)
    return b.finish(make_materials(np.array([[0.8, 0.8, 0.8]], np.float32)))


def small_room(n_side=6, seed=3):
    """A few-hundred-triangle closed room with two boxes: small enough for brute-force oracles."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    S = 4.0
    planes = [((0, 0, 0), (S, 0, 0), (0, 0, S), (0, 1, 0)), ((0, S, 0), (S, 0, 0), (0, 0, S), (0, -1, 0)),
              ((0, 0, 0), (S, 0, 0), (0, S, 0), (0, 0, 1)), ((0, 0, S), (S, 0, 0), (0, S, 0), (0, 0, -1)),
              ((0, 0, 0), (0, 0, S), (0, S, 0), (1, 0, 0))]            # x = S side left open: sky visible
    for k, (o, du, dv, n) in enumerate(planes):
        h = 0.02 * rng.standard_normal((n_side + 1, n_side + 1))
        b.add_instance(b.add_mesh(*_plane(o, du, dv, n_side, n_side, n, h)), identity12(), k % 5)
    cube = default_cube()
    m = b.add_mesh(cube["vertices"]["position"][:, :3], cube["vertices"]["normal"], cube["indices"].reshape(-1, 3))
    b.add_instance(m, trs12((1.2, 0.5, 1.5), (0.5, 0.5, 0.5), rot_y=0.5), 7)
    b.add_instance(m, trs12((2.8, 0.9, 2.6), (0.4, 0.9, 0.3), rot_y=-0.8), 9)
    return b.finish(make_materials())


def pattern_texture(width, height, seed=0):
    """Procedural RGBA8 texture: coloured checker + gradient + per-texel noise (so neighbouring texels differ)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width]
    cell = ((x * 8 // max(width, 1)) + (y * 8 // max(height, 1))) & 1
    base = np.stack([80 + 120 * cell, 60 + 150 * x / max(width - 1, 1), 200 - 140 * y / max(height - 1, 1)], -1)
    t = np.zeros((height, width, 4), np.uint8)
    t[..., :3] = np.clip(base + rng.integers(-25, 26, base.shape), 0, 255).astype(np.uint8)
    t[..., 3] = 255
    return t


def small_room_textured(n_side=6, seed=3):
    """small_room with texture coordinates on every vertex and four baseColor textures (one per sampler flavour:
    linear/repeat, linear/clamp, linear/mirror, nearest/repeat) on materials 0, 1, 2 and 7; uv0 runs past [0, 1]
    so that the address modes matter. The remaining materials keep the factor path (env_map.rchit:43-47)."""
    from . import WRAP_REPEAT, WRAP_CLAMP_TO_EDGE, WRAP_MIRRORED_REPEAT, FILTER_LINEAR, FILTER_NEAREST
    scene = small_room(n_side, seed)
    rng = np.random.default_rng(seed + 100)
    v = scene["vertices"]
    p = v["position"][:, :3]
    # planar-ish mapping with an irrational scale and an offset: covers about [-0.7, 2.3]
    v["uv0"][:, 0] = 0.37 * p[:, 0] + 0.29 * p[:, 2] - 0.7 + 0.01 * rng.standard_normal(len(v))
    v["uv0"][:, 1] = 0.41 * p[:, 1] + 0.23 * p[:, 2] - 0.6 + 0.01 * rng.standard_normal(len(v))
    m = scene["materials"]
    for slot, mat in enumerate((0, 1, 2, 7)):
        m["textures"][mat, 2, 0] = slot                # Textures::baseColor.index (structures.h:52-61)
    scene["textures"] = [
        {"texels": pattern_texture(16, 8, 1), "wrap_u": WRAP_REPEAT, "wrap_v": WRAP_REPEAT, "filter": FILTER_LINEAR},
        {"texels": pattern_texture(5, 7, 2), "wrap_u": WRAP_CLAMP_TO_EDGE, "wrap_v": WRAP_CLAMP_TO_EDGE, "filter": FILTER_LINEAR},
        {"texels": pattern_texture(32, 32, 3), "wrap_u": WRAP_MIRRORED_REPEAT, "wrap_v": WRAP_REPEAT, "filter": FILTER_LINEAR},
        {"texels": pattern_texture(4, 4, 4), "wrap_u": WRAP_REPEAT, "wrap_v": WRAP_MIRRORED_REPEAT, "filter": FILTER_NEAREST},
    ]
    return scene


def hdr_sky(width, height, seed=1):
    """BASELINE C1 input: smooth HDR sky rgb = a + b*max(0, d.s)^p plus U[0,0.1) per-texel noise,
    alpha 1, RGBA32F."""
    rng = np.random.default_rng(seed)
    y = (np.arange(height) + 0.5) / height * np.pi
    x = (np.arange(width) + 0.5) / width * 2 * np.pi
    st, ct = np.sin(y)[:, None], np.cos(y)[:, None]
    d = np.stack([st * np.sin(x)[None, :], np.broadcast_to(ct, (height, width)), st * np.cos(x)[None, :]], -1)
    sun = np.array([0.35, 0.8, 0.48])
    sun /= np.linalg.norm(sun)
    lobe = np.maximum(0.0, d @ sun) ** 24
    a = np.array([0.25, 0.35, 0.55])[None, None, :] * (0.6 + 0.4 * np.clip(d[..., 1:2], -1, 1))
    img = np.empty((height, width, 4), np.float32)
    img[..., :3] = a + np.array([6.0, 5.0, 3.5])[None, None, :] * lobe[..., None] + rng.uniform(0, 0.1, (height, width, 3))
    img[..., 3] = 1.0
    return img


def atrium_settings(probes=(16, 8, 16), dirs=(32, 32), order=2, bounds=None):
    """BakeSettings of BASELINE configs 2/3: grid spanning the bounds, equirect direction grid,
    flags SHADOW_RAYS | SKYBOX_ON_MISS | SRGB_ENCODE, sun moved into the hall."""
    from . import SHADOW_RAYS, SKYBOX_ON_MISS, SRGB_ENCODE, default_settings, settings_from_bounds
    s = default_settings()
    s.probes[:] = probes
    s.dir_w, s.dir_h = dirs
    s.sh_order = order
    s.light_pos[:] = ATRIUM_LIGHT
    s.flags = SHADOW_RAYS | SKYBOX_ON_MISS | SRGB_ENCODE
    if bounds is None:
        bounds = (0.0, 0.0, 0.0) + HALL
    settings_from_bounds(s, bounds)
    return s


def write_gltf(scene, path, index_dtype=np.uint16, embed=True, images_in_views=False):
    """Writes a scene dict (the arrays the C ABI takes) as a glTF 2.0 file: one mesh + one node per
    instance, the instance transform as the node's column-major `matrix`, materials as
    pbrMetallicRoughness.baseColorFactor. Loading the file back with vlb_scene_load_gltf must give
    the same geometry (tests/test_gltf.py). `embed`: base64 data URI, else a .bin next to the file."""
    import base64
    import json
    import os
    verts, idx, insts, mats = scene["vertices"], scene["indices"], scene["instances"], scene["materials"]
    textures = scene.get("textures", [])
    blob = bytearray()
    views, accessors, meshes, nodes = [], [], [], []

    def add_view(data, target=None):
        while len(blob) % 4:
            blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}
        if target:
            v["target"] = target
        blob.extend(data)
        views.append(v)
        return len(views) - 1

    for inst in insts:
        fv, nv = int(inst["first_vertex"]), int(inst["vertex_count"])
        fi, ni = int(inst["first_index"]), int(inst["index_count"])
        pos = np.ascontiguousarray(verts["position"][fv:fv + nv, :3], np.float32)
        nrm = np.ascontiguousarray(verts["normal"][fv:fv + nv], np.float32)
        ii = np.ascontiguousarray(idx[fi:fi + ni])
        dt = index_dtype if (nv <= np.iinfo(index_dtype).max + 1) else np.uint32
        comp = {np.uint8: 5121, np.uint16: 5123, np.uint32: 5125}[np.dtype(dt).type]
        a_pos = len(accessors)
        accessors.append({"bufferView": add_view(pos.tobytes(), 34962), "componentType": 5126, "count": nv, "type": "VEC3",
                          "min": [float(x) for x in pos.min(0)] if nv else [0, 0, 0],
                          "max": [float(x) for x in pos.max(0)] if nv else [0, 0, 0]})
        accessors.append({"bufferView": add_view(nrm.tobytes(), 34962), "componentType": 5126, "count": nv, "type": "VEC3"})
        accessors.append({"bufferView": add_view(ii.astype(dt).tobytes(), 34963), "componentType": comp, "count": ni, "type": "SCALAR"})
        prim = {"attributes": {"POSITION": a_pos, "NORMAL": a_pos + 1}, "indices": a_pos + 2}
        if textures:
            uv = np.ascontiguousarray(verts["uv0"][fv:fv + nv], np.float32)
            prim["attributes"]["TEXCOORD_0"] = len(accessors)
            accessors.append({"bufferView": add_view(uv.tobytes(), 34962), "componentType": 5126, "count": nv, "type": "VEC2"})
        if int(inst["material_index"]) < len(mats):
            prim["material"] = int(inst["material_index"])
        meshes.append({"primitives": [prim]})
        m = np.asarray(inst["transform"], np.float32).reshape(3, 4)
        col_major = [float(m[r, c]) for c in range(4) for r in range(3)]
        matrix = col_major[0:3] + [0.0] + col_major[3:6] + [0.0] + col_major[6:9] + [0.0] + col_major[9:12] + [1.0]
        nodes.append({"mesh": len(meshes) - 1, "matrix": matrix})
    jmats = []
    for m in mats:
        jm = {"pbrMetallicRoughness": {"baseColorFactor": [float(x) for x in m["base_color_factor"]]}}
        if int(m["textures"][2, 0]) >= 0:
            jm["pbrMetallicRoughness"]["baseColorTexture"] = {"index": int(m["textures"][2, 0])}
        jmats.append(jm)
    doc = {"asset": {"version": "2.0", "generator": "vulkan-light-bakery_b200.scenes.write_gltf"},
           "scene": 0, "scenes": [{"nodes": list(range(len(nodes)))}], "nodes": nodes, "meshes": meshes,
           "materials": jmats, "accessors": accessors, "bufferViews": views}
    if textures:
        # textures as PNG images (RGB when the alpha is all 255, else RGBA), alternating between data URIs / files and
        # bufferViews so that both image sources of the loader are exercised; one sampler per texture
        import io
        from PIL import Image
        wrap = {0: 10497, 1: 33071, 2: 33648}
        doc["images"], doc["samplers"], doc["textures"] = [], [], []
        for k, t in enumerate(textures):
            px = np.ascontiguousarray(t["texels"], np.uint8)
            img = Image.fromarray(np.ascontiguousarray(px[..., :3])) if (px[..., 3] == 255).all() and (k & 1) else Image.fromarray(px)
            buf = io.BytesIO()
            img.save(buf, format="PNG")
            png = buf.getvalue()
            if k % 3 == 0 or images_in_views:
                doc["images"].append({"bufferView": add_view(png), "mimeType": "image/png"})
            elif embed:
                doc["images"].append({"uri": "data:image/png;base64," + base64.b64encode(png).decode()})
            else:
                name = "%s_tex%d.png" % (os.path.splitext(os.path.basename(path))[0], k)
                with open(os.path.join(os.path.dirname(path), name), "wb") as f:
                    f.write(png)
                doc["images"].append({"uri": name})
            nearest = int(t.get("filter", 0)) == 1
            doc["samplers"].append({"wrapS": wrap[int(t.get("wrap_u", 0))], "wrapT": wrap[int(t.get("wrap_v", 0))],
                                    "magFilter": 9728 if nearest else 9729, "minFilter": 9728 if nearest else 9987})
            doc["textures"].append({"source": k, "sampler": k})
    if embed:
        doc["buffers"] = [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(bytes(blob)).decode()}]
    else:
        bin_name = os.path.splitext(os.path.basename(path))[0] + ".bin"
        with open(os.path.join(os.path.dirname(path), bin_name), "wb") as f:
            f.write(bytes(blob))
        doc["buffers"] = [{"byteLength": len(blob), "uri": bin_name}]
    with open(path, "w") as f:
        json.dump(doc, f)
    return path


def write_glb(scene, path):
    """Same content as write_gltf in the binary .glb container (JSON chunk + BIN chunk)."""
    import json
    import os
    import struct
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        tmp = write_gltf(scene, os.path.join(d, "x.gltf"), embed=False, images_in_views=True)
        doc = json.load(open(tmp))
        blob = open(os.path.join(d, "x.bin"), "rb").read()
    del doc["buffers"][0]["uri"]
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    blob += b"\0" * (-len(blob) % 4)
    with open(path, "wb") as f:
        f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(blob)))
        f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        f.write(struct.pack("<II", len(blob), 0x004E4942) + blob)
    return path

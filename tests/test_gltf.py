"""glTF ingest (vlb_scene_load_gltf / vlb_gltf_probe), the replacement of the reference's tinygltf
loader (src/scene_manager.cpp:32-67, 257-337, 463-538, 837-871). CPU tests use the host-only probe;
GPU tests check that a scene loaded from a file traces and bakes bit-identically to the same scene
passed as arrays."""
import json
import math
import os

import numpy as np
import pytest

REF_CUBE = "/root/reference/src/vendor/default_blender_cube.gltf"


def test_probe_counts_and_bounds_roundtrip(vlb, scenes, tmp_path):
    sc = scenes.small_room()
    for name, writer in (("a.gltf", scenes.write_gltf), ("b.glb", scenes.write_glb)):
        p = writer(sc, str(tmp_path / name))
        counts, bounds = vlb.gltf_probe(p)
        # the writer emits one mesh per instance, so shared meshes are duplicated in the file
        assert counts["vertices"] == int(sc["instances"]["vertex_count"].sum())
        assert counts["indices"] == int(sc["instances"]["index_count"].sum())
        assert counts["instances"] == len(sc["instances"])
        assert counts["materials"] == len(sc["materials"]) + 1          # + trailing default (scene_manager.cpp:851)
        assert counts["triangles"] == sum(int(i["index_count"]) // 3 for i in sc["instances"])
        assert np.all(bounds[:3] <= 0) and np.all(bounds[3:] >= 0)       # reference bounds start at the origin
    p = scenes.write_gltf(sc, str(tmp_path / "ext.gltf"), embed=False)  # external .bin buffer
    assert vlb.gltf_probe(p)[0]["triangles"] == counts["triangles"]


@pytest.mark.skipif(not os.path.exists(REF_CUBE), reason="reference fixture not present on this box")
def test_reference_default_cube_kat(vlb):
    # 24 vertices, 36 indices (12 triangles), 1 material (+ default), bounds [(-1,-1,-1),(1,1,1)]
    counts, bounds = vlb.gltf_probe(REF_CUBE)
    assert counts == {"vertices": 24, "indices": 36, "instances": 1, "materials": 2, "triangles": 12}
    assert np.allclose(bounds, [-1, -1, -1, 1, 1, 1])
    s = vlb.default_settings()
    vlb.settings_from_bounds(s, bounds)
    assert np.allclose(list(s.step), [2.0 / 6.0] * 3)                   # gridStep of the 7x7x7 grid


def _write(tmp_path, doc, name="t.gltf"):
    p = str(tmp_path / name)
    json.dump(doc, open(p, "w"))
    return p


def _tri_doc(nodes, extra=None):
    import base64
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    idx = np.array([0, 1, 2], np.uint8)
    blob = pos.tobytes() + idx.tobytes() + b"\0"
    doc = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0]}], "nodes": nodes,
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                         {"bufferView": 1, "componentType": 5121, "count": 3, "type": "SCALAR"}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 3}],
           "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    if extra:
        doc.update(extra)
    return doc


def test_reference_bounds_quirk_uses_local_matrix_only(vlb, tmp_path):
    # parent translates by +10 in x, child (holding the mesh) by +1: the reference applies only the
    # child's LOCAL matrix to the two AABB corners (scene_manager.cpp:497-507) -> max x = 2, not 12
    nodes = [{"translation": [10, 0, 0], "children": [1]}, {"translation": [1, 0, 0], "mesh": 0}]
    counts, bounds = vlb.gltf_probe(_write(tmp_path, _tri_doc(nodes)))
    assert counts["triangles"] == 1 and counts["instances"] == 1
    assert np.allclose(bounds, [0, 0, 0, 2, 1, 0])


def test_unsupported_and_malformed_inputs_fail_loudly(vlb, tmp_path):
    doc = _tri_doc([{"mesh": 0}])
    doc["meshes"][0]["primitives"][0]["mode"] = 1                       # LINES
    with pytest.raises(vlb.VlbError) as e:
        vlb.gltf_probe(_write(tmp_path, doc))
    assert e.value.code == vlb.ERR_UNSUPPORTED
    doc = _tri_doc([{"mesh": 0}])
    doc["accessors"][1]["componentType"] = 5126                         # float indices: the reference throws too
    with pytest.raises(vlb.VlbError):
        vlb.gltf_probe(_write(tmp_path, doc))
    doc = _tri_doc([{"mesh": 0}])
    doc["accessors"][0]["count"] = 300                                  # exceeds the buffer
    with pytest.raises(vlb.VlbError) as e:
        vlb.gltf_probe(_write(tmp_path, doc))
    assert e.value.code == vlb.ERR_IO
    with pytest.raises(vlb.VlbError):
        vlb.gltf_probe(str(tmp_path / "missing.gltf"))
    open(str(tmp_path / "bad.gltf"), "w").write("{not json")
    with pytest.raises(vlb.VlbError):
        vlb.gltf_probe(str(tmp_path / "bad.gltf"))


def test_crafted_accessor_numbers_are_rejected_not_wrapped(vlb, tmp_path):
    # numbers straight from the file must be validated without wrapping arithmetic: count = 2^60 with byteStride 16 and
    # byteOffset 4 makes off + (count - 1) * stride + elem wrap to a small value (this used to pass the check and read far
    # out of bounds); negative and fractional numbers, and a stride smaller than the element, are refused as well
    for patch in ({"count": 2 ** 60}, {"count": -3}, {"count": 2.5}, {"byteOffset": -4}, {"byteOffset": 2 ** 62}):
        doc = _tri_doc([{"mesh": 0}])
        doc["bufferViews"][0]["byteStride"] = 16
        doc["accessors"][0]["byteOffset"] = 4
        doc["accessors"][0].update(patch)
        with pytest.raises(vlb.VlbError):
            vlb.gltf_probe(_write(tmp_path, doc))
    doc = _tri_doc([{"mesh": 0}])
    doc["bufferViews"][0]["byteStride"] = 8                             # smaller than a float VEC3
    with pytest.raises(vlb.VlbError):
        vlb.gltf_probe(_write(tmp_path, doc))
    doc = _tri_doc([{"mesh": 0}])
    doc["bufferViews"][0]["byteOffset"] = 2 ** 63                       # view offset + accessor offset would wrap
    with pytest.raises(vlb.VlbError):
        vlb.gltf_probe(_write(tmp_path, doc))


def test_non_float_normals_and_uvs_are_refused_not_dropped(vlb, tmp_path):
    import base64
    for attr, comp, typ in (("TEXCOORD_0", 5123, "VEC2"), ("TEXCOORD_0", 5126, "SCALAR"), ("NORMAL", 5120, "VEC3")):
        doc = _tri_doc([{"mesh": 0}])
        extra = np.zeros(3 * 16, np.uint8).tobytes()
        blob = base64.b64decode(doc["buffers"][0]["uri"].split(",", 1)[1]) + extra
        doc["buffers"][0] = {"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}
        doc["bufferViews"].append({"buffer": 0, "byteOffset": 40, "byteLength": 48})
        doc["accessors"].append({"bufferView": 2, "componentType": comp, "count": 3, "type": typ})
        doc["meshes"][0]["primitives"][0]["attributes"][attr] = 2
        with pytest.raises(vlb.VlbError) as e:
            vlb.gltf_probe(_write(tmp_path, doc))
        assert e.value.code == vlb.ERR_UNSUPPORTED


# ------------------------------------------------------------------------------- GPU ------------
@pytest.mark.gpu
@pytest.mark.parametrize("container", ["gltf", "glb"])
def test_loaded_scene_equals_array_scene(ctx, vlb, scenes, tmp_path, container):
    sc = scenes.small_room()
    p = (scenes.write_gltf if container == "gltf" else scenes.write_glb)(sc, str(tmp_path / ("room." + container)))
    rng = np.random.default_rng(3)
    o = rng.uniform(0.2, 3.8, (5000, 3)).astype(np.float32)
    d = rng.normal(size=(5000, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    ctx.set_scene(sc)
    ids_a, tuv_a = ctx.trace_rays(o, d)
    tight_a = ctx.scene_bounds(tight=True)
    s = vlb.default_settings()
    s.probes[:] = (2, 2, 2); s.dir_w, s.dir_h = 16, 8; s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SRGB_ENCODE
    vlb.settings_from_bounds(s, tight_a)
    bake_a = ctx.bake_probes(s)
    ctx.load_gltf(p)
    ids_b, tuv_b = ctx.trace_rays(o, d)
    assert np.array_equal(ids_a, ids_b) and np.array_equal(tuv_a, tuv_b)
    assert np.array_equal(tight_a, ctx.scene_bounds(tight=True))
    # normals are re-normalised by the loader (glm::normalize, scene_manager.cpp:272): last-ulp differences
    from conftest import rel_l2
    assert rel_l2(ctx.bake_probes(s), bake_a) <= 1e-5
    assert (ids_a >= 0).mean() > 0.5


@pytest.mark.gpu
def test_node_trs_hierarchy(ctx, vlb, tmp_path):
    # child TRS under a rotated + scaled parent: world = parent * child (Node_t::getMatrix), M = T*R*S
    a = math.radians(90.0)
    q = [0.0, math.sin(a / 2), 0.0, math.cos(a / 2)]                    # 90 degrees about +y
    nodes = [{"rotation": q, "scale": [2, 2, 2], "children": [1]}, {"translation": [1, 0, 0], "mesh": 0}]
    ctx.load_gltf(_write(tmp_path, _tri_doc(nodes)))
    b = ctx.scene_bounds(tight=True)
    # triangle (0,0,0),(1,0,0),(0,1,0) -> +1 in x -> scaled by 2 -> rotated +90 about y: (x,y,z) -> (z, y, -x)
    assert np.allclose(b, [0, 0, -4, 0, 2, -2], atol=1e-5)


@pytest.mark.gpu
def test_load_errors_keep_ctx_usable(ctx, vlb, scenes, tmp_path):
    with pytest.raises(vlb.VlbError):
        ctx.load_gltf(str(tmp_path / "nope.gltf"))
    ctx.set_scene(scenes.small_room())
    assert ctx.build_bvh().n_triangles > 0

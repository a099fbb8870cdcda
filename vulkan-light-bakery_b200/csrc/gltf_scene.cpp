// gltf_scene.cpp — glTF 2.0 scene ingest for the bake path; replaces the tinygltf-based loader of
// the reference (SceneManager::pushScene -> Scene_t, src/scene_manager.cpp:32-67, 257-337, 463-538,
// 695-857, 1013-1034) for exactly what the bake consumes: per-primitive vertex / index arrays in the
// reference's shader::Vertex layout, one instance per (node, primitive) in loadNode's pre-order with
// the node's world matrix, and the material table with its trailing default material.
// Host-side, off the timed path (O(file size)); everything downstream runs on the GPU.
//
// Handled: .gltf (JSON) and .glb containers; buffers as base64 data URIs, external files next to
// the .gltf, or the GLB BIN chunk; node TRS / matrix hierarchies; POSITION / NORMAL / TEXCOORD_0/1
// float attributes with arbitrary byteStride; u8 / u16 / u32 indices; baseColorFactor and the other
// factor fields of shader::Factors. Not handled (fail with VLB_ERR_UNSUPPORTED, never silently):
// sparse accessors, non-triangle primitive modes, Draco / meshopt compression. Textures that a material
// names as baseColorTexture are decoded (PNG: png_decode.cpp, baseline JPEG: jpeg_decode.cpp) with their samplers (loadTextures /
// loadSamplers, :941-973, 650-690); the other texture slots are recorded in the material table only, as the
// bake shader never samples them (env_map.rchit:36-49).
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "vlb_context.h"
#include "vlb_json.h"

namespace vlb {

namespace {

struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };
}  // namespace
void png_decode_rgba8(const uint8_t* data, size_t size, std::vector<uint8_t>& rgba, int& width, int& height, bool& unsupported);
void jpeg_decode_rgba8(const uint8_t* data, size_t size, std::vector<uint8_t>& rgba, int& width, int& height, bool& unsupported);
void hdr_decode_rgba32f(const uint8_t* data, size_t size, std::vector<float>& rgba, int& width, int& height, bool& unsupported);
namespace {

bool read_all(const std::string& path, std::string& out) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    std::ostringstream ss; ss << f.rdbuf();
    out = ss.str();
    return true;
}

// ---- column-major 4x4 float matrices with glm's operation order ----------------------------
struct M4 { float c[4][4]; };   // c[col][row]
M4 identity() { M4 m{}; for (int i = 0; i < 4; ++i) m.c[i][i] = 1.f; return m; }
M4 mul(const M4& a, const M4& b) {   // glm: Result[j] = A[0]*B[j][0] + A[1]*B[j][1] + A[2]*B[j][2] + A[3]*B[j][3]
    M4 r{};
    for (int j = 0; j < 4; ++j)
        for (int row = 0; row < 4; ++row)
            r.c[j][row] = ((a.c[0][row] * b.c[j][0] + a.c[1][row] * b.c[j][1]) + a.c[2][row] * b.c[j][2]) + a.c[3][row] * b.c[j][3];
    return r;
}

// Scene_t::loadMatrix (src/scene_manager.cpp:463-477): M = translate * rotation * scale * matrix, the
// factors narrowed from the file's doubles to float first.
M4 node_matrix(const Json& node) {
    float t[3] = {0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f};
    double q[4] = {0.0, 0.0, 0.0, 1.0};   // x y z w
    M4 xf = identity();
    if (const Json* a = node.find("translation")) if (a->a.size() == 3) for (int i = 0; i < 3; ++i) t[i] = (float)a->a[i].num();
    if (const Json* a = node.find("scale")) if (a->a.size() == 3) for (int i = 0; i < 3; ++i) s[i] = (float)a->a[i].num();
    if (const Json* a = node.find("rotation")) if (a->a.size() == 4) for (int i = 0; i < 4; ++i) q[i] = a->a[i].num();
    if (const Json* a = node.find("matrix")) if (a->a.size() == 16) for (int i = 0; i < 16; ++i) xf.c[i / 4][i % 4] = (float)a->a[i].num();
    // glm::mat4_cast of a dquat, then narrowed
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    M4 rot = identity();
    rot.c[0][0] = (float)(1.0 - 2.0 * (y * y + z * z)); rot.c[0][1] = (float)(2.0 * (x * y + w * z)); rot.c[0][2] = (float)(2.0 * (x * z - w * y));
    rot.c[1][0] = (float)(2.0 * (x * y - w * z)); rot.c[1][1] = (float)(1.0 - 2.0 * (x * x + z * z)); rot.c[1][2] = (float)(2.0 * (y * z + w * x));
    rot.c[2][0] = (float)(2.0 * (x * z + w * y)); rot.c[2][1] = (float)(2.0 * (y * z - w * x)); rot.c[2][2] = (float)(1.0 - 2.0 * (x * x + y * y));
    M4 m = identity();
    for (int row = 0; row < 4; ++row)   // glm::translate: m[3] = m[0]*t.x + m[1]*t.y + m[2]*t.z + m[3]
        m.c[3][row] = ((m.c[0][row] * t[0] + m.c[1][row] * t[1]) + m.c[2][row] * t[2]) + m.c[3][row];
    m = mul(m, rot);
    for (int k = 0; k < 3; ++k) for (int row = 0; row < 4; ++row) m.c[k][row] *= s[k];   // glm::scale
    return mul(m, xf);
}

struct Accessor {
    const uint8_t* base = nullptr;   // first element
    size_t stride = 0, count = 0;
    int component = 0;               // 5120.. glTF component types
    int ncomp = 0;
};

struct Document {
    Json json;
    std::vector<std::vector<uint8_t>> buffers;

    const Json& arr(const char* key) const {
        static const Json empty = Json::array();
        const Json* a = json.find(key);
        return a && a->is(Json::Array) ? *a : empty;
    }
    Accessor accessor(long long index) const {
        const Json& accs = arr("accessors");
        if (index < 0 || (size_t)index >= accs.a.size()) throw std::runtime_error("glTF: accessor index out of range");
        const Json& a = accs.a[(size_t)index];
        if (a.find("sparse")) throw Unsupported("glTF: sparse accessors are not supported");
        const Json* bvi = a.find("bufferView");
        if (!bvi) throw Unsupported("glTF: accessor without bufferView (zero-filled) is not supported");
        const Json& views = arr("bufferViews");
        if (bvi->integer_value() < 0 || (size_t)bvi->integer_value() >= views.a.size()) throw std::runtime_error("glTF: bufferView index out of range");
        const Json& v = views.a[(size_t)bvi->integer_value()];
        const Json* bi = v.find("buffer");
        if (!bi || bi->integer_value() < 0 || (size_t)bi->integer_value() >= buffers.size()) throw std::runtime_error("glTF: buffer index out of range");
        const std::vector<uint8_t>& buf = buffers[(size_t)bi->integer_value()];
        Accessor r;
        // every number below comes straight from the file: reject negatives and validate without wrapping arithmetic
        auto field = [](const Json& o, const char* key) -> size_t {
            const Json* f = o.find(key);
            if (!f) return 0;
            const double d = f->num();
            if (!(d >= 0.0) || d > 9.0e15 || d != std::floor(d)) throw std::runtime_error(std::string("glTF: ") + key + " is not a non-negative integer");
            return (size_t)d;
        };
        r.component = (int)field(a, "componentType");
        if (r.component < 5120 || r.component > 5126 || r.component == 5124) throw Unsupported("glTF: accessor componentType is not one glTF 2.0 defines");
        r.count = field(a, "count");
        const std::string type = a.find("type") ? a.find("type")->s : "";
        r.ncomp = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : 0;
        if (!r.ncomp) throw Unsupported("glTF: accessor type " + type + " is not used by the bake path");
        const size_t csize = (r.component == 5120 || r.component == 5121) ? 1 : (r.component == 5122 || r.component == 5123) ? 2 : 4;
        const size_t elem = csize * r.ncomp;
        const size_t off_a = field(a, "byteOffset"), off_v = field(v, "byteOffset");
        const size_t bstride = field(v, "byteStride");
        r.stride = bstride ? bstride : elem;
        if (r.stride < elem) throw std::runtime_error("glTF: byteStride smaller than the accessor's element");
        if (off_a > buf.size() || off_v > buf.size() - off_a) throw std::runtime_error("glTF: accessor exceeds its buffer");
        const size_t off = off_a + off_v;
        if (r.count) {
            if (elem > buf.size() - off || r.count - 1 > (buf.size() - off - elem) / r.stride) throw std::runtime_error("glTF: accessor exceeds its buffer");
        }
        r.base = buf.data() + off;
        return r;
    }
};

std::string dir_of(const std::string& path) {
    const size_t p = path.find_last_of("/\\");
    return p == std::string::npos ? std::string() : path.substr(0, p + 1);
}

void load_document(const std::string& path, Document& doc) {
    std::string text;
    if (!read_all(path, text)) throw std::runtime_error("cannot read " + path);
    std::vector<uint8_t> glb_bin;
    bool have_bin = false;
    if (text.size() >= 12 && std::memcmp(text.data(), "glTF", 4) == 0) {   // .glb container (LoadBinaryFromFile, :49-51)
        size_t pos = 12;
        std::string json_text;
        while (pos + 8 <= text.size()) {
            uint32_t len, type;
            std::memcpy(&len, text.data() + pos, 4); std::memcpy(&type, text.data() + pos + 4, 4);
            pos += 8;
            if (pos + len > text.size()) throw std::runtime_error("glb: truncated chunk");
            if (type == 0x4E4F534Au) json_text.assign(text.data() + pos, len);                       // "JSON"
            else if (type == 0x004E4942u && !have_bin) { glb_bin.assign(text.begin() + pos, text.begin() + pos + len); have_bin = true; }   // "BIN\0"
            pos += (len + 3u) & ~size_t(3);
        }
        doc.json = json_parse(json_text);
    } else {
        doc.json = json_parse(text);
    }
    if (!doc.json.is(Json::Object)) throw std::runtime_error("glTF: root is not an object");
    const std::string base = dir_of(path);
    for (const Json& b : doc.arr("buffers").a) {
        const Json* uri = b.find("uri");
        std::vector<uint8_t> data;
        if (!uri) {
            if (!have_bin) throw std::runtime_error("glTF: buffer without uri outside a .glb");
            data = glb_bin;
        } else if (uri->s.compare(0, 5, "data:") == 0) {
            const size_t comma = uri->s.find(',');
            if (comma == std::string::npos || uri->s.find(";base64") == std::string::npos) throw Unsupported("glTF: data URI is not base64");
            data = base64_decode(uri->s.substr(comma + 1));
        } else {
            std::string raw;
            if (!read_all(base + uri->s, raw)) throw std::runtime_error("glTF: cannot read buffer file " + base + uri->s);
            data.assign(raw.begin(), raw.end());
        }
        const size_t want = (size_t)(b.find("byteLength") ? b.find("byteLength")->integer_value() : 0);
        if (data.size() < want) throw std::runtime_error("glTF: buffer shorter than its byteLength");
        doc.buffers.push_back(std::move(data));
    }
}

float read_float(const Accessor& a, size_t i, int c) {
    float v; std::memcpy(&v, a.base + i * a.stride + 4 * (size_t)c, 4); return v;
}

void set_texture(vlb_texture_ref& t, const Json* info) {
    t.index = -1; t.coord_set = 0;
    if (!info) return;
    if (const Json* i = info->find("index")) t.index = (int32_t)i->integer_value();
    if (const Json* c = info->find("texCoord")) t.coord_set = (int32_t)c->integer_value();
}

// loadMaterials / loadFactors / matchTextures (src/scene_manager.cpp:695-857)
void load_materials(const Document& doc, std::vector<vlb_material>& out) {
    auto blank = [] {
        vlb_material m;
        std::memset(&m, 0, sizeof m);
        vlb_texture_ref* t = &m.normal;
        for (int k = 0; k < 8; ++k) { t[k].index = -1; t[k].coord_set = 0; }
        return m;
    };
    for (const Json& jm : doc.arr("materials").a) {
        vlb_material m = blank();
        const Json* pbr = jm.find("pbrMetallicRoughness");
        if (const Json* am = jm.find("alphaMode")) if (am->s == "MASK") m.alpha_cutoff = 0.5f;
        if (const Json* ac = jm.find("alphaCutoff")) m.alpha_cutoff = (float)ac->num();
        if (pbr) {
            if (const Json* f = pbr->find("baseColorFactor")) if (f->a.size() == 4) for (int i = 0; i < 4; ++i) m.base_color_factor[i] = (float)f->a[i].num();
            if (const Json* f = pbr->find("metallicFactor")) m.metallic = (float)f->num();
            if (const Json* f = pbr->find("roughnessFactor")) m.roughness = (float)f->num();
            set_texture(m.base_color, pbr->find("baseColorTexture"));
            set_texture(m.metallic_roughness, pbr->find("metallicRoughnessTexture"));
        }
        if (const Json* f = jm.find("emissiveFactor")) if (f->a.size() == 3) { for (int i = 0; i < 3; ++i) m.emissive_factor[i] = (float)f->a[i].num(); m.emissive_factor[3] = 1.f; }
        set_texture(m.normal, jm.find("normalTexture"));
        set_texture(m.occlusion, jm.find("occlusionTexture"));
        set_texture(m.emissive, jm.find("emissiveTexture"));
        if (const Json* ext = jm.find("extensions")) if (const Json* sg = ext->find("KHR_materials_pbrSpecularGlossiness")) {
            if (const Json* f = sg->find("diffuseFactor")) for (size_t i = 0; i < f->a.size() && i < 4; ++i) m.diffuse_factor[i] = (float)f->a[i].num();
            if (const Json* f = sg->find("specularFactor")) for (size_t i = 0; i < f->a.size() && i < 3; ++i) m.specular_factor[i] = (float)f->a[i].num();
            set_texture(m.diffuse_ext, sg->find("diffuseTexture"));
            set_texture(m.specular_ext, sg->find("specularGlossinessTexture"));
        }
        out.push_back(m);
    }
    out.push_back(blank());   // the trailing default material (:851); primitives without a material use it (:510)
}

struct HostTexture { std::vector<uint8_t> rgba; int width = 1, height = 1, wrap_u = 0, wrap_v = 0, filter = 0; bool used = false; };

struct HostScene {
    std::vector<vlb_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<vlb_instance> instances;
    std::vector<vlb_material> materials;
    std::vector<HostTexture> textures;          // one per glTF texture; only those used as baseColor are decoded
    float ref_bounds[6] = {0, 0, 0, 0, 0, 0};   // Scene_t::bounds starts at the origin (scene_manager.hpp:170)
};

// Scene_t::loadNode (src/scene_manager.cpp:479-538), pre-order over scenes[0]
void load_node(const Document& doc, long long index, const M4& parent_world, HostScene& hs, int depth) {
    const Json& nodes = doc.arr("nodes");
    if (index < 0 || (size_t)index >= nodes.a.size()) throw std::runtime_error("glTF: node index out of range");
    if (depth > 256) throw std::runtime_error("glTF: node hierarchy too deep (cycle?)");
    const Json& node = nodes.a[(size_t)index];
    const M4 local = node_matrix(node);
    const M4 world = mul(parent_world, local);   // Node_t::getMatrix: p->matrix * matrix up the chain (:445-461)
    const Json* mesh_i = node.find("mesh");
    if (mesh_i && mesh_i->integer_value() >= 0) {
        const Json& meshes = doc.arr("meshes");
        if ((size_t)mesh_i->integer_value() >= meshes.a.size()) throw std::runtime_error("glTF: mesh index out of range");
        const Json* prims = meshes.a[(size_t)mesh_i->integer_value()].find("primitives");
        for (const Json& prim : (prims ? prims->a : std::vector<Json>())) {
            if (const Json* mode = prim.find("mode")) if (mode->integer_value() != 4) throw Unsupported("glTF: only TRIANGLES primitives are supported");
            if (const Json* ext = prim.find("extensions")) if (ext->find("KHR_draco_mesh_compression")) throw Unsupported("glTF: Draco compression is not supported");
            const Json* attrs = prim.find("attributes");
            const Json* pos_i = attrs ? attrs->find("POSITION") : nullptr;
            if (!pos_i) throw std::runtime_error("glTF: primitive without POSITION");
            const Accessor pos = doc.accessor(pos_i->integer_value());
            if (pos.component != 5126 || pos.ncomp != 3) throw Unsupported("glTF: POSITION must be float VEC3");
            Accessor nrm, uv0, uv1;
            if (const Json* a = attrs->find("NORMAL")) nrm = doc.accessor(a->integer_value());
            if (const Json* a = attrs->find("TEXCOORD_0")) uv0 = doc.accessor(a->integer_value());
            if (const Json* a = attrs->find("TEXCOORD_1")) uv1 = doc.accessor(a->integer_value());
            // only float attributes are read (fetchVertices, :257-290); anything else is refused, never dropped silently
            if (nrm.base && (nrm.component != 5126 || nrm.ncomp != 3)) throw Unsupported("glTF: NORMAL must be a float VEC3 accessor");
            for (const Accessor* uv : {&uv0, &uv1})
                if (uv->base && (uv->component != 5126 || uv->ncomp < 2)) throw Unsupported("glTF: TEXCOORD_n must be a float VEC2 accessor (normalised integer texture coordinates are not supported)");
            vlb_instance inst;
            std::memset(&inst, 0, sizeof inst);
            inst.first_vertex = (uint32_t)hs.vertices.size();
            inst.vertex_count = (uint32_t)pos.count;
            float lo[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, hi[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
            for (size_t v = 0; v < pos.count; ++v) {   // fetchVertices (:257-290)
                vlb_vertex vx;
                std::memset(&vx, 0, sizeof vx);
                for (int c = 0; c < 3; ++c) vx.position[c] = read_float(pos, v, c);
                vx.position[3] = 1.0f;
                if (nrm.base && nrm.component == 5126 && nrm.ncomp == 3 && v < nrm.count) {
                    const float x = read_float(nrm, v, 0), y = read_float(nrm, v, 1), z = read_float(nrm, v, 2);
                    const float inv = 1.0f / std::sqrt(x * x + y * y + z * z);   // glm::normalize = v * inversesqrt(dot(v, v))
                    vx.normal[0] = x * inv; vx.normal[1] = y * inv; vx.normal[2] = z * inv;
                }
                if (uv0.base && uv0.component == 5126 && v < uv0.count) { vx.uv0[0] = read_float(uv0, v, 0); vx.uv0[1] = read_float(uv0, v, 1); }
                if (uv1.base && uv1.component == 5126 && v < uv1.count) { vx.uv1[0] = read_float(uv1, v, 0); vx.uv1[1] = read_float(uv1, v, 1); }
                for (int c = 0; c < 3; ++c) { lo[c] = std::fmin(lo[c], vx.position[c]); hi[c] = std::fmax(hi[c], vx.position[c]); }
                hs.vertices.push_back(vx);
            }
            inst.first_index = (uint32_t)hs.indices.size();
            const Json* idx_i = prim.find("indices");
            if (idx_i && idx_i->integer_value() >= 0) {   // fetchIndices (:292-337)
                const Accessor ia = doc.accessor(idx_i->integer_value());
                for (size_t i = 0; i < ia.count; ++i) {
                    const uint8_t* p = ia.base + i * ia.stride;
                    uint32_t v;
                    if (ia.component == 5125) std::memcpy(&v, p, 4);
                    else if (ia.component == 5123) { uint16_t t; std::memcpy(&t, p, 2); v = t; }
                    else if (ia.component == 5121) v = *p;
                    else throw Unsupported("glTF: index component type not supported");   // the reference throws too (:332)
                    if (v >= pos.count) throw std::runtime_error("glTF: index out of range");
                    hs.indices.push_back(v);
                }
                inst.index_count = (uint32_t)ia.count;
            } else {   // non-indexed primitive: the reference dereferences accessors[-1]; defined here as 0..n-1
                for (size_t i = 0; i < pos.count; ++i) hs.indices.push_back((uint32_t)i);
                inst.index_count = (uint32_t)pos.count;
            }
            const Json* mat = prim.find("material");
            inst.material_index = (mat && mat->integer_value() >= 0) ? (uint32_t)mat->integer_value() : (uint32_t)hs.materials.size() - 1;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) inst.transform[4 * r + c] = world.c[c][r];   // transpose -> 3x4 row-major
            hs.instances.push_back(inst);
            // Scene bounds quirk (:497-507): only the node's LOCAL matrix is applied, to the two AABB corners
            if (pos.count) {
                for (int c = 0; c < 3; ++c) {
                    const float a = ((local.c[0][c] * lo[0] + local.c[1][c] * lo[1]) + local.c[2][c] * lo[2]) + local.c[3][c] * 1.0f;
                    const float b = ((local.c[0][c] * hi[0] + local.c[1][c] * hi[1]) + local.c[2][c] * hi[2]) + local.c[3][c] * 1.0f;
                    hs.ref_bounds[c] = std::fmin(hs.ref_bounds[c], a);
                    hs.ref_bounds[3 + c] = std::fmax(hs.ref_bounds[3 + c], b);
                }
            }
        }
    }
    if (const Json* ch = node.find("children")) for (const Json& c : ch->a) load_node(doc, c.integer_value(), world, hs, depth + 1);
}

// Bytes of glTF image `index`: data URI, file next to the .gltf, or a bufferView (images in a .glb).
std::vector<uint8_t> image_bytes(const Document& doc, const std::string& base_dir, long long index) {
    const Json& images = doc.arr("images");
    if (index < 0 || (size_t)index >= images.a.size()) throw std::runtime_error("glTF: image index out of range");
    const Json& img = images.a[(size_t)index];
    if (const Json* uri = img.find("uri")) {
        if (uri->s.compare(0, 5, "data:") == 0) {
            const size_t comma = uri->s.find(',');
            if (comma == std::string::npos || uri->s.find(";base64") == std::string::npos) throw Unsupported("glTF: image data URI is not base64");
            return base64_decode(uri->s.substr(comma + 1));
        }
        std::string raw;
        if (!read_all(base_dir + uri->s, raw)) throw std::runtime_error("glTF: cannot read image file " + base_dir + uri->s);
        return std::vector<uint8_t>(raw.begin(), raw.end());
    }
    const Json* bvi = img.find("bufferView");
    const Json& views = doc.arr("bufferViews");
    if (!bvi || bvi->integer_value() < 0 || (size_t)bvi->integer_value() >= views.a.size()) throw std::runtime_error("glTF: image without uri or bufferView");
    const Json& v = views.a[(size_t)bvi->integer_value()];
    const Json* bi = v.find("buffer");
    if (!bi || bi->integer_value() < 0 || (size_t)bi->integer_value() >= doc.buffers.size()) throw std::runtime_error("glTF: buffer index out of range");
    const std::vector<uint8_t>& buf = doc.buffers[(size_t)bi->integer_value()];
    const size_t off = (size_t)(v.find("byteOffset") ? v.find("byteOffset")->integer_value() : 0);
    const size_t len = (size_t)(v.find("byteLength") ? v.find("byteLength")->integer_value() : 0);
    if (off + len > buf.size()) throw std::runtime_error("glTF: image bufferView exceeds its buffer");
    return std::vector<uint8_t>(buf.begin() + off, buf.begin() + off + len);
}

// Scene_t::loadSamplers + loadTextures (src/scene_manager.cpp:650-690, 941-973). Texture indices stay those of
// the glTF `textures` array (material.textures.baseColor.index, matchTextures :785-835); entries that no
// material uses as baseColor stay 1x1 white placeholders, like the reference's dummy texture (:964-970).
void load_textures(const Document& doc, const std::string& base_dir, HostScene& hs) {
    const Json& textures = doc.arr("textures");
    const Json& samplers = doc.arr("samplers");
    hs.textures.assign(textures.a.size(), HostTexture());
    for (HostTexture& t : hs.textures) t.rgba.assign(4, 255);
    std::vector<char> used(textures.a.size(), 0);
    for (const vlb_material& m : hs.materials) {
        if (m.base_color.index < 0) continue;
        if ((size_t)m.base_color.index >= textures.a.size()) throw std::runtime_error("glTF: baseColorTexture index out of range");
        used[(size_t)m.base_color.index] = 1;
    }
    for (size_t i = 0; i < textures.a.size(); ++i) {
        if (!used[i]) continue;
        const Json& jt = textures.a[i];
        HostTexture& t = hs.textures[i];
        t.used = true;
        const Json* si = jt.find("sampler");
        if (si && si->integer_value() >= 0) {
            if ((size_t)si->integer_value() >= samplers.a.size()) throw std::runtime_error("glTF: sampler index out of range");
            const Json& js = samplers.a[(size_t)si->integer_value()];
            auto wrap = [](const Json* w) {                              // toVkWrapMode (:654-668); glTF default 10497 = repeat
                const long long g = w ? w->integer_value() : 10497;
                return g == 33071 ? VLB_WRAP_CLAMP_TO_EDGE : (g == 33648 ? VLB_WRAP_MIRRORED_REPEAT : VLB_WRAP_REPEAT);
            };
            auto nearest = [](const Json* f) {                           // toVkFilterMode (:670-680)
                const long long g = f ? f->integer_value() : -1;
                return g == 9728 || g == 9984 || g == 9985;
            };
            t.wrap_u = wrap(js.find("wrapS")); t.wrap_v = wrap(js.find("wrapT"));
            // one filter per texture at the base level: the sample is a magnification or a minification depending on
            // footprint, which a ray-tracing stage does not have; magFilter decides (minFilter only if mag is absent)
            const Json* mag = js.find("magFilter");
            t.filter = nearest(mag ? mag : js.find("minFilter")) ? VLB_FILTER_NEAREST : VLB_FILTER_LINEAR;
        }
        const Json* src = jt.find("source");
        if (!src) throw Unsupported("glTF: texture without source (extension-only image) is not supported");
        const std::vector<uint8_t> bytes = image_bytes(doc, base_dir, src->integer_value());
        bool unsupported = false;
        try {
            if (bytes.size() >= 2 && bytes[0] == 0xFF && bytes[1] == 0xD8)
                jpeg_decode_rgba8(bytes.data(), bytes.size(), t.rgba, t.width, t.height, unsupported);
            else
                png_decode_rgba8(bytes.data(), bytes.size(), t.rgba, t.width, t.height, unsupported);
        } catch (const std::exception& e) {
            const std::string msg = "glTF: baseColor texture " + std::to_string(i) + ": " + e.what();
            if (unsupported) throw Unsupported(msg);
            throw std::runtime_error(msg);
        }
    }
}

void load_host_scene(const char* path, HostScene& hs) {
    Document doc;
    load_document(path, doc);
    load_materials(doc, hs.materials);
    load_textures(doc, dir_of(path), hs);
    // Scene_t::loadNodes (:860-871): the roots of scenes[0]
    const Json& scenes = doc.arr("scenes");
    if (scenes.a.empty()) throw std::runtime_error("glTF: no scenes (the reference reads model.scenes[0])");
    const Json* roots = scenes.a[0].find("nodes");
    for (const Json& r : (roots ? roots->a : std::vector<Json>())) load_node(doc, r.integer_value(), identity(), hs, 0);
}

int fail_thread(int code, const std::string& msg) { set_thread_error(msg.c_str()); return code; }

}  // namespace

}  // namespace vlb

using namespace vlb;

extern "C" {

int vlb_gltf_probe(const char* path, uint64_t counts[5], float ref_bounds[6]) {
    if (!path) return fail_thread(VLB_ERR_INVALID, "vlb_gltf_probe: NULL path");
    try {
        HostScene hs;
        load_host_scene(path, hs);
        if (counts) {
            uint64_t tris = 0;
            for (const vlb_instance& i : hs.instances) tris += i.index_count / 3;
            counts[0] = hs.vertices.size(); counts[1] = hs.indices.size(); counts[2] = hs.instances.size();
            counts[3] = hs.materials.size(); counts[4] = tris;
        }
        if (ref_bounds) std::memcpy(ref_bounds, hs.ref_bounds, sizeof hs.ref_bounds);
    } catch (const Unsupported& e) { return fail_thread(VLB_ERR_UNSUPPORTED, e.what());
    } catch (const std::exception& e) { return fail_thread(VLB_ERR_IO, e.what()); }
    return VLB_OK;
}

int vlb_image_load_rgba8(const char* path, void* texels, uint64_t capacity, int32_t size[2]) {
    if (!path || !size) return fail_thread(VLB_ERR_INVALID, "vlb_image_load_rgba8: NULL argument");
    std::string raw;
    if (!read_all(path, raw)) return fail_thread(VLB_ERR_IO, std::string("cannot read ") + path);
    const uint8_t* d = reinterpret_cast<const uint8_t*>(raw.data());
    std::vector<uint8_t> rgba;
    int w = 0, h = 0;
    bool unsupported = false;
    try {
        if (raw.size() >= 2 && d[0] == 0xFF && d[1] == 0xD8) jpeg_decode_rgba8(d, raw.size(), rgba, w, h, unsupported);
        else png_decode_rgba8(d, raw.size(), rgba, w, h, unsupported);
    } catch (const std::exception& e) {
        return fail_thread(unsupported ? VLB_ERR_UNSUPPORTED : VLB_ERR_IO, std::string(path) + ": " + e.what());
    }
    size[0] = w; size[1] = h;
    if (texels && capacity >= rgba.size()) std::memcpy(texels, rgba.data(), rgba.size());
    return VLB_OK;
}

int vlb_image_load_rgba32f(const char* path, void* texels, uint64_t capacity, int32_t size[2]) {
    if (!path || !size) return fail_thread(VLB_ERR_INVALID, "vlb_image_load_rgba32f: NULL argument");
    std::string raw;
    if (!read_all(path, raw)) return fail_thread(VLB_ERR_IO, std::string("cannot read ") + path);
    std::vector<float> rgba;
    int w = 0, h = 0;
    bool unsupported = false;
    try {
        hdr_decode_rgba32f(reinterpret_cast<const uint8_t*>(raw.data()), raw.size(), rgba, w, h, unsupported);
    } catch (const std::exception& e) {
        return fail_thread(unsupported ? VLB_ERR_UNSUPPORTED : VLB_ERR_IO, std::string(path) + ": " + e.what());
    }
    size[0] = w; size[1] = h;
    if (texels && capacity >= rgba.size() * sizeof(float)) std::memcpy(texels, rgba.data(), rgba.size() * sizeof(float));
    return VLB_OK;
}

int vlb_gltf_texture(const char* path, uint32_t index, void* texels, uint64_t capacity, int32_t info[6]) {
    if (!path || !info) return fail_thread(VLB_ERR_INVALID, "vlb_gltf_texture: NULL argument");
    try {
        HostScene hs;
        load_host_scene(path, hs);
        if (index >= hs.textures.size()) return fail_thread(VLB_ERR_INVALID, "vlb_gltf_texture: texture index out of range");
        const HostTexture& t = hs.textures[index];
        info[0] = t.width; info[1] = t.height; info[2] = t.wrap_u; info[3] = t.wrap_v; info[4] = t.filter; info[5] = t.used ? 1 : 0;
        if (texels && capacity >= t.rgba.size()) std::memcpy(texels, t.rgba.data(), t.rgba.size());
    } catch (const Unsupported& e) { return fail_thread(VLB_ERR_UNSUPPORTED, e.what());
    } catch (const std::exception& e) { return fail_thread(VLB_ERR_IO, e.what()); }
    return VLB_OK;
}

int vlb_scene_load_gltf(vlb_ctx* ctx, const char* path) {
    if (!ctx) return VLB_ERR_INVALID;
    if (!path) return ctx->fail(VLB_ERR_INVALID, "vlb_scene_load_gltf: NULL path");
    HostScene hs;
    try {
        load_host_scene(path, hs);
    } catch (const Unsupported& e) { return ctx->fail(VLB_ERR_UNSUPPORTED, "%s", e.what());
    } catch (const std::exception& e) { return ctx->fail(VLB_ERR_IO, "%s", e.what()); }
    const int r = vlb_scene_set_triangles(ctx, hs.vertices.data(), hs.vertices.size(), hs.indices.data(), hs.indices.size(),
                                          hs.instances.data(), (uint32_t)hs.instances.size(), hs.materials.data(),
                                          (uint32_t)hs.materials.size());
    if (r != VLB_OK) return r;
    std::memcpy(ctx->ref_bounds, hs.ref_bounds, sizeof hs.ref_bounds);   // local-matrix quirk, see load_node
    std::vector<vlb_texture> tex(hs.textures.size());
    for (size_t i = 0; i < tex.size(); ++i) {
        const HostTexture& t = hs.textures[i];
        tex[i].texels = t.rgba.data(); tex[i].width = t.width; tex[i].height = t.height;
        tex[i].wrap_u = t.wrap_u; tex[i].wrap_v = t.wrap_v; tex[i].filter = t.filter; tex[i].reserved = 0;
    }
    return vlb_scene_set_textures(ctx, tex.data(), (uint32_t)tex.size());
}

}  // extern "C"

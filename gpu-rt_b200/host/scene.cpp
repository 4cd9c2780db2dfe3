/*
 * scene.cpp — glTF 2.0 / GLB loader and Scene_Desc packing for the C ABI.
 *
 * Behavioural restatement of Scene::load / Scene::parse_mesh (src/scene/scene.cpp:48-391) on top
 * of this repo's own JSON reader instead of tinygltf: one Object per glTF primitive, ids from 1,
 * only node.matrix honoured (TRS ignored), pose recovered through
 * Mat4::decompose -> Euler -> Pose::transform, textures as RGBA8.
 */
#include <sys/stat.h>
#include "scene.h"

#include <cstdio>
#include <cstring>
#include <functional>
#include <zlib.h>

#include "json.h"

namespace gpurt {

void Mesh::set(std::vector<Vertex>&& v, std::vector<uint32_t>&& i) {
    verts = std::move(v);
    idx = std::move(i);
    BBox b;
    for(auto& p : verts) b.enclose(Vec3{p.pos[0], p.pos[1], p.pos[2]});
    bbox = b;
}

void Scene::clear() {
    objs.clear();
    insertion.clear();
    textures.clear();
}

unsigned int Scene::add(Object&& obj) {
    unsigned int id = obj.id;
    if(objs.emplace(std::make_pair(id, std::move(obj))).second) insertion.push_back(id);
    return id;
}

/* src/vk/rt.cpp:26-76 */
void Scene::build_desc(std::vector<SceneDesc>& descs, std::vector<SceneLight>& lights) const {
    descs.clear();
    lights.clear();
    std::unordered_map<unsigned int, unsigned int> obj_to_idx;
    for_objs([&](const Object& obj) {
        SceneDesc d;
        std::memset((void*)&d, 0, sizeof(d));
        d.index = (uint32_t)descs.size();
        d.model = obj.has_model ? obj.model : Mat4::scale(Vec3{scale}) * obj.pose.transform();
        d.modelIT = d.model.inverse().T();
        d.albedo_tex = obj.material.albedo_tex;
        d.metal_rough_tex = obj.material.metal_rough_tex;
        d.emissive_tex = obj.material.emissive_tex;
        d.normal_tex = obj.material.normal_tex;
        d.albedo[0] = obj.material.albedo.x, d.albedo[1] = obj.material.albedo.y,
        d.albedo[2] = obj.material.albedo.z;
        d.emissive[0] = obj.material.emissive.x, d.emissive[1] = obj.material.emissive.y,
        d.emissive[2] = obj.material.emissive.z;
        d.metal_rough[0] = obj.material.metal_rough.x, d.metal_rough[1] = obj.material.metal_rough.y;
        obj_to_idx[obj.id] = (unsigned int)descs.size();
        descs.push_back(d);
    });
    for_objs([&](const Object& obj) {
        if(obj.material.emissive != Vec3{} || obj.material.emissive_tex != -1) {
            SceneLight l;
            std::memset(&l, 0, sizeof(l));
            l.index = obj_to_idx[obj.id];
            l.n_triangles = (uint32_t)(obj.mesh.idx.size() / 3);
            BBox box = obj.mesh.bbox;
            box.transform(descs[l.index].model);
            l.bmin[0] = box.min.x, l.bmin[1] = box.min.y, l.bmin[2] = box.min.z;
            l.bmax[0] = box.max.x, l.bmax[1] = box.max.y, l.bmax[2] = box.max.z;
            lights.push_back(l);
        }
    });
}

/* ---------------------------------------------------------------------------------------------- */
namespace {

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if(!f) return false;
    struct stat st; /* a uri may name a directory (fopen succeeds, ftell answers LONG_MAX) or a device */
    if(fstat(fileno(f), &st) != 0 || !S_ISREG(st.st_mode)) {
        fclose(f);
        return false;
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = n > 0 ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == out.size();
}

bool base64(const std::string& in, std::vector<uint8_t>& out) {
    auto val = [](char c) -> int {
        if(c >= 'A' && c <= 'Z') return c - 'A';
        if(c >= 'a' && c <= 'z') return c - 'a' + 26;
        if(c >= '0' && c <= '9') return c - '0' + 52;
        if(c == '+' || c == '-') return 62;
        if(c == '/' || c == '_') return 63;
        return -1;
    };
    uint32_t acc = 0;
    int bits = 0;
    for(char c : in) {
        int v = val(c);
        if(v < 0) continue;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if(bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)((acc >> bits) & 0xFF));
        }
    }
    return true;
}

bool load_uri(const std::string& base, const std::string& uri, std::vector<uint8_t>& out) {
    if(uri.rfind("data:", 0) == 0) {
        size_t comma = uri.find(',');
        if(comma == std::string::npos) return false;
        return base64(uri.substr(comma + 1), out);
    }
    std::string dec;
    for(size_t i = 0; i < uri.size(); i++) { /* percent-decoding */
        if(uri[i] == '%' && i + 2 < uri.size()) {
            dec += (char)strtol(uri.substr(i + 1, 2).c_str(), nullptr, 16);
            i += 2;
        } else
            dec += uri[i];
    }
    return read_file(base + dec, out);
}

int comp_size(int ct) {
    switch(ct) {
    case 5120: case 5121: return 1;
    case 5122: case 5123: return 2;
    case 5124: case 5125: case 5126: return 4;
    case 5130: return 8;
    }
    return -1;
}
int type_comps(const std::string& t) {
    if(t == "SCALAR") return 1;
    if(t == "VEC2") return 2;
    if(t == "VEC3") return 3;
    if(t == "VEC4") return 4;
    if(t == "MAT2") return 4;
    if(t == "MAT3") return 9;
    if(t == "MAT4") return 16;
    return -1;
}

struct Accessor {
    const uint8_t* ptr = nullptr;
    size_t stride = 0, count = 0;
    int ctype = 0, ncomp = 0;
};

struct Gltf {
    Json root;
    std::vector<std::vector<uint8_t>> buffers;

    /* tinygltf Accessor::ByteStride: bufferView.byteStride, or tightly packed when 0 */
    bool accessor(int i, Accessor& a) const {
        const Json& acc = root["accessors"][(size_t)i];
        if(acc.type != Json::Obj) return false;
        const Json& bv = root["bufferViews"][(size_t)acc["bufferView"].integer(-1)];
        if(bv.type != Json::Obj) return false;
        size_t b = (size_t)bv["buffer"].integer(0);
        if(b >= buffers.size()) return false;
        a.ctype = acc["componentType"].integer(0);
        a.ncomp = type_comps(acc["type"].string());
        int cs = comp_size(a.ctype);
        if(cs < 0 || a.ncomp < 0) return false;
        /* untrusted numbers: finite, non-negative and no larger than the buffer before anything is multiplied */
        const double size = (double)buffers[b].size();
        auto sane = [&](double v) { return v == v && v >= 0.0 && v <= size; };
        const double d_count = acc["count"].number(0), d_stride = bv["byteStride"].number(0);
        const double d_off0 = bv["byteOffset"].number(0), d_off1 = acc["byteOffset"].number(0);
        if(!sane(d_count) || !sane(d_stride) || !sane(d_off0) || !sane(d_off1) || !sane(d_off0 + d_off1)) return false;
        a.count = (size_t)d_count;
        size_t bs = (size_t)d_stride;
        a.stride = bs ? bs : (size_t)cs * a.ncomp;
        size_t off = (size_t)d_off0 + (size_t)d_off1;
        const size_t elem = (size_t)cs * a.ncomp, avail = buffers[b].size() - off; /* off <= size */
        if(a.count && (elem > avail || (a.count - 1) > (avail - elem) / a.stride)) return false;
        a.ptr = buffers[b].data() + off;
        return true;
    }
};

template <typename T> T rd(const uint8_t* p) {
    T v;
    std::memcpy(&v, p, sizeof(T));
    return v;
}

/* float or double vector attribute -> floats (scene.cpp:157-270) */
bool read_vec(const Accessor& a, int want, std::vector<float>& out) {
    if(a.ncomp != want) return false;
    if(a.ctype != 5126 && a.ctype != 5130) return false;
    out.resize(a.count * want);
    for(size_t i = 0; i < a.count; i++)
        for(int c = 0; c < want; c++)
            out[i * want + c] = a.ctype == 5126 ? rd<float>(a.ptr + i * a.stride + 4 * c)
                                                : (float)rd<double>(a.ptr + i * a.stride + 8 * c);
    return true;
}

} // namespace

/* ---- PNG (8-bit, non-interlaced) via zlib ---------------------------------------------------- */
bool decode_image(const std::vector<uint8_t>& f, Texture& out, std::string& err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if(f.size() >= 3 && f[0] == 0xFF && f[1] == 0xD8 && f[2] == 0xFF) return decode_jpeg(f, out, err);
    if(f.size() < 8 || std::memcmp(f.data(), sig, 8) != 0) {
        err = "unsupported image format (PNG and baseline JPEG are decoded)";
        return false;
    }
    auto be32 = [&](size_t o) {
        return ((uint32_t)f[o] << 24) | ((uint32_t)f[o + 1] << 16) | ((uint32_t)f[o + 2] << 8) | f[o + 3];
    };
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    for(size_t o = 8; o + 12 <= f.size();) {
        uint32_t len = be32(o);
        std::string tag((const char*)&f[o + 4], 4);
        if(o + 12 + len > f.size()) break;
        const uint8_t* d = &f[o + 8];
        if(tag == "IHDR") {
            if(len < 13) break;
            w = be32(o + 8), h = be32(o + 12);
            depth = d[8], ctype = d[9], interlace = d[12];
        } else if(tag == "IDAT") idat.insert(idat.end(), d, d + len);
        else if(tag == "PLTE") plte.assign(d, d + len);
        else if(tag == "tRNS") trns.assign(d, d + len);
        else if(tag == "IEND") break;
        o += 12 + len;
    }
    if(!w || !h || depth != 8 || interlace) {
        err = "PNG variant not supported (need 8-bit, non-interlaced)";
        return false;
    }
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if(!ch) {
        err = "bad PNG colour type";
        return false;
    }
    size_t row = (size_t)w * ch;
    /* untrusted dimensions: deflate expands by at most 1032 : 1, so a header that asks for more than the IDAT bytes can
     * inflate to is wrong before anything is allocated (row, h < 2^34 and 2^32: the product fits 64 bits) */
    if((row + 1) * h > (idat.size() + 16) * 1032) {
        err = "PNG inflate failed";
        return false;
    }
    std::vector<uint8_t> raw((row + 1) * h);
    uLongf rawlen = (uLongf)raw.size();
    if(uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) {
        err = "PNG inflate failed";
        return false;
    }
    std::vector<uint8_t> img(row * h);
    for(uint32_t y = 0; y < h; y++) {
        const uint8_t* in = &raw[(row + 1) * y];
        uint8_t* cur = &img[row * y];
        const uint8_t* up = y ? &img[row * (y - 1)] : nullptr;
        int ft = in[0];
        for(size_t x = 0; x < row; x++) {
            int a = x >= (size_t)ch ? cur[x - ch] : 0, b = up ? up[x] : 0,
                c = (up && x >= (size_t)ch) ? up[x - ch] : 0, v = in[1 + x];
            switch(ft) {
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) / 2; break;
            case 4: {
                int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            } break;
            }
            cur[x] = (uint8_t)v;
        }
    }
    out.w = w, out.h = h;
    out.rgba.resize((size_t)w * h * 4);
    for(size_t i = 0; i < (size_t)w * h; i++) {
        uint8_t r, g, b, a = 255;
        const uint8_t* p = &img[i * ch];
        if(ctype == 0) r = g = b = p[0];
        else if(ctype == 2) r = p[0], g = p[1], b = p[2];
        else if(ctype == 4) r = g = b = p[0], a = p[1];
        else if(ctype == 6) r = p[0], g = p[1], b = p[2], a = p[3];
        else {
            size_t k = p[0];
            r = k * 3 + 2 < plte.size() ? plte[k * 3] : 0;
            g = k * 3 + 2 < plte.size() ? plte[k * 3 + 1] : 0;
            b = k * 3 + 2 < plte.size() ? plte[k * 3 + 2] : 0;
            a = k < trns.size() ? trns[k] : 255;
        }
        out.rgba[i * 4] = r, out.rgba[i * 4 + 1] = g, out.rgba[i * 4 + 2] = b, out.rgba[i * 4 + 3] = a;
    }
    return true;
}

/* ---------------------------------------------------------------------------------------------- */
bool Scene::load(const std::string& file, std::string& err) {
    clear();
    next_id = 1;

    Gltf g;
    std::string base;
    size_t slash = file.find_last_of("/\\");
    if(slash != std::string::npos) base = file.substr(0, slash + 1);

    std::vector<uint8_t> raw;
    if(!read_file(file, raw)) {
        err = "cannot read " + file;
        return false;
    }
    std::vector<uint8_t> glb_bin;
    std::string text;
    /* scene.cpp:329-333: "glb" anywhere in the path selects the binary container */
    if(file.find("glb") != std::string::npos) {
        if(raw.size() < 20 || std::memcmp(raw.data(), "glTF", 4) != 0) {
            err = "not a GLB container";
            return false;
        }
        size_t o = 12;
        while(o + 8 <= raw.size()) {
            uint32_t len = rd<uint32_t>(&raw[o]), type = rd<uint32_t>(&raw[o + 4]);
            if(o + 8 + len > raw.size()) break;
            if(type == 0x4E4F534A) text.assign((const char*)&raw[o + 8], len);
            else if(type == 0x004E4942) glb_bin.assign(&raw[o + 8], &raw[o + 8] + len);
            o += 8 + ((len + 3) & ~3u);
        }
    } else if(file.find("gltf") != std::string::npos) {
        text.assign((const char*)raw.data(), raw.size());
    } else {
        err = "unknown scene file type";
        return false;
    }
    if(!Json::parse(text, g.root, err)) return false;

    const Json& bufs = g.root["buffers"];
    for(size_t i = 0; i < bufs.size(); i++) {
        std::vector<uint8_t> data;
        if(bufs[i].has("uri")) {
            if(!load_uri(base, bufs[i]["uri"].string(), data)) {
                err = "cannot load buffer " + bufs[i]["uri"].string().substr(0, 64);
                return false;
            }
        } else
            data = glb_bin;
        g.buffers.push_back(std::move(data));
    }

    const Json& accessors = g.root["accessors"];
    const Json& materials = g.root["materials"];

    /* scene.cpp:48-315 */
    auto parse_mesh = [&](const Json& mesh, Pose pose) -> bool {
        const Json& prims = mesh["primitives"];
        for(size_t p = 0; p < prims.size(); p++) {
            const Json& prim = prims[p];
            std::vector<uint32_t> indices;
            Accessor ia;
            int iacc = prim["indices"].integer(-1);
            if(iacc < 0 || (size_t)iacc >= accessors.size() || !g.accessor(iacc, ia)) {
                err = "primitive without a valid index accessor";
                return false;
            }
            indices.reserve(ia.count);
            for(size_t i = 0; i < ia.count; i++) { /* scene.cpp:71-108 */
                const uint8_t* q = ia.ptr + ia.stride * i;
                switch(ia.ctype) {
                case 5120: indices.push_back((uint32_t)(int)rd<int8_t>(q)); break;
                case 5121: indices.push_back(rd<uint8_t>(q)); break;
                case 5122: indices.push_back((uint32_t)(int)rd<int16_t>(q)); break;
                case 5123: indices.push_back(rd<uint16_t>(q)); break;
                case 5124: indices.push_back((uint32_t)rd<int32_t>(q)); break;
                case 5125: indices.push_back(rd<uint32_t>(q)); break;
                default: break;
                }
            }
            int mode = prim["mode"].integer(4);
            if(mode == 6) { /* fan -> list, scene.cpp:114-129 */
                std::vector<uint32_t> fan = std::move(indices);
                indices.clear();
                for(size_t i = 2; i < fan.size(); ++i)
                    indices.push_back(fan[0]), indices.push_back(fan[i - 1]), indices.push_back(fan[i]);
            } else if(mode == 5) { /* strip -> list, scene.cpp:130-143 */
                std::vector<uint32_t> strip = std::move(indices);
                indices.clear();
                for(size_t i = 2; i < strip.size(); ++i)
                    indices.push_back(strip[i - 2]), indices.push_back(strip[i - 1]), indices.push_back(strip[i]);
            }
            std::vector<float> pos, nrm, tan, uv;
            if(mode == 4 || mode == 5 || mode == 6) {
                const Json& attrs = prim["attributes"];
                for(auto& kv : attrs.obj) {
                    Accessor a;
                    if(!g.accessor(kv.second.integer(-1), a)) continue;
                    if(kv.first == "POSITION") read_vec(a, 3, pos);
                    else if(kv.first == "NORMAL") read_vec(a, 3, nrm);
                    else if(kv.first == "TANGENT") read_vec(a, 4, tan);
                    else if(kv.first == "TEXCOORD_0") read_vec(a, 2, uv);
                }
            }
            if(!(mode == 4 || mode == 5 || mode == 6)) {
                /* points / lines: the reference keeps the index list next to an empty vertex array (scene.cpp:278-283)
                 * and dies in the Vulkan upload; here the object stays (ids and object order as in the reference)
                 * but holds no triangles */
                indices.clear();
                err += "[warn] primitive with mode " + std::to_string(mode) + " is not triangle based, ignored; ";
            }
            /* scene.cpp:286-302; tinygltf defaults: baseColor 1, metallic 1, roughness 1 */
            Material mat;
            const Json& gm = materials[(size_t)prim["material"].integer(-1)];
            const Json& pbr = gm["pbrMetallicRoughness"];
            const Json& bc = pbr["baseColorFactor"];
            mat.albedo = Vec3{(float)bc[0].number(1.0), (float)bc[1].number(1.0), (float)bc[2].number(1.0)};
            mat.albedo_tex = pbr["baseColorTexture"]["index"].integer(-1);
            const Json& em = gm["emissiveFactor"];
            mat.emissive = Vec3{(float)em[0].number(0.0), (float)em[1].number(0.0), (float)em[2].number(0.0)};
            mat.emissive_tex = gm["emissiveTexture"]["index"].integer(-1);
            mat.metal_rough.x = (float)pbr["metallicFactor"].number(1.0);
            mat.metal_rough.y = (float)pbr["roughnessFactor"].number(1.0);
            mat.metal_rough_tex = pbr["metallicRoughnessTexture"]["index"].integer(-1);
            mat.normal_tex = gm["normalTexture"]["index"].integer(-1);

            /* scene.cpp:304-311 */
            size_t nv = pos.size() / 3;
            std::vector<Vertex> verts(nv);
            for(size_t i = 0; i < nv; i++) {
                Vertex& v = verts[i];
                std::memset(&v, 0, sizeof(v));
                v.pos[0] = pos[3 * i], v.pos[1] = pos[3 * i + 1], v.pos[2] = pos[3 * i + 2];
                if(i < nrm.size() / 3) v.norm[0] = nrm[3 * i], v.norm[1] = nrm[3 * i + 1], v.norm[2] = nrm[3 * i + 2];
                if(i < tan.size() / 4) std::memcpy(v.tang, &tan[4 * i], 16);
                if(i < uv.size() / 2) v.pos[3] = uv[2 * i], v.norm[3] = uv[2 * i + 1];
            }
            Object obj;
            obj.id = reserve_id();
            obj.pose = pose;
            obj.mesh.set(std::move(verts), std::move(indices));
            obj.material = mat;
            add(std::move(obj));
        }
        return true;
    };

    /* scene.cpp:347-371 */
    const Json& nodes = g.root["nodes"];
    bool ok = true;
    int node_depth = 0; /* a cyclic or absurdly deep node graph must not overflow the stack */
    std::function<void(int, Mat4)> load_node = [&](int n, Mat4 T) {
        struct Depth {
            int& d;
            explicit Depth(int& d) : d(d) { d++; }
            ~Depth() { d--; }
        } guard(node_depth);
        if(node_depth > 256) {
            ok = false;
            err = "node graph deeper than 256 levels (cycle?)";
            return;
        }
        const Json& node = nodes[(size_t)n];
        Mat4 M;
        const Json& m = node["matrix"];
        for(size_t i = 0; i < 16 && i < m.size(); i++) M.data()[i] = (float)m[i].number(0.0);
        T = T * M.T();
        Pose pose;
        T.decompose(pose.pos, pose.scale, pose.euler);
        int mesh = node["mesh"].integer(-1);
        if(mesh >= 0 && ok) ok = parse_mesh(g.root["meshes"][(size_t)mesh], pose);
        const Json& ch = node["children"];
        for(size_t i = 0; i < ch.size(); i++) load_node(ch[i].integer(0), T);
    };
    const Json& scenes = g.root["scenes"];
    for(size_t s = 0; s < scenes.size(); s++) {
        const Json& roots = scenes[s]["nodes"];
        for(size_t r = 0; r < roots.size(); r++) load_node(roots[r].integer(0), Mat4());
    }
    if(!ok) return false;

    /* scene.cpp:373-388 */
    const Json& texs = g.root["textures"];
    const Json& imgs = g.root["images"];
    for(size_t t = 0; t < texs.size(); t++) {
        int src = texs[t]["source"].integer(-1);
        if(src < 0 || (size_t)src >= imgs.size()) continue;
        const Json& im = imgs[(size_t)src];
        std::vector<uint8_t> bytes;
        if(im.has("uri")) {
            if(!load_uri(base, im["uri"].string(), bytes)) {
                err = "cannot load image " + im["uri"].string().substr(0, 64);
                return false;
            }
        } else {
            const Json& bv = g.root["bufferViews"][(size_t)im["bufferView"].integer(-1)];
            size_t b = (size_t)bv["buffer"].integer(0);
            if(b < g.buffers.size()) {
                const double d_off = bv["byteOffset"].number(0), d_len = bv["byteLength"].number(0), size = (double)g.buffers[b].size();
                size_t off = 0, len = 0;
                if(d_off == d_off && d_len == d_len && d_off >= 0 && d_len >= 0 && d_off <= size && d_len <= size)
                    off = (size_t)d_off, len = (size_t)d_len;
                if(len && off <= g.buffers[b].size() - len)
                    bytes.assign(g.buffers[b].begin() + off, g.buffers[b].begin() + off + len);
            }
        }
        Texture tex;
        std::string ierr;
        if(!decode_image(bytes, tex, ierr)) {
            /* keep texture indices aligned: 1x1 white placeholder, error reported as a warning */
            tex.w = tex.h = 1;
            tex.rgba = {255, 255, 255, 255};
            err += "[warn] texture " + std::to_string(t) + ": " + ierr + "; ";
        }
        textures.push_back(std::move(tex));
    }
    return true;
}

} // namespace gpurt

/*
 * host_api.cpp — host-only half of the C ABI: error state, scene loading / packing, camera.
 * Needs no GPU; the CUDA half lives in csrc/api.cu.
 */
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "camera.h"
#include "internal.h"

namespace gpurt {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
} // namespace gpurt

using namespace gpurt;

bool gpurt_scene::pack() try {
    if(!dirty) return true;
    PackedScene& P = packed;
    P = PackedScene();
    scene.build_desc(P.descs, P.lights);
    /* a material may name a texture that does not exist (a glTF image that failed to load shifts the table; add_object
     * copies the ids verbatim): the reference would index past its descriptor array, here the slot falls back to the
     * constant factor (-1) and the last-error string says so */
    int ntex = (int)scene.textures.size(), dropped = 0;
    for(SceneDesc& d : P.descs)
        for(int32_t* t : {&d.albedo_tex, &d.emissive_tex, &d.metal_rough_tex, &d.normal_tex})
            if(*t >= ntex) *t = -1, dropped++;
    if(dropped) set_error("warning: " + std::to_string(dropped) + " material texture id(s) beyond the texture table were reset to -1");
    P.tri_off.push_back(0);
    P.vert_off.push_back(0);
    bool ok = true;
    scene.for_objs([&](const Object& o) {
        size_t nv = o.mesh.verts.size();
        for(uint32_t i : o.mesh.idx)
            if(i >= nv) ok = false;
        P.verts.insert(P.verts.end(), o.mesh.verts.begin(), o.mesh.verts.end());
        P.idx.insert(P.idx.end(), o.mesh.idx.begin(), o.mesh.idx.begin() + 3 * (o.mesh.idx.size() / 3));
        P.tri_off.push_back((uint32_t)(P.idx.size() / 3));
        P.vert_off.push_back((uint32_t)P.verts.size());
    });
    if(!ok) {
        set_error("scene has vertex indices out of range");
        return false;
    }
    dirty = false;
    version++;
    geom_version++;
    return true;
} catch(const std::exception& e) { /* packing a scene copies it: out of host memory is an error code, not a terminate */
    packed = PackedScene();
    dirty = true;
    set_error(std::string("scene could not be packed: ") + e.what());
    return false;
}

static void material_from_abi(Material& dst, const GpurtMaterial* m) {
    dst.albedo = Vec3{m->albedo[0], m->albedo[1], m->albedo[2]};
    dst.albedo_tex = m->albedo_tex;
    dst.emissive = Vec3{m->emissive[0], m->emissive[1], m->emissive[2]};
    dst.emissive_tex = m->emissive_tex;
    dst.metal_rough = Vec2{m->metal_rough[0], m->metal_rough[1]};
    dst.metal_rough_tex = m->metal_rough_tex;
    dst.normal_tex = m->normal_tex;
}

extern "C" {

const char* gpurt_last_error(void) { return g_error.c_str(); }
const char* gpurt_version(void) { return "gpurt-b200 0.1 (sm_100a)"; }

int gpurt_scene_create(gpurt_ctx* ctx, gpurt_scene** out) {
    if(!out) return set_error("out is NULL"), GPURT_E_INVALID;
    *out = new gpurt_scene;
    (*out)->ctx = ctx;
    return GPURT_OK;
}
int gpurt_scene_destroy(gpurt_scene* s) {
    delete s;
    return GPURT_OK;
}
int gpurt_scene_load_gltf(gpurt_scene* s, const char* path, float scale) {
    if(!s || !path) return set_error("NULL argument"), GPURT_E_INVALID;
    std::string err;
    s->dirty = true;
    s->scene.scale = scale;
    /* no exception leaves the C ABI: a file that asks for more memory than there is (or than a vector may hold) is an
     * unreadable file, not a reason to terminate the caller */
    try {
        if(!s->scene.load(path, err)) return set_error(err), GPURT_E_IO;
        if(!err.empty()) set_error(err); /* warnings */
        s->label = path;
        return s->pack() ? GPURT_OK : GPURT_E_INVALID;
    } catch(const std::exception& e) {
        s->scene.clear();
        return set_error(std::string("scene file rejected: ") + e.what()), GPURT_E_IO;
    }
}
int gpurt_scene_make_sponza_standin(gpurt_scene* s) {
    if(!s) return set_error("NULL argument"), GPURT_E_INVALID;
    s->dirty = true;
    s->scene.scale = 1.0f;
    make_sponza_standin(s->scene);
    s->label = "sponza_standin";
    return s->pack() ? GPURT_OK : GPURT_E_INVALID;
}
int gpurt_scene_add_object(gpurt_scene* s, const void* verts48, uint32_t nv, const uint32_t* idx,
                           uint32_t ni, const float model[16], const GpurtMaterial* m, uint32_t* out) {
    if(!s || (!verts48 && nv) || (!idx && ni) || !model) return set_error("NULL argument"), GPURT_E_INVALID;
    for(uint32_t i = 0; i < ni; i++)
        if(idx[i] >= nv) return set_error("index out of range"), GPURT_E_INVALID;
    try {
    Object o;
    o.id = s->scene.reserve_id();
    std::vector<Vertex> v(nv);
    if(nv) std::memcpy(v.data(), verts48, (size_t)nv * sizeof(Vertex));
    std::vector<uint32_t> ix(idx, idx + ni);
    o.mesh.set(std::move(v), std::move(ix));
    o.has_model = true;
    std::memcpy(o.model.data(), model, 64);
    if(m) material_from_abi(o.material, m);
    else {
        o.material.albedo = Vec3{1.0f};
        o.material.metal_rough = Vec2{1.0f, 1.0f};
    }
    unsigned int id = o.id;
    s->scene.add(std::move(o));
    s->dirty = true;
    if(out) { /* position in for_objs order */
        uint32_t k = 0, found = 0;
        s->scene.for_objs([&](const Object& ob) {
            if(ob.id == id) found = k;
            k++;
        });
        *out = found;
    }
    return GPURT_OK;
    } catch(const std::exception& e) {
        return set_error(std::string("object not added: ") + e.what()), GPURT_E_INVALID;
    }
}
int gpurt_scene_get_texture(const gpurt_scene* s, uint32_t tex, uint32_t* w, uint32_t* h, uint8_t* out) {
    if(!s || tex >= s->scene.textures.size()) return set_error("texture index out of range"), GPURT_E_INVALID;
    const Texture& t = s->scene.textures[tex];
    if(w) *w = t.w;
    if(h) *h = t.h;
    if(out) std::memcpy(out, t.rgba.data(), t.rgba.size());
    return GPURT_OK;
}
int gpurt_scene_set_transform(gpurt_scene* s, uint32_t obj, const float model[16]) {
    if(!s || !model) return set_error("NULL argument"), GPURT_E_INVALID;
    Object* o = s->scene.at_index(obj);
    if(!o) return set_error("object index out of range"), GPURT_E_INVALID;
    o->has_model = true;
    std::memcpy(o->model.data(), model, 64);
    if(!s->dirty) { /* geometry already packed: only Scene_Desc / Scene_Light change (rt.cpp:26-76) */
        s->scene.build_desc(s->packed.descs, s->packed.lights);
        s->version++;
    }
    return GPURT_OK;
}
int gpurt_scene_set_material(gpurt_scene* s, uint32_t obj, const GpurtMaterial* m) {
    if(!s || !m) return set_error("NULL argument"), GPURT_E_INVALID;
    Object* o = s->scene.at_index(obj);
    if(!o) return set_error("object index out of range"), GPURT_E_INVALID;
    material_from_abi(o->material, m);
    if(!s->dirty) { /* geometry already packed: only Scene_Desc / Scene_Light change (rt.cpp:26-76) */
        s->scene.build_desc(s->packed.descs, s->packed.lights);
        int ntex = (int)s->scene.textures.size();
        for(SceneDesc& d : s->packed.descs)
            for(int32_t* t : {&d.albedo_tex, &d.emissive_tex, &d.metal_rough_tex, &d.normal_tex})
                if(*t >= ntex) *t = -1;
        s->version++;
    }
    return GPURT_OK;
}
int gpurt_scene_set_ordered(gpurt_scene* s, int ordered) {
    if(!s) return set_error("NULL argument"), GPURT_E_INVALID;
    if(s->scene.ordered != (ordered != 0)) s->scene.ordered = ordered != 0, s->dirty = true;
    return GPURT_OK;
}
int gpurt_scene_clear_textures(gpurt_scene* s) {
    if(!s) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!s->scene.textures.empty()) s->scene.textures.clear(), s->dirty = true;
    return GPURT_OK;
}
int gpurt_scene_add_texture(gpurt_scene* s, const uint8_t* rgba, uint32_t w, uint32_t h, int32_t* out) {
    if(!s || !rgba || !w || !h) return set_error("bad texture"), GPURT_E_INVALID;
    Texture t;
    t.w = w, t.h = h;
    try {
        t.rgba.assign(rgba, rgba + (size_t)w * h * 4);
        s->scene.textures.push_back(std::move(t));
    } catch(const std::exception& e) {
        return set_error(std::string("texture not added: ") + e.what()), GPURT_E_INVALID;
    }
    s->dirty = true;
    if(out) *out = (int32_t)s->scene.textures.size() - 1;
    return GPURT_OK;
}
int gpurt_scene_counts(const gpurt_scene* cs, uint32_t* no, uint32_t* nt, uint32_t* nl, uint32_t* ntex) {
    gpurt_scene* s = const_cast<gpurt_scene*>(cs);
    if(!s) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!s->pack()) return GPURT_E_INVALID;
    if(no) *no = (uint32_t)s->packed.descs.size();
    if(nt) *nt = s->packed.tri_off.back();
    if(nl) *nl = (uint32_t)s->packed.lights.size();
    if(ntex) *ntex = (uint32_t)s->scene.textures.size();
    return GPURT_OK;
}
int gpurt_scene_tri_offsets(const gpurt_scene* cs, uint32_t* out) {
    gpurt_scene* s = const_cast<gpurt_scene*>(cs);
    if(!s || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!s->pack()) return GPURT_E_INVALID;
    std::memcpy(out, s->packed.tri_off.data(), s->packed.tri_off.size() * 4);
    return GPURT_OK;
}
int gpurt_scene_get_descs(const gpurt_scene* cs, GpurtSceneDesc* out) {
    gpurt_scene* s = const_cast<gpurt_scene*>(cs);
    if(!s || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!s->pack()) return GPURT_E_INVALID;
    static_assert(sizeof(GpurtSceneDesc) == sizeof(SceneDesc), "desc layout");
    std::memcpy(out, s->packed.descs.data(), s->packed.descs.size() * sizeof(SceneDesc));
    return GPURT_OK;
}
int gpurt_scene_get_lights(const gpurt_scene* cs, GpurtSceneLight* out) {
    gpurt_scene* s = const_cast<gpurt_scene*>(cs);
    if(!s || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!s->pack()) return GPURT_E_INVALID;
    static_assert(sizeof(GpurtSceneLight) == sizeof(SceneLight), "light layout");
    std::memcpy(out, s->packed.lights.data(), s->packed.lights.size() * sizeof(SceneLight));
    return GPURT_OK;
}
int gpurt_scene_object_sizes(const gpurt_scene* cs, uint32_t obj, uint32_t* nv, uint32_t* ni) {
    gpurt_scene* s = const_cast<gpurt_scene*>(cs);
    if(!s) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!s->pack()) return GPURT_E_INVALID;
    if((size_t)obj + 1 >= s->packed.tri_off.size()) return set_error("object index out of range"), GPURT_E_INVALID;
    if(nv) *nv = s->packed.vert_off[obj + 1] - s->packed.vert_off[obj];
    if(ni) *ni = 3 * (s->packed.tri_off[obj + 1] - s->packed.tri_off[obj]);
    return GPURT_OK;
}
int gpurt_scene_get_object(const gpurt_scene* cs, uint32_t obj, void* verts, uint32_t* idx) {
    gpurt_scene* s = const_cast<gpurt_scene*>(cs);
    if(!s) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!s->pack()) return GPURT_E_INVALID;
    if((size_t)obj + 1 >= s->packed.tri_off.size()) return set_error("object index out of range"), GPURT_E_INVALID;
    const PackedScene& P = s->packed;
    if(verts)
        std::memcpy(verts, P.verts.data() + P.vert_off[obj],
                    (size_t)(P.vert_off[obj + 1] - P.vert_off[obj]) * sizeof(Vertex));
    if(idx)
        std::memcpy(idx, P.idx.data() + 3ull * P.tri_off[obj],
                    12ull * (P.tri_off[obj + 1] - P.tri_off[obj]));
    return GPURT_OK;
}

/* Camera + RTPipe::update_uniforms (rt.cpp:121-127) */
int gpurt_camera_make(int mode, float w, float h, const float pos[3], const float center[3],
                      float vfov, GpurtCamera* out) {
    if(!out || w <= 0 || h <= 0) return set_error("bad camera arguments"), GPURT_E_INVALID;
    Camera cam(Vec2{w, h});
    if(mode == 1) {
        if(!pos || !center) return set_error("look_at camera needs pos and center"), GPURT_E_INVALID;
        cam.look_at(Vec3{center[0], center[1], center[2]}, Vec3{pos[0], pos[1], pos[2]});
        cam.set_fov(vfov);
    }
    Mat4 V = cam.get_view(), P = cam.get_proj();
    Mat4 iV = V.inverse(), iP = P.inverse();
    std::memset(out, 0, sizeof(*out));
    std::memcpy(out->V, V.data(), 64);
    std::memcpy(out->P, P.data(), 64);
    std::memcpy(out->iV, iV.data(), 64);
    std::memcpy(out->iP, iP.data(), 64);
    Mat4 pv = P * V; /* first frame: prev_PV = P*V of the same camera */
    std::memcpy(out->prev_PV, pv.data(), 64);
    out->new_samples = 4;          /* res_samples, rt.h:53 */
    out->temporal_multiplier = 16; /* temporal_scale, rt.h:50 */
    return GPURT_OK;
}

int gpurt_pipe_params_default(GpurtPipeParams* p) { /* rt.h:38-53 */
    if(!p) return set_error("NULL argument"), GPURT_E_INVALID;
    std::memset(p, 0, sizeof(*p));
    p->max_frames = 256, p->samples_per_frame = 8, p->max_depth = 8;
    p->clear[0] = p->clear[1] = p->clear[2] = 0.3f;
    p->env[0] = p->env[1] = p->env[2] = 1.0f;
    p->env_scale = 0.0f;
    p->use_normal_map = 0, p->use_rr = 1, p->use_metalness = 0, p->use_qmc = 0, p->use_temporal = 1;
    p->integrator = 0, p->temporal_scale = 16, p->brdf = 0, p->debug_view = 0, p->res_samples = 4;
    p->seed = 0;
    p->spatial_samples = 0, p->spatial_radius = 16.0f;
    p->light_sampling = 0;
    return GPURT_OK;
}

} /* extern "C" */

/* ---- OpenEXR output (no reference code: the reference links tinyexr) ------------------------------------------------
 * File layout (OpenEXR file-layout document): magic 20000630, version 2 (no flag bits: single part, scanlines, short
 * names), header = attributes (name\0 type\0 size value) ended by \0, offset table (one u64 per scanline block; without
 * compression a block is one scanline), blocks (y, byte count, then the row channel by channel in header order). */
namespace {
struct ExrOut {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void i32(int32_t v) {
        for(int k = 0; k < 4; k++) b.push_back((uint8_t)((uint32_t)v >> (8 * k)));
    }
    void u64(uint64_t v) {
        for(int k = 0; k < 8; k++) b.push_back((uint8_t)(v >> (8 * k)));
    }
    void f32(float v) {
        uint32_t u;
        std::memcpy(&u, &v, 4);
        i32((int32_t)u);
    }
    void str(const char* s) {
        while(*s) b.push_back((uint8_t)*s++);
        b.push_back(0);
    }
    void attr(const char* name, const char* type, int32_t size) { str(name), str(type), i32(size); }
};
} // namespace
int gpurt_write_exr(const char* path, const float* rgba, uint32_t w, uint32_t h) {
    if(!path || !rgba || !w || !h || w > (1u << 24) || h > (1u << 24)) return set_error("write_exr: bad argument"), GPURT_E_INVALID;
    ExrOut o;
    o.i32(20000630), o.i32(2);
    static const char* const names[4] = {"A", "B", "G", "R"};   /* channels are stored in alphabetical order */
    static const int source[4] = {3, 2, 1, 0};                   /* index into the RGBA pixel */
    o.attr("channels", "chlist", 4 * (2 + 16) + 1);
    for(int c = 0; c < 4; c++) {
        o.str(names[c]);
        o.i32(2);                                               /* FLOAT */
        o.u8(0), o.u8(0), o.u8(0), o.u8(0);                     /* pLinear + reserved */
        o.i32(1), o.i32(1);                                     /* x / y sampling */
    }
    o.u8(0);
    o.attr("compression", "compression", 1), o.u8(0);
    o.attr("dataWindow", "box2i", 16), o.i32(0), o.i32(0), o.i32((int32_t)w - 1), o.i32((int32_t)h - 1);
    o.attr("displayWindow", "box2i", 16), o.i32(0), o.i32(0), o.i32((int32_t)w - 1), o.i32((int32_t)h - 1);
    o.attr("lineOrder", "lineOrder", 1), o.u8(0);               /* increasing y */
    o.attr("pixelAspectRatio", "float", 4), o.f32(1.0f);
    o.attr("screenWindowCenter", "v2f", 8), o.f32(0.0f), o.f32(0.0f);
    o.attr("screenWindowWidth", "float", 4), o.f32(1.0f);
    o.u8(0);
    const uint64_t row_bytes = (uint64_t)w * 16, block = 8 + row_bytes, first = o.b.size() + 8ull * h;
    for(uint32_t y = 0; y < h; y++) o.u64(first + block * y);
    FILE* f = fopen(path, "wb");
    if(!f) return set_error(std::string("write_exr: cannot open ") + path), GPURT_E_IO;
    bool ok = fwrite(o.b.data(), 1, o.b.size(), f) == o.b.size();
    std::vector<float> row((size_t)w * 4);
    for(uint32_t y = 0; y < h && ok; y++) {
        ExrOut hd;
        hd.i32((int32_t)y), hd.i32((int32_t)row_bytes);
        const float* src = rgba + (size_t)y * w * 4;
        for(int c = 0; c < 4; c++)
            for(uint32_t x = 0; x < w; x++) row[(size_t)c * w + x] = src[4ull * x + source[c]];
        ok = fwrite(hd.b.data(), 1, 8, f) == 8 && fwrite(row.data(), 4, row.size(), f) == row.size(); /* little-endian host */
    }
    ok = fclose(f) == 0 && ok;
    if(!ok) return set_error(std::string("write_exr: short write to ") + path), GPURT_E_IO;
    return GPURT_OK;
}

/*
 * scene.h — host scene front end: the reference's Scene / Object / Pose / Material
 * (src/scene/ headers) and VK::Mesh's host half (src/vk/mesh.h:14-58), feeding the C ABI.
 */
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "hmath.h"

namespace gpurt {

/* VK::Mesh::Vertex (src/vk/mesh.h:16-22): pos.xyz|u, norm.xyz|v, tangent.xyzw — 48 bytes */
struct Vertex {
    float pos[4], norm[4], tang[4];
};
static_assert(sizeof(Vertex) == 48, "Mesh::Vertex stride");

struct Mesh {
    std::vector<Vertex> verts;
    std::vector<uint32_t> idx;
    BBox bbox; /* object space, src/vk/mesh.cpp:73-75 */
    void set(std::vector<Vertex>&& v, std::vector<uint32_t>&& i);
};

/* src/scene/material.h:15-21 */
struct Material {
    Vec3 albedo;
    int albedo_tex = -1;
    Vec3 emissive;
    int emissive_tex = -1;
    Vec2 metal_rough;
    int metal_rough_tex = -1;
    int normal_tex = -1;
};

/* src/scene/pose.h:8-13, pose.cpp:4-10 (GLOBAL_SCALE = 1) */
struct Pose {
    Vec3 pos, euler, scale{1.0f};
    Mat4 transform() const {
        return Mat4::translate(pos) * Mat4::euler(euler) * Mat4::scale(scale) * Mat4::scale(Vec3{1.0f});
    }
};

struct Object {
    unsigned int id = 0;
    Pose pose;
    Mesh mesh;
    Material material;
    /* When set, `model` overrides scale*pose.transform() (objects added through the C ABI). */
    bool has_model = false;
    Mat4 model;
};

struct Texture {
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> rgba;
};

/* src/vk/rt.h:67-84 host mirrors (== GpurtSceneDesc / GpurtSceneLight) */
struct SceneDesc {
    Mat4 model, modelIT;
    float albedo[4], emissive[4], metal_rough[4];
    int32_t albedo_tex, emissive_tex, metal_rough_tex, normal_tex;
    uint32_t index;
    uint32_t pad[3];
};
struct SceneLight {
    float bmin[4], bmax[4];
    uint32_t index, n_triangles, pad[2];
};
static_assert(sizeof(SceneDesc) == 208 && sizeof(SceneLight) == 48, "std430 mirrors");

class Scene {
public:
    /* src/scene/scene.cpp:317-391 */
    bool load(const std::string& file, std::string& err);
    void clear();
    unsigned int add(Object&& obj); /* scene.cpp:19-23 */
    unsigned int reserve_id() { return next_id++; }
    size_t size() const { return objs.size(); }

    /* iteration order of std::unordered_map<unsigned, Object> — SURVEY Q2: never sort.  `ordered` (set by a caller
     * that passes instances in an order of its own, e.g. the BLAS_T order of GPURT::build_accel) iterates in insertion
     * order instead, so that instance i is object i. */
    template <typename F> void for_objs(F&& f) const {
        if(ordered) {
            for(unsigned int id : insertion) f(objs.at(id));
            return;
        }
        for(auto& o : objs) f(o.second);
    }

    /* k-th object in for_objs order (the obj_id the shaders see), or nullptr */
    Object* at_index(size_t k) {
        if(ordered) return k < insertion.size() ? &objs.at(insertion[k]) : nullptr;
        for(auto& o : objs)
            if(k-- == 0) return &o.second;
        return nullptr;
    }
    bool ordered = false;

    /* RTPipe::build_desc, src/vk/rt.cpp:26-76 */
    void build_desc(std::vector<SceneDesc>& descs, std::vector<SceneLight>& lights) const;

    float scale = 1.0f; /* scene.h:42 */
    std::vector<Texture> textures;

private:
    std::unordered_map<unsigned int, Object> objs; /* scene.h:46 */
    std::vector<unsigned int> insertion;           /* ids in the order add() saw them */
    unsigned int next_id = 1;
};

/* Procedural stand-in for media/sponza (geometry blob missing from the snapshot). */
void make_sponza_standin(Scene& scene);

/* glTF texture decode to RGBA8 with the reference's (tinygltf -> stb_image, 4 channels requested)
 * results: PNG (zlib; 8-bit, non-interlaced) and baseline JPEG (jpeg.cpp). */
bool decode_image(const std::vector<uint8_t>& file, Texture& out, std::string& err);
bool decode_jpeg(const std::vector<uint8_t>& file, Texture& out, std::string& err);

} // namespace gpurt

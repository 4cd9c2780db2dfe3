/* internal.h — structures shared by the host half (host_api.cpp, g++) and the CUDA half
 * (csrc/api.cu, nvcc) of libgpurt.so.  Not part of the public ABI. */
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/gpurt.h"
#include "scene.h"

namespace gpurt {

/* Flat arrays in Scene::for_objs order: what RTPipe::build_desc (rt.cpp:26-76) and the
 * per-object vbuf/ibuf descriptor arrays (rt.cpp:662-673) hand to the GPU. */
struct PackedScene {
    std::vector<SceneDesc> descs;
    std::vector<SceneLight> lights;
    std::vector<uint32_t> tri_off;  /* n_objs+1, in triangles */
    std::vector<uint32_t> vert_off; /* n_objs+1, in vertices */
    std::vector<Vertex> verts;      /* concatenated, 48-byte stride */
    std::vector<uint32_t> idx;      /* concatenated, object-local vertex indices */
};

void set_error(const std::string& msg);

} // namespace gpurt

struct gpurt_scene {
    gpurt_ctx* ctx = nullptr;
    gpurt::Scene scene;
    gpurt::PackedScene packed;
    bool dirty = true;
    uint64_t version = 0;      /* bumps on every change of `packed` */
    uint64_t geom_version = 0; /* bumps only when geometry / textures were repacked (not on pose changes) */
    std::string label = "custom";
    /* (re)build `packed` from `scene`; returns false (with error set) on invalid indices */
    bool pack();
};

/*
 * scene_view.cuh — the flat scene arrays as the kernels see them.  Free of CUDA runtime types so
 * that shade.cuh also compiles for the host (tests/emu replays the shading code on the CPU; that
 * harness is test-only and never part of the product).
 */
#pragma once
#include "../host/internal.h"
#include "common.cuh"

namespace gpurt {

/* Device copy of PackedScene: the descriptor arrays of src/vk/rt.cpp:529-741 as plain pointers. */
struct DeviceScene {
    Vertex* verts = nullptr;
    uint32_t* idx = nullptr;
    uint32_t* tri_off = nullptr;
    uint32_t* vert_off = nullptr;
    SceneDesc* descs = nullptr;
    SceneLight* lights = nullptr;
    uint32_t n_objs = 0, n_tris = 0, n_lights = 0, n_verts = 0;
    uint64_t version = 0, geom_version = 0;
    /* textures: RGBA8 texels, one allocation, per-texture (offset,w,h) table */
    uint8_t* texels = nullptr;
    uint4* tex_info = nullptr; /* x: texel offset, y: w, z: h */
    uint32_t n_textures = 0;
};

} // namespace gpurt

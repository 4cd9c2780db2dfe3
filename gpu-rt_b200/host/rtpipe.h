/*
 * rtpipe.h — C++ shim that re-creates the reference's VK::Accel / VK::RTPipe call surface on top of
 * the C ABI (include/gpurt.h), so that GPURT's call sites (src/gpurt.cpp:39-45, :216-241) keep their
 * shape:
 *
 *     reference                                   here
 *     VK::Accel(obj.mesh()) per object +          gpurt::Accel accel(scene);        // BLAS+TLAS in one
 *       TLAS->recreate(BLAS, BLAS_T)
 *     rt_pipe.recreate(scene)                     gpurt::RTPipe rt_pipe(scene, accel);
 *     rt_pipe.use_image / use_accel               (owned by the pipe)
 *     rt_pipe.update_uniforms(cam)                folded into trace()
 *     rt_pipe.trace(cam, cmds, ext)               rt_pipe.trace(cam, ext.width, ext.height)
 *     rt_pipe.reset_frame()                       rt_pipe.reset_frame()
 *     public tunables (rt.h:38-53)                same names, same defaults
 *
 * Header-only; link against libgpurt.so.
 */
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gpurt.h"

namespace gpurt {

inline void check(int rc) {
    if(rc < 0) throw std::runtime_error(std::string("gpurt: ") + gpurt_last_error());
}

class Context {
public:
    explicit Context(int device = 0) { check(gpurt_ctx_create(device, &h)); }
    ~Context() { gpurt_ctx_destroy(h); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    gpurt_ctx* h = nullptr;
};

/* Scene (src/scene/scene.h:18-42) */
class SceneHandle {
public:
    explicit SceneHandle(Context& ctx) { check(gpurt_scene_create(ctx.h, &h)); }
    ~SceneHandle() { gpurt_scene_destroy(h); }
    SceneHandle(const SceneHandle&) = delete;
    SceneHandle& operator=(const SceneHandle&) = delete;
    /* Scene::load(file, cam) — scene.cpp:317 */
    void load(const std::string& file, float scale = 1.0f) { check(gpurt_scene_load_gltf(h, file.c_str(), scale)); }
    void make_sponza_standin() { check(gpurt_scene_make_sponza_standin(h)); }
    gpurt_scene* h = nullptr;
};

/* VK::Accel (src/vk/vulkan.h:256-284) */
class Accel {
public:
    explicit Accel(SceneHandle& scene, uint32_t flags = GPURT_BUILD_DEFAULT) { check(gpurt_accel_build(scene.h, flags, &h)); }
    ~Accel() { gpurt_accel_destroy(h); }
    Accel(const Accel&) = delete;
    Accel& operator=(const Accel&) = delete;
    /* GPURT::build_accel after edit_scene (src/gpurt.cpp:220-241, :378-385) */
    void update() { check(gpurt_accel_update(h)); }
    GpurtAccelInfo info() const {
        GpurtAccelInfo i;
        check(gpurt_accel_info(h, &i));
        return i;
    }
    gpurt_accel* h = nullptr;
};

/* VK::RTPipe (src/vk/rt.h:14-142) */
class RTPipe {
public:
    RTPipe(SceneHandle& scene, Accel& accel) { /* recreate(scene) + use_accel(tlas) */
        check(gpurt_pipe_create(scene.h, accel.h, &h));
    }
    ~RTPipe() { gpurt_pipe_destroy(h); }
    RTPipe(const RTPipe&) = delete;
    RTPipe& operator=(const RTPipe&) = delete;

    /* rt.h:38-53 */
    int max_frames = 256;
    int samples_per_frame = 8;
    int max_depth = 8;
    float clear[3] = {0.3f, 0.3f, 0.3f};
    float env[3] = {1.0f, 1.0f, 1.0f};
    float env_scale = 0.0f;
    bool use_normal_map = false;
    bool use_rr = true;
    bool use_metalness = false;
    bool use_qmc = false;
    bool use_temporal = true;
    int integrator = 0;
    int temporal_scale = 16;
    int brdf = 0;
    int debug_view = 0;
    int res_samples = 4;
    unsigned seed = 0; /* replaces clockARB() (rt.rgen:569) */
    int spatial_samples = 0;     /* extension: ReSTIR spatial reuse (include/gpurt.h), off by default */
    float spatial_radius = 16.0f;
    int light_sampling = 0;      /* extension: 1 = light triangles chosen in proportion to their power (include/gpurt.h) */

    void reset_frame() { check(gpurt_pipe_reset_frame(h)); }

    GpurtPipeParams params() const {
        GpurtPipeParams p;
        gpurt_pipe_params_default(&p);
        p.max_frames = max_frames, p.samples_per_frame = samples_per_frame, p.max_depth = max_depth;
        for(int k = 0; k < 3; k++) p.clear[k] = clear[k], p.env[k] = env[k];
        p.env_scale = env_scale;
        p.use_normal_map = use_normal_map, p.use_rr = use_rr, p.use_metalness = use_metalness, p.use_qmc = use_qmc;
        p.use_temporal = use_temporal, p.integrator = integrator, p.temporal_scale = temporal_scale, p.brdf = brdf;
        p.debug_view = debug_view, p.res_samples = res_samples, p.seed = seed;
        p.spatial_samples = spatial_samples, p.spatial_radius = spatial_radius, p.light_sampling = light_sampling;
        return p;
    }
    /* multi-GPU, frame-parallel: render frame f into a device buffer / fold a frame mean (include/gpurt.h) */
    void render_frame_mean(const GpurtCamera& cam, unsigned width, unsigned height, int frame, void* mean_out_device) {
        GpurtPipeParams p = params();
        check(gpurt_pipe_render_frame_mean(h, &p, &cam, width, height, frame, mean_out_device));
        w_ = width, h_ = height;
    }
    void accumulate_mean(const void* mean_device, int frame) { check(gpurt_pipe_accumulate_mean(h, mean_device, frame, w_, h_)); }

    /* update_uniforms(cam) + trace(cam, cmds, ext): false when frame >= max_frames (rt.cpp:353) */
    bool trace(const GpurtCamera& cam, unsigned width, unsigned height) {
        GpurtPipeParams p = params();
        int rc = gpurt_pipe_render_frame(h, &p, &cam, width, height);
        check(rc);
        w_ = width, h_ = height;
        return rc == 0;
    }
    std::vector<float> read_image() {
        std::vector<float> img((size_t)w_ * h_ * 4);
        check(gpurt_pipe_read_image(h, img.data(), GPURT_MEM_HOST));
        return img;
    }
    void read_image_async(float* out_pinned) { check(gpurt_pipe_read_image_async(h, out_pinned)); }
    void read_image_wait() { check(gpurt_pipe_read_image_wait(h)); }
    /* EffectPipe::tonemap (src/vk/effect.cpp:32-61) + save_rt's framebuffer read (gpurt.cpp:258-262) */
    std::vector<uint8_t> tonemap(int op = 1, float exposure = 1.0f, float gamma = 2.2f) {
        std::vector<uint8_t> out((size_t)w_ * h_ * 4);
        check(gpurt_tonemap(h, op, exposure, gamma, out.data(), GPURT_MEM_HOST));
        return out;
    }
    gpurt_pipe* h = nullptr;

private:
    unsigned w_ = 0, h_ = 0;
};

} // namespace gpurt

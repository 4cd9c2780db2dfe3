/*
 * vk_dropin.h — VK::Accel and VK::RTPipe with the REFERENCE'S OWN SIGNATURES on top of the C ABI
 * (include/gpurt.h), so that GPURT's call sites compile unchanged:
 *
 *     src/gpurt.cpp:39-45    rt_pipe.use_image(rt_target_view); rt_pipe.use_accel(TLAS);
 *                            rt_pipe.update_uniforms(cam); rt_pipe.trace(cam, cmds, {w, h});
 *     src/gpurt.cpp:216-218  rt_pipe.recreate(scene);
 *     src/gpurt.cpp:220-241  BLAS.push_back({VK::Accel(obj.mesh())}); ... TLAS.drop(); TLAS->recreate(BLAS, BLAS_T);
 *     src/gpurt.cpp:283-316  the public tunables (rt.h:38-53), rt_pipe.reset_frame()
 *
 * It replaces the declarations at src/vk/vulkan.h:256-284 (Accel), :305-336 (Drop) and src/vk/rt.h:13-142 (RTPipe).
 * Everything else it touches is the reference's own and must be declared before this header is included:
 *
 *     VK::Mesh     verts() / inds(), Vertex = 48 bytes                        src/vk/mesh.h:15-45
 *     Mat4         64 bytes, column-major                                      src/lib/mat4.h:65-72
 *     Scene        for_objs(f), scale, images()                                src/scene/scene.h:12-48
 *     Object       id(), mesh(), pose.transform(), material                    src/scene/object.h
 *     Camera       get_view(), get_proj()                                      src/util/camera.h:20-22
 *     VkCommandBuffer, VkExtent2D                                              (only named; a typedef / {width,height})
 *
 * Header-only; link against libgpurt.so.  Errors: the reference exits the process on any Vulkan error
 * (VK_CHECK -> die, vulkan.h:26-33); here every failed C call throws VK::DropinError carrying gpurt_last_error().
 *
 * What maps to what:
 *   Accel(mesh)               a BLAS keeps the mesh's host arrays (vulkan.cpp:881-936 uploads + builds per object; here the
 *                             geometry of all instances goes to the device in ONE build when the TLAS is created)
 *   Accel::recreate(blas, T)  gpurt_scene_add_object per instance (instance i = object i = gl_InstanceCustomIndexEXT,
 *                             vulkan.cpp:785-805) + gpurt_accel_build.  The reference's Drop<T>::drop() defers destruction
 *                             to an erase queue (vulkan.h:562-564); the same deferral here parks the dropped TLAS until
 *                             the next one is built, and a TLAS over the SAME BLAS objects adopts it: only the instance
 *                             matrices are rewritten and the BVH is refitted in place (gpurt_scene_set_transform +
 *                             gpurt_accel_update_auto) — the reference's rebuild_tlas-without-rebuild_blas case (gpurt.cpp:228-237).
 *   RTPipe::recreate(scene)   build_textures + build_desc (rt.cpp:16-24, :26-76, :430-455): materials and textures are
 *                             captured here and applied to the TLAS's scene in use_accel (gpurt_scene_set_material,
 *                             gpurt_accel_sync_scene); reset_frame()
 *   update_uniforms(cam)      V, P, iV = V.inverse(), iP = P.inverse() (rt.cpp:121-127); prev_PV / the frame reset on a
 *                             camera change happen inside gpurt_pipe_render_frame exactly as rt.cpp:128-135
 *   trace(cam, cmds, ext)     gpurt_pipe_render_frame; returns false once frame >= max_frames (rt.cpp:353)
 *   use_image(view)           the target image is owned by the pipe (rt_target, RGBA32F): read it with read_image() /
 *                             device_image() / tonemap()
 */
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gpurt.h"

#if !defined(VULKAN_CORE_H_) && !defined(GPURT_DROPIN_HAVE_VK_HANDLES)
typedef struct VkCommandBuffer_T* VkCommandBuffer;
struct VkExtent2D {
    uint32_t width, height;
};
#endif

namespace VK {

struct DropinError : std::runtime_error {
    DropinError(int code, const std::string& what) : std::runtime_error(what), code(code) {}
    int code;
};
inline int dropin_check(int rc) {
    if(rc < 0) throw DropinError(rc, std::string("gpurt: ") + gpurt_last_error());
    return rc;
}
/* VK::vk() (vulkan.cpp:17-20): one context per process, on device GPURT_DEVICE (default 0) */
inline gpurt_ctx* dropin_ctx() {
    struct Holder {
        gpurt_ctx* h = nullptr;
        Holder() {
            const char* e = std::getenv("GPURT_DEVICE");
            dropin_check(gpurt_ctx_create(e ? std::atoi(e) : 0, &h));
        }
        ~Holder() { gpurt_ctx_destroy(h); }
    };
    static Holder holder;
    return holder.h;
}

#if !defined(GPURT_DROPIN_NO_DROP)
/* vulkan.h:305-336, :559-564 */
template <typename T> struct Drop {
    Drop() = default;
    Drop(T&& resource) : resource(std::move(resource)) {}
    ~Drop() = default;
    Drop(const Drop&) = delete;
    Drop(Drop&& src) = default;
    Drop& operator=(const Drop&) = delete;
    Drop& operator=(Drop&& src) = default;
    T* operator->() { return &resource; }
    const T* operator->() const { return &resource; }
    operator T&() { return resource; }
    operator T const&() const { return resource; }
    void drop() {
        T gone = std::move(resource);
        resource = T();
        (void)gone; /* destroyed here; Accel parks its device state (see Accel::destroy) */
    }

private:
    T resource;
};
#endif

/* src/vk/vulkan.h:256-284 */
struct Accel {
    Accel() = default;
    Accel(const Mesh& mesh) { recreate(mesh); }
    Accel(const std::vector<Drop<Accel>>& blas, const std::vector<Mat4>& inst) { recreate(blas, inst); }
    ~Accel() { destroy(); }

    Accel(const Accel&) = delete;
    Accel(Accel&& src) { *this = std::move(src); }
    Accel& operator=(const Accel&) = delete;
    Accel& operator=(Accel&& src) {
        if(this != &src) {
            destroy();
            geom = std::move(src.geom), top = std::move(src.top);
        }
        return *this;
    }

    /* BLAS (vulkan.cpp:881-936): vertices at stride 48 / RGB32F positions, u32 indices, opaque */
    void recreate(const Mesh& mesh) {
        destroy();
        static_assert(sizeof(typename std::remove_reference<decltype(mesh.verts()[0])>::type) == 48, "Mesh::Vertex is 48 bytes");
        auto g = std::make_shared<Geometry>();
        const auto& v = mesh.verts();
        const auto& ix = mesh.inds();
        g->verts.resize(v.size() * 48);
        if(!v.empty()) std::memcpy(g->verts.data(), v.data(), g->verts.size());
        g->idx.assign(ix.begin(), ix.end());
        g->idx.resize(g->idx.size() / 3 * 3);
        geom = std::move(g);
    }
    /* TLAS (vulkan.cpp:777-856): instance i = blas[i] under inst[i], instanceCustomIndex = i */
    void recreate(const std::vector<Drop<Accel>>& blas, const std::vector<Mat4>& inst) {
        destroy();
        if(blas.size() != inst.size()) throw DropinError(GPURT_E_INVALID, "Accel::recreate: blas / inst sizes differ");
        std::vector<std::shared_ptr<const Geometry>> parts;
        for(const auto& b : blas) {
            if(!b->geom) throw DropinError(GPURT_E_INVALID, "Accel::recreate: an element of blas is not a bottom-level Accel");
            parts.push_back(b->geom);
        }
        std::shared_ptr<Top>& park = parked();
        if(park && park->parts == parts) { /* same BLAS objects: pose edit only */
            top = std::move(park);
            for(size_t i = 0; i < inst.size(); i++)
                if(std::memcmp(&top->inst[i], &inst[i], 64) != 0) {
                    dropin_check(gpurt_scene_set_transform(top->scene, (uint32_t)i, reinterpret_cast<const float*>(&inst[i])));
                    top->inst[i] = inst[i];
                }
            dropin_check(gpurt_accel_update_auto(top->accel, 0.0f)); /* refit while the old topology still fits */
            top->generation = next_generation();
            return;
        }
        park.reset();
        auto t = std::make_shared<Top>();
        dropin_check(gpurt_scene_create(dropin_ctx(), &t->scene));
        dropin_check(gpurt_scene_set_ordered(t->scene, 1));
        for(size_t i = 0; i < parts.size(); i++) {
            uint32_t index = 0;
            dropin_check(gpurt_scene_add_object(t->scene, parts[i]->verts.data(), (uint32_t)(parts[i]->verts.size() / 48),
                                                parts[i]->idx.data(), (uint32_t)parts[i]->idx.size(),
                                                reinterpret_cast<const float*>(&inst[i]), nullptr, &index));
            if(index != i) throw DropinError(GPURT_E_STATE, "Accel::recreate: instance order was not kept");
        }
        dropin_check(gpurt_accel_build(t->scene, build_flags(), &t->accel));
        t->parts = std::move(parts), t->inst = inst, t->generation = next_generation();
        top = std::move(t);
    }
    void recreate(const Drop<Accel>& blas, Mat4 inst) {
        destroy();
        std::vector<std::shared_ptr<const Geometry>> one{blas->geom};
        if(!one[0]) throw DropinError(GPURT_E_INVALID, "Accel::recreate: not a bottom-level Accel");
        auto t = std::make_shared<Top>();
        dropin_check(gpurt_scene_create(dropin_ctx(), &t->scene));
        dropin_check(gpurt_scene_set_ordered(t->scene, 1));
        dropin_check(gpurt_scene_add_object(t->scene, one[0]->verts.data(), (uint32_t)(one[0]->verts.size() / 48), one[0]->idx.data(),
                                            (uint32_t)one[0]->idx.size(), reinterpret_cast<const float*>(&inst), nullptr, nullptr));
        dropin_check(gpurt_accel_build(t->scene, build_flags(), &t->accel));
        t->parts = std::move(one), t->inst = {inst}, t->generation = next_generation();
        top = std::move(t);
    }
    void destroy() {
        geom.reset();
        if(top) parked() = std::move(top); /* deferred, like the erase queue: freed when the next TLAS is built or parked */
    }

    /* ---- B200 side ---- */
    struct Geometry {
        std::vector<unsigned char> verts; /* 48-byte Mesh::Vertex records */
        std::vector<uint32_t> idx;
    };
    struct Top {
        gpurt_scene* scene = nullptr;
        gpurt_accel* accel = nullptr;
        std::vector<std::shared_ptr<const Geometry>> parts;
        std::vector<Mat4> inst;
        uint64_t uid = 0;        /* identity of this device build (kept when a pose edit adopts it) */
        uint64_t generation = 0; /* bumps on every recreate */
        Top() : uid(next_generation()) {}
        Top(const Top&) = delete;
        Top& operator=(const Top&) = delete;
        ~Top() {
            if(accel) gpurt_accel_destroy(accel);
            if(scene) gpurt_scene_destroy(scene);
        }
    };
    bool is_top() const { return (bool)top; }
    gpurt_accel* handle() const { return top ? top->accel : nullptr; }
    gpurt_scene* scene() const { return top ? top->scene : nullptr; }
    uint64_t generation() const { return top ? top->generation : 0; }
    /* shared ownership of the device build: a pipe bound to it keeps it alive past TLAS.drop() */
    std::shared_ptr<const Top> share() const { return top; }
    /* release the parked TLAS now (the reference: Manager::destroy drains the erase queues) */
    static void release_deferred() { parked().reset(); }
    /* GPURT_BUILD_* flags for the next TLAS builds (the reference asks for PREFER_FAST_TRACE, vulkan.cpp:780, :884) */
    static uint32_t& build_flags() {
        static uint32_t flags = GPURT_BUILD_DEFAULT;
        return flags;
    }

private:
    static std::shared_ptr<Top>& parked() {
        static std::shared_ptr<Top> p;
        return p;
    }
    static uint64_t next_generation() {
        static uint64_t g = 0;
        return ++g;
    }
    std::shared_ptr<const Geometry> geom;
    std::shared_ptr<Top> top;
};

/* src/vk/rt.h:13-142 */
struct RTPipe {
    RTPipe() = default;
    RTPipe(const Scene& scene) { recreate(scene); }
    ~RTPipe() { destroy(); }

    RTPipe(const RTPipe&) = delete;
    RTPipe(RTPipe&& src) { *this = std::move(src); }
    RTPipe& operator=(const RTPipe&) = delete;
    RTPipe& operator=(RTPipe&& src) {
        if(this != &src) {
            destroy();
            max_frames = src.max_frames, samples_per_frame = src.samples_per_frame, max_depth = src.max_depth;
            clear = src.clear, env = src.env, env_scale = src.env_scale, use_normal_map = src.use_normal_map;
            use_rr = src.use_rr, use_metalness = src.use_metalness, use_qmc = src.use_qmc, use_temporal = src.use_temporal;
            integrator = src.integrator, temporal_scale = src.temporal_scale, brdf = src.brdf, debug_view = src.debug_view;
            res_samples = src.res_samples, seed = src.seed, spatial_samples = src.spatial_samples, spatial_radius = src.spatial_radius, light_sampling = src.light_sampling;
            mats = std::move(src.mats), texs = std::move(src.texs), mats_version = src.mats_version;
            pipe = src.pipe, src.pipe = nullptr;
            bound = std::move(src.bound), bound_generation = src.bound_generation, applied_version = src.applied_version;
            cam_ubo = src.cam_ubo, have_cam = src.have_cam, w_ = src.w_, h_ = src.h_;
        }
        return *this;
    }

    /* rt.cpp:16-24: build_textures(scene) + build_desc(scene) + reset_frame() */
    void recreate(const Scene& scene) {
        mats.clear(), texs.clear();
        scene.for_objs([&](const Object& obj) { /* rt.cpp:31-45 */
            GpurtMaterial m;
            m.albedo[0] = obj.material.albedo.x, m.albedo[1] = obj.material.albedo.y, m.albedo[2] = obj.material.albedo.z;
            m.emissive[0] = obj.material.emissive.x, m.emissive[1] = obj.material.emissive.y, m.emissive[2] = obj.material.emissive.z;
            m.metal_rough[0] = obj.material.metal_rough.x, m.metal_rough[1] = obj.material.metal_rough.y;
            m.albedo_tex = obj.material.albedo_tex, m.emissive_tex = obj.material.emissive_tex;
            m.metal_rough_tex = obj.material.metal_rough_tex, m.normal_tex = obj.material.normal_tex;
            mats.push_back(m);
        });
        for(const auto& image : scene.images()) { /* rt.cpp:432-444: R8G8B8A8_SRGB */
            auto dim = image.dim();
            Tex t;
            t.w = dim.first, t.h = dim.second;
            t.rgba.assign(image.data(), image.data() + (size_t)t.w * t.h * 4);
            texs.push_back(std::move(t));
        }
        mats_version++;
        reset_frame();
    }
    void destroy() {
        if(pipe) gpurt_pipe_destroy(pipe);
        pipe = nullptr, bound.reset(), bound_generation = 0;
    }
    void recreate_swap(const Scene&) {} /* the target image follows the extent passed to trace() */

    /* rt.cpp:121-138 */
    void update_uniforms(const Camera& cam) {
        Mat4 V = cam.get_view(), P = cam.get_proj();
        Mat4 iV = V.inverse(), iP = P.inverse();
        std::memset(&cam_ubo, 0, sizeof(cam_ubo));
        std::memcpy(cam_ubo.V, &V, 64), std::memcpy(cam_ubo.P, &P, 64);
        std::memcpy(cam_ubo.iV, &iV, 64), std::memcpy(cam_ubo.iP, &iP, 64);
        have_cam = true;
    }
    /* rt.cpp:140-176: bind the TLAS (and, here, bring its Scene_Desc / textures in line with recreate(scene)) */
    void use_accel(const Accel& tlas) {
        if(!tlas.is_top()) throw DropinError(GPURT_E_INVALID, "RTPipe::use_accel: not a top-level Accel");
        const bool new_accel = !bound || bound->uid != tlas.share()->uid;
        if(new_accel || applied_version != mats_version || bound_generation != tlas.generation()) {
            uint32_t n_objs = 0;
            dropin_check(gpurt_scene_counts(tlas.scene(), &n_objs, nullptr, nullptr, nullptr));
            if(n_objs != mats.size()) throw DropinError(GPURT_E_STATE, "RTPipe: recreate(scene) saw a different object count than the TLAS");
            dropin_check(gpurt_scene_clear_textures(tlas.scene()));
            for(const Tex& t : texs) dropin_check(gpurt_scene_add_texture(tlas.scene(), t.rgba.data(), t.w, t.h, nullptr));
            for(size_t i = 0; i < mats.size(); i++) dropin_check(gpurt_scene_set_material(tlas.scene(), (uint32_t)i, &mats[i]));
            dropin_check(gpurt_accel_sync_scene(tlas.handle()));
            applied_version = mats_version, bound_generation = tlas.generation();
        }
        if(new_accel) {
            if(pipe) gpurt_pipe_destroy(pipe);
            pipe = nullptr;
            dropin_check(gpurt_pipe_create(tlas.scene(), tlas.handle(), &pipe));
            bound = tlas.share();
        }
    }
    template <typename View> void use_image(const View&) {} /* rt_target lives in the pipe: read_image() / device_image() */
    void reset_frame() { /* rt.cpp:396-398 */
        if(pipe) dropin_check(gpurt_pipe_reset_frame(pipe));
        pending_reset = !pipe;
    }

    /* rt.cpp:346-394 */
    bool trace(const Camera& cam, VkCommandBuffer&, VkExtent2D ext) {
        if(!pipe) throw DropinError(GPURT_E_STATE, "RTPipe::trace before use_accel");
        if(!have_cam) update_uniforms(cam);
        if(pending_reset) dropin_check(gpurt_pipe_reset_frame(pipe)), pending_reset = false;
        GpurtPipeParams p = params();
        int rc = dropin_check(gpurt_pipe_render_frame(pipe, &p, &cam_ubo, ext.width, ext.height));
        w_ = ext.width, h_ = ext.height;
        return rc == 0;
    }

    /* rt.h:38-53 */
    int max_frames = 256;
    int samples_per_frame = 8;
    int max_depth = 8;
    Vec3 clear = Vec3{0.3f};
    Vec3 env = Vec3{1.0f};
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
    unsigned seed = 0; /* stands in for clockARB() in the per-pixel seed (rt.rgen:569), see DESIGN.md Q1 */
    int spatial_samples = 0;     /* extension: ReSTIR spatial reuse (include/gpurt.h), off by default */
    float spatial_radius = 16.0f;
    int light_sampling = 0;      /* extension: 1 = light triangles chosen in proportion to their power (include/gpurt.h) */

    /* ---- results (the reference samples rt_target in EffectPipe::tonemap and reads the framebuffer in save_rt) ---- */
    std::vector<float> read_image() const { /* RGBA32F, w*h*4 */
        std::vector<float> img((size_t)w_ * h_ * 4);
        dropin_check(gpurt_pipe_read_image(pipe, img.data(), GPURT_MEM_HOST));
        return img;
    }
    /* the same copy queued behind the frames traced so far; it overlaps the following trace() calls (page-locked `out`) */
    void read_image_async(float* out) const { dropin_check(gpurt_pipe_read_image_async(pipe, out)); }
    void read_image_wait() const { dropin_check(gpurt_pipe_read_image_wait(pipe)); }
    const float* device_image() const {
        void* p = nullptr;
        dropin_check(gpurt_pipe_device_image(pipe, &p));
        return static_cast<const float*>(p);
    }
    /* EffectPipe::tonemap (src/vk/effect.cpp:32-61, tonemap.frag) + the sRGB framebuffer store */
    std::vector<uint8_t> tonemap(int op = 1, float exposure = 1.0f, float gamma = 2.2f) const {
        std::vector<uint8_t> out((size_t)w_ * h_ * 4);
        dropin_check(gpurt_tonemap(pipe, op, exposure, gamma, out.data(), GPURT_MEM_HOST));
        return out;
    }
    int frame() const {
        int32_t f = -1;
        if(pipe) dropin_check(gpurt_pipe_frame_index(pipe, &f));
        return f;
    }
    gpurt_pipe* handle() const { return pipe; }

    GpurtPipeParams params() const {
        GpurtPipeParams p;
        gpurt_pipe_params_default(&p);
        p.max_frames = max_frames, p.samples_per_frame = samples_per_frame, p.max_depth = max_depth;
        p.clear[0] = clear.x, p.clear[1] = clear.y, p.clear[2] = clear.z;
        p.env[0] = env.x, p.env[1] = env.y, p.env[2] = env.z;
        p.env_scale = env_scale;
        p.use_normal_map = use_normal_map, p.use_rr = use_rr, p.use_metalness = use_metalness, p.use_qmc = use_qmc;
        p.use_temporal = use_temporal, p.integrator = integrator, p.temporal_scale = temporal_scale, p.brdf = brdf;
        p.debug_view = debug_view, p.res_samples = res_samples, p.seed = seed;
        p.spatial_samples = spatial_samples, p.spatial_radius = spatial_radius, p.light_sampling = light_sampling;
        return p;
    }

private:
    struct Tex {
        uint32_t w = 0, h = 0;
        std::vector<uint8_t> rgba;
    };
    std::vector<GpurtMaterial> mats;
    std::vector<Tex> texs;
    uint64_t mats_version = 0, applied_version = ~0ull, bound_generation = 0;
    gpurt_pipe* pipe = nullptr;
    std::shared_ptr<const Accel::Top> bound; /* keeps the TLAS's device build alive while the pipe uses it */
    GpurtCamera cam_ubo;
    bool have_cam = false, pending_reset = false;
    unsigned w_ = 0, h_ = 0;
};

} // namespace VK

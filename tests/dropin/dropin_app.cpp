/*
 * dropin_app.cpp — TEST INFRASTRUCTURE: the reference application's ray-tracing call sites, compiled against
 * gpu-rt_b200/host/vk_dropin.h.  `App` holds the members of class GPURT the path uses (src/gpurt.h:44-60) and the
 * bodies of GPURT::build_rt (src/gpurt.cpp:216-218), GPURT::build_accel (:220-241) and the use_rt branch of
 * GPURT::render (:39-45) as they stand in the reference — those statements are what "drop-in" has to keep compiling.
 *
 *   default              types from tests/dropin/app_types.h (no reference tree needed)
 *   -DDROPIN_REAL_TYPES  Scene / Object / VK::Mesh / Camera / Mat4 are the reference's own headers and code
 *                        (oracle/Makefile builds it next to oracle/_ref/libgpurt_ref.so; the reference's Vulkan-backed
 *                        Accel / RTPipe / Drop declarations are renamed out of the way for this translation unit)
 *
 * usage: dropin_app scene.gltf out.f32 W H frames integrator brdf spp depth seed [edit_obj dx]
 * Renders `frames` calls of render(); with edit_obj >= 0 the pose of the edit_obj-th object (for_objs order) is then
 * moved by dx along x — GPURT::edit_scene's pose edit, which sets rebuild_tlas — and `frames` more calls follow.
 * Writes the RGBA32F image.
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#ifdef DROPIN_REAL_TYPES
#define Accel RefVkAccel
#define RTPipe RefVkRTPipe
#define Drop RefVkDrop
#include <scene/scene.h>
#include <util/camera.h>
#undef Accel
#undef RTPipe
#undef Drop
#define GPURT_DROPIN_HAVE_VK_HANDLES
namespace VK {
/* link-time stand-ins for the renamed Vulkan-backed class (never instantiated; cf. oracle/ref_shim/vk_stubs.cpp) */
RefVkAccel::~RefVkAccel() {}
RefVkAccel::RefVkAccel(RefVkAccel&&) {}
RefVkAccel& RefVkAccel::operator=(RefVkAccel&&) { return *this; }
struct DropinImage { /* stands in for rt_target (a VkImage in the reference): only w / h are read at the call site */
    unsigned int w = 0, h = 0;
};
struct DropinImageView {};
} // namespace VK
#define RT_IMAGE VK::DropinImage
#define RT_IMAGE_VIEW VK::DropinImageView
#else
#include "app_types.h"
#define RT_IMAGE VK::Image
#define RT_IMAGE_VIEW VK::ImageView
#endif

#include "../../gpu-rt_b200/host/vk_dropin.h"

struct App {
    /* src/gpurt.h:44-60 */
    Scene scene;
    Camera cam;
    bool use_rt = true, rebuild_blas = true, rebuild_tlas = true;
    std::vector<VK::Drop<VK::Accel>> BLAS;
    std::vector<Mat4> BLAS_T;
    VK::Drop<VK::Accel> TLAS;
    VK::RTPipe rt_pipe;
    VK::Drop<RT_IMAGE> rt_target;
    VK::Drop<RT_IMAGE_VIEW> rt_target_view;
    bool more = true;

    explicit App(Vec2 dim) : cam(dim) {}

    /* src/gpurt.cpp:216-218 */
    void build_rt() {
        rt_pipe.recreate(scene);
    }

    /* src/gpurt.cpp:220-241 */
    void build_accel() {

        if(rebuild_blas) {
            BLAS.clear();
            scene.for_objs([this](const Object& obj) { BLAS.push_back({VK::Accel(obj.mesh())}); });
            rebuild_blas = false;
        }

        if(rebuild_tlas) {

            BLAS_T.clear();
            scene.for_objs([this](const Object& obj) {
                BLAS_T.push_back(Mat4::scale(Vec3{scene.scale}) * obj.pose.transform());
            });

            TLAS.drop();
            TLAS->recreate(BLAS, BLAS_T);
            rebuild_tlas = false;

            rt_pipe.recreate(scene);
        }
    }

    /* src/gpurt.cpp:32-45 (use_rt branch; `cmds` comes from vk.begin() in the reference) */
    void render() {
        VkCommandBuffer cmds = {};

        if(use_rt) {
            build_accel();

            rt_pipe.use_image(rt_target_view);
            rt_pipe.use_accel(TLAS);
            rt_pipe.update_uniforms(cam);
            more = rt_pipe.trace(cam, cmds, {rt_target->w, rt_target->h});
        }
    }
};

int main(int argc, char** argv) {
    if(argc < 11) {
        std::fprintf(stderr, "usage: %s scene.gltf out.f32 W H frames integrator brdf spp depth seed [edit_obj dx]\n", argv[0]);
        return 2;
    }
    try {
        unsigned w = (unsigned)std::atoi(argv[3]), h = (unsigned)std::atoi(argv[4]);
        int frames = std::atoi(argv[5]);
        App app(Vec2{(float)w, (float)h});
        std::string err = app.scene.load(argv[1], app.cam);
        if(!err.empty()) {
            std::fprintf(stderr, "load: %s\n", err.c_str());
            return 3;
        }
        app.rt_target->w = w, app.rt_target->h = h;
        app.rt_pipe.integrator = std::atoi(argv[6]), app.rt_pipe.brdf = std::atoi(argv[7]);
        app.rt_pipe.samples_per_frame = std::atoi(argv[8]), app.rt_pipe.max_depth = std::atoi(argv[9]);
        app.rt_pipe.seed = (unsigned)std::strtoul(argv[10], nullptr, 10);
        app.rt_pipe.max_frames = 1 << 20;
        for(int f = 0; f < frames; f++) app.render();
        if(argc >= 13 && std::atoi(argv[11]) >= 0) {
            int k = std::atoi(argv[11]), i = 0;
            float dx = (float)std::atof(argv[12]);
            app.scene.for_objs([&](Object& obj) {
                if(i++ == k) obj.pose.pos.x += dx;
            });
            app.rebuild_tlas = true; /* GPURT::edit_scene, src/gpurt.cpp:286-289 */
            for(int f = 0; f < frames; f++) app.render();
        }
        std::vector<float> img = app.rt_pipe.read_image();
        FILE* fp = std::fopen(argv[2], "wb");
        if(!fp || std::fwrite(img.data(), 4, img.size(), fp) != img.size()) return 4;
        std::fclose(fp);
        std::printf("frame %d, %zu objects, %ux%u\n", app.rt_pipe.frame(), app.BLAS.size(), w, h);
    } catch(const VK::DropinError& e) {
        std::fprintf(stderr, "dropin error %d: %s\n", e.code, e.what());
        return e.code == GPURT_E_NO_DEVICE ? 42 : 1;
    }
    VK::Accel::release_deferred();
    return 0;
}

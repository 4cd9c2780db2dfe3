/*
 * main_render.cpp — headless front end replacing GPURT::loop (src/gpurt.cpp:411-435) and main.cpp's
 * `-s scene` option (src/main.cpp:6-18): load a scene, build the acceleration structure, render
 * max_frames progressive frames, tonemap and write a PNG (GPURT::save_rt, src/gpurt.cpp:258-262).
 * Every RTPipe tunable the reference exposes through ImGui (src/gpurt.cpp:283-314) is a flag.
 *
 *   gpurt_render -s media/cbox/cbox.gltf -o out.png --frames 16 --spp 8 --integrator 2
 */
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rtpipe.h"

static void put32(std::vector<uint8_t>& v, uint32_t x) {
    for(int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s));
}
static void chunk(std::vector<uint8_t>& png, const char* tag, const std::vector<uint8_t>& data) {
    put32(png, (uint32_t)data.size());
    size_t start = png.size();
    png.insert(png.end(), tag, tag + 4);
    png.insert(png.end(), data.begin(), data.end());
    put32(png, (uint32_t)crc32(0, png.data() + start, (uInt)(png.size() - start)));
}
static bool write_png(const std::string& path, const std::vector<uint8_t>& rgba, unsigned w, unsigned h) {
    std::vector<uint8_t> raw;
    raw.reserve((size_t)(w * 4 + 1) * h);
    for(unsigned y = 0; y < h; y++) {
        raw.push_back(0);
        raw.insert(raw.end(), rgba.begin() + (size_t)y * w * 4, rgba.begin() + (size_t)(y + 1) * w * 4);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if(compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    comp.resize(clen);
    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
    put32(ihdr, w), put32(ihdr, h);
    ihdr.insert(ihdr.end(), {8, 6, 0, 0, 0});
    chunk(png, "IHDR", ihdr);
    chunk(png, "IDAT", comp);
    chunk(png, "IEND", {});
    FILE* f = fopen(path.c_str(), "wb");
    if(!f) return false;
    fwrite(png.data(), 1, png.size(), f);
    fclose(f);
    return true;
}

int main(int argc, char** argv) {
    std::string scene_file, out = "out.png";
    unsigned w = 1280, h = 720; /* src/platform/window.cpp:36-38 */
    int device = 0, tonemap_op = 1, cam_mode = 0;
    float exposure = 1.0f, gamma = 2.2f, scale = 1.0f, vfov = 90.0f;
    float pos[3] = {0, 0, 0}, at[3] = {0, 0, 0};
    bool standin = false;
    struct Opt {
        int max_frames = 256, spp = 8, depth = 8, integrator = 0, brdf = 0, rr = 1, qmc = 0, temporal = 1,
            temporal_scale = 16, res_samples = 4, normal_map = 0, metalness = 0, debug_view = 0;
        float clear = 0.3f, env_scale = 0.0f;
        unsigned seed = 0;
    } o;
    for(int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : "0"; };
        if(a == "-s" || a == "--scene") scene_file = next();
        else if(a == "--sponza-standin") standin = true;
        else if(a == "-o") out = next();
        else if(a == "--size") w = (unsigned)atoi(next()), h = (unsigned)atoi(next());
        else if(a == "--device") device = atoi(next());
        else if(a == "--frames") o.max_frames = atoi(next());
        else if(a == "--spp") o.spp = atoi(next());
        else if(a == "--depth") o.depth = atoi(next());
        else if(a == "--integrator") o.integrator = atoi(next());
        else if(a == "--brdf") o.brdf = atoi(next());
        else if(a == "--no-rr") o.rr = 0;
        else if(a == "--qmc") o.qmc = 1;
        else if(a == "--no-temporal") o.temporal = 0;
        else if(a == "--temporal-scale") o.temporal_scale = atoi(next());
        else if(a == "--res-samples") o.res_samples = atoi(next());
        else if(a == "--normal-map") o.normal_map = 1;
        else if(a == "--metalness") o.metalness = 1;
        else if(a == "--debug-view") o.debug_view = atoi(next());
        else if(a == "--clear") o.clear = (float)atof(next());
        else if(a == "--env-scale") o.env_scale = (float)atof(next());
        else if(a == "--seed") o.seed = (unsigned)atoi(next());
        else if(a == "--scale") scale = (float)atof(next());
        else if(a == "--tonemap") tonemap_op = atoi(next());
        else if(a == "--exposure") exposure = (float)atof(next());
        else if(a == "--gamma") gamma = (float)atof(next());
        else if(a == "--camera") {
            cam_mode = 1;
            for(int k = 0; k < 3; k++) pos[k] = (float)atof(next());
            for(int k = 0; k < 3; k++) at[k] = (float)atof(next());
            vfov = (float)atof(next());
        } else {
            fprintf(stderr, "unknown option %s\n", a.c_str());
            return 2;
        }
    }
    if(scene_file.empty() && !standin) {
        fprintf(stderr, "usage: gpurt_render -s scene.gltf|--sponza-standin [-o out.png|out.exr] [--size W H] [--frames N] [--spp N] "
                        "[--depth N] [--integrator 0..4] [--brdf 0|1] [--camera px py pz ax ay az vfov] ...\n");
        return 2;
    }
    try {
        gpurt::Context ctx(device);
        gpurt::SceneHandle scene(ctx);
        if(standin) scene.make_sponza_standin();
        else scene.load(scene_file, scale);
        auto t0 = std::chrono::steady_clock::now();
        gpurt::Accel accel(scene);
        GpurtAccelInfo info = accel.info();
        gpurt::RTPipe pipe(scene, accel);
        pipe.max_frames = o.max_frames, pipe.samples_per_frame = o.spp, pipe.max_depth = o.depth;
        pipe.integrator = o.integrator, pipe.brdf = o.brdf, pipe.use_rr = o.rr, pipe.use_qmc = o.qmc;
        pipe.use_temporal = o.temporal, pipe.temporal_scale = o.temporal_scale, pipe.res_samples = o.res_samples;
        pipe.use_normal_map = o.normal_map, pipe.use_metalness = o.metalness, pipe.debug_view = o.debug_view;
        pipe.clear[0] = pipe.clear[1] = pipe.clear[2] = o.clear;
        pipe.env_scale = o.env_scale, pipe.seed = o.seed;
        GpurtCamera cam;
        gpurt::check(gpurt_camera_make(cam_mode, (float)w, (float)h, pos, at, vfov, &cam));
        auto t1 = std::chrono::steady_clock::now();
        int frames = 0;
        while(pipe.trace(cam, w, h)) frames++; /* GPURT::render until converged (rt.cpp:353) */
        auto t2 = std::chrono::steady_clock::now();
        if(out.size() > 4 && out.compare(out.size() - 4, 4, ".exr") == 0) { /* linear radiance, untouched by the tonemap pass */
            auto lin = pipe.read_image();
            t2 = std::chrono::steady_clock::now();
            gpurt::check(gpurt_write_exr(out.c_str(), lin.data(), w, h));
        } else {
            auto img = pipe.tonemap(tonemap_op, exposure, gamma);
            t2 = std::chrono::steady_clock::now();
            if(!write_png(out, img, w, h)) throw std::runtime_error("cannot write " + out);
        }
        printf("%u tris, %u wide nodes (depth %u), build %.2f ms; %d frames x %d spp at %ux%u in %.1f ms -> %s\n", info.n_tris,
               info.n_wide_nodes, info.wide_depth, info.build_ms, frames, o.spp, w, h,
               std::chrono::duration<double, std::milli>(t2 - t1).count(), out.c_str());
        (void)t0;
    } catch(const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}

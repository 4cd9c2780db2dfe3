/*
 * fuzz_loader.cpp — mutation fuzzer for the host half (glTF / GLB loader, JSON reader, PNG and JPEG decoders, scene packing),
 * meant to be built with AddressSanitizer + UndefinedBehaviorSanitizer (tools/fuzz_loader.sh).  No GPU, no CUDA: it links
 * host/*.cpp only and goes through the C ABI the way a caller would (gpurt_scene_load_gltf, then every read-out entry point).
 *
 *   fuzz_loader <seed file> <iterations> [rng seed]
 *
 * Seeds: .gltf (text mutations: numeric tokens replaced by hostile values, spans deleted / duplicated, bytes flipped, file
 * truncated), .glb / .png / .jpg (binary mutations; images are wrapped in a one-triangle glTF that references them).
 * A run passes when the sanitizers stay silent and every call returns (an error code is a fine answer).
 */
#include "../include/gpurt.h"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
    g_rng ^= g_rng << 13, g_rng ^= g_rng >> 7, g_rng ^= g_rng << 17;
    return g_rng;
}
static size_t rnd_below(size_t n) { return n ? (size_t)(rnd() % n) : 0; }

static std::vector<uint8_t> read_file(const std::string& p) {
    std::vector<uint8_t> out;
    if(FILE* f = fopen(p.c_str(), "rb")) {
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        out.resize(n > 0 ? (size_t)n : 0);
        if(n > 0 && fread(out.data(), 1, (size_t)n, f) != (size_t)n) out.clear();
        fclose(f);
    }
    return out;
}
static void write_file(const std::string& p, const std::vector<uint8_t>& d) {
    if(FILE* f = fopen(p.c_str(), "wb")) {
        if(!d.empty()) fwrite(d.data(), 1, d.size(), f);
        fclose(f);
    }
}

static const char* kHostile[] = {"-1", "0", "1", "255", "65535", "65536", "2147483647", "2147483648", "4294967295", "4294967296",
                                 "9223372036854775807", "18446744073709551616", "1e30", "-1e30", "1e308", "1e-320", "1e999",
                                 "-0.0", "0.5", "3", "7", "12", "5120", "5121", "5123", "5125", "5126", "5130", "4", "5", "6",
                                 "null", "true", "[]", "{}", "\"\"", "NaN"};

static void mutate_binary(std::vector<uint8_t>& d) {
    if(d.empty()) return;
    int n = 1 + (int)rnd_below(6);
    for(int k = 0; k < n; k++) {
        size_t i = rnd_below(d.size());
        switch(rnd_below(6)) {
        case 0: d[i] ^= (uint8_t)(1u << rnd_below(8)); break;
        case 1: d[i] = (uint8_t)rnd(); break;
        case 2: d[i] = (rnd() & 1) ? 0xff : 0x00; break;
        case 3: /* 32-bit field */
            if(i + 4 <= d.size()) {
                static const uint32_t v[] = {0, 1, 0x7fffffffu, 0x80000000u, 0xffffffffu, 0xfffffff0u, 0x10000u};
                uint32_t x = v[rnd_below(7)];
                memcpy(&d[i], &x, 4);
            }
            break;
        case 4: d.resize(i); break; /* truncate */
        case 5: { /* duplicate a span */
            size_t len = 1 + rnd_below(64);
            if(i + len <= d.size()) {
                std::vector<uint8_t> span(d.begin() + (long)i, d.begin() + (long)(i + len));
                d.insert(d.begin() + (long)i, span.begin(), span.end());
            }
        } break;
        }
        if(d.empty()) return;
    }
}

static void mutate_text(std::vector<uint8_t>& d) {
    if(d.empty()) return;
    int n = 1 + (int)rnd_below(4);
    for(int k = 0; k < n && !d.empty(); k++) {
        if(rnd_below(4) == 0) {
            mutate_binary(d);
            continue;
        }
        if(rnd_below(4) == 0) { /* type confusion: a whole [...] / {...} / "..." value becomes something else */
            size_t b = rnd_below(d.size());
            while(b < d.size() && d[b] != '[' && d[b] != '{' && d[b] != '"') b++;
            if(b >= d.size()) continue;
            size_t e = b + 1;
            if(d[b] == '"') {
                while(e < d.size() && d[e] != '"') e += d[e] == '\\' ? 2 : 1;
                e = e < d.size() ? e + 1 : d.size();
            } else {
                int level = 1;
                for(; e < d.size() && level; e++) level += (d[e] == '[' || d[e] == '{') - (d[e] == ']' || d[e] == '}');
            }
            if(e > d.size()) e = d.size();
            static const char* repl[] = {"null", "0", "-1", "[]", "{}", "\"x\"", "[[]]", "{\"a\":{}}", "1e30", "true"};
            const char* h = repl[rnd_below(sizeof(repl) / sizeof(repl[0]))];
            d.erase(d.begin() + (long)b, d.begin() + (long)e);
            d.insert(d.begin() + (long)b, (const uint8_t*)h, (const uint8_t*)h + strlen(h));
            continue;
        }
        /* find a numeric token starting at a random position and replace it */
        size_t start = rnd_below(d.size()), i = start;
        auto is_num = [](uint8_t c) { return (c >= '0' && c <= '9') || c == '-' || c == '.' || c == 'e' || c == 'E' || c == '+'; };
        while(i < d.size() && !(d[i] >= '0' && d[i] <= '9')) i++;
        if(i >= d.size()) continue;
        size_t b = i, e = i;
        while(b > 0 && is_num(d[b - 1])) b--;
        while(e < d.size() && is_num(d[e])) e++;
        const char* h = kHostile[rnd_below(sizeof(kHostile) / sizeof(kHostile[0]))];
        d.erase(d.begin() + (long)b, d.begin() + (long)e);
        d.insert(d.begin() + (long)b, (const uint8_t*)h, (const uint8_t*)h + strlen(h));
    }
}

/* every read-out entry point of a loaded scene: packing, light list, object copies, textures */
static void exercise(gpurt_scene* s) {
    uint32_t no = 0, nt = 0, nl = 0, ntex = 0;
    if(gpurt_scene_counts(s, &no, &nt, &nl, &ntex) != GPURT_OK) return;
    if(no) {
        std::vector<uint32_t> off(no + 1);
        gpurt_scene_tri_offsets(s, off.data());
        std::vector<GpurtSceneDesc> descs(no);
        gpurt_scene_get_descs(s, descs.data());
    }
    if(nl) {
        std::vector<GpurtSceneLight> lights(nl);
        gpurt_scene_get_lights(s, lights.data());
    }
    for(uint32_t o = 0; o < no && o < 64; o++) {
        uint32_t nv = 0, ni = 0;
        if(gpurt_scene_object_sizes(s, o, &nv, &ni) != GPURT_OK) continue;
        std::vector<uint8_t> verts((size_t)nv * 48);
        std::vector<uint32_t> idx(ni);
        gpurt_scene_get_object(s, o, verts.data(), idx.data());
    }
    for(uint32_t t = 0; t < ntex && t < 16; t++) {
        uint32_t w = 0, h = 0;
        if(gpurt_scene_get_texture(s, t, &w, &h, nullptr) != GPURT_OK || (uint64_t)w * h > (64u << 20)) continue;
        std::vector<uint8_t> px((size_t)w * h * 4);
        gpurt_scene_get_texture(s, t, &w, &h, px.data());
    }
}

int main(int argc, char** argv) {
    if(argc < 3) return fprintf(stderr, "usage: %s <seed file> <iterations> [rng seed]\n", argv[0]), 2;
    std::string seed_path = argv[1];
    long iters = atol(argv[2]);
    if(argc > 3) g_rng ^= strtoull(argv[3], nullptr, 0) * 0xD1342543DE82EF95ull + 1;
    std::vector<uint8_t> seed = read_file(seed_path);
    if(seed.empty()) return fprintf(stderr, "cannot read %s\n", seed_path.c_str()), 2;
    std::string ext = seed_path.substr(seed_path.find_last_of('.') + 1);
    std::string seed_dir = seed_path.substr(0, seed_path.find_last_of('/') + 1);

    char tmpl[] = "/tmp/gpurt_fuzz_XXXXXX";
    if(!mkdtemp(tmpl)) return 2;
    std::string dir = std::string(tmpl) + "/";
    bool image = ext == "png" || ext == "jpg" || ext == "jpeg";
    std::string target = dir + (image ? "scene.gltf" : "scene." + ext);
    if(image) {
        /* one triangle whose material samples the image under test */
        std::string g = "{\"asset\":{\"version\":\"2.0\"},\"buffers\":[{\"byteLength\":48,\"uri\":\"data:application/octet-stream;base64,"
                        "AAAAAAAAAAAAAAAAAACAPwAAAAAAAAAAAAAAAAAAgD8AAAAAAAAAAAEAAAACAAAA\"}],"
                        "\"bufferViews\":[{\"buffer\":0,\"byteOffset\":0,\"byteLength\":36},{\"buffer\":0,\"byteOffset\":36,\"byteLength\":12}],"
                        "\"accessors\":[{\"bufferView\":0,\"componentType\":5126,\"count\":3,\"type\":\"VEC3\"},"
                        "{\"bufferView\":1,\"componentType\":5125,\"count\":3,\"type\":\"SCALAR\"}],"
                        "\"images\":[{\"uri\":\"img." + ext + "\"}],\"textures\":[{\"source\":0}],"
                        "\"materials\":[{\"pbrMetallicRoughness\":{\"baseColorTexture\":{\"index\":0}}}],"
                        "\"meshes\":[{\"primitives\":[{\"attributes\":{\"POSITION\":0},\"indices\":1,\"material\":0}]}],"
                        "\"nodes\":[{\"mesh\":0}],\"scenes\":[{\"nodes\":[0]}]}";
        write_file(target, std::vector<uint8_t>(g.begin(), g.end()));
    } else if(ext == "gltf") {
        /* external buffers / images the seed names sit next to it: link them into the scratch directory */
        std::string cmd = "for f in \"$(realpath '" + (seed_dir.empty() ? std::string("./") : seed_dir) + "')\"/*; do ln -sf \"$f\" '" + dir + "'; done";
        if(system(cmd.c_str()) != 0) return 2;
        unlink(target.c_str());
    }

    long ok = 0, rejected = 0;
    for(long it = 0; it < iters; it++) {
        std::vector<uint8_t> d = seed;
        if(it) {
            if(ext == "gltf") mutate_text(d);
            else mutate_binary(d);
        }
        write_file(image ? dir + "img." + ext : target, d);
        gpurt_scene* s = nullptr;
        if(gpurt_scene_create(nullptr, &s) != GPURT_OK) return 1;
        int rc = gpurt_scene_load_gltf(s, target.c_str(), 1.0f);
        if(rc == GPURT_OK) ok++, exercise(s);
        else rejected++;
        gpurt_scene_destroy(s);
        if(it == 0 && rc != GPURT_OK) return fprintf(stderr, "unmutated seed rejected: %s\n", gpurt_last_error()), 1;
    }
    printf("%s: %ld iterations, %ld loaded, %ld rejected\n", seed_path.c_str(), iters, ok, rejected);
    std::string rm = "rm -rf '" + dir + "'";
    return system(rm.c_str()) != 0;
}

/*
 * glsl_tonemap_ref.cpp — the reference's tonemap.frag compiled as C++ (oracle/_ref/tonemap_gen.inc, see
 * glsl_compat.h), followed by what the Vulkan framebuffer does with the shader's output in the reference
 * (R8G8B8A8_SRGB target, src/gpurt.cpp:176: clamp, sRGB encode of rgb, round to 8 bits).  TEST INFRASTRUCTURE:
 * part of oracle/_ref/libglsl_ref.so, pins orc_tonemap / gpurt_tonemap.
 */
#include <cstring>

#include "glsl_compat.h"

namespace glsl_tonemap {
using namespace glsl;
static vec2 fragTexcoord;
static vec4 outColor;
static sampler2D image;
static vec4 g_texel;
static vec4 texture(sampler2D, vec2) { return g_texel; }
#include "../_ref/tonemap_gen.inc"
} // namespace glsl_tonemap

extern "C" void ref_glsl_tonemap(const float* rgba, unsigned long long n, int op, float exposure, float gamma, uint8_t* out) {
    using namespace glsl_tonemap;
    consts.exposure = exposure, consts.gamma = gamma, consts.type = op;
    for(unsigned long long i = 0; i < n; i++) {
        g_texel = glsl::vec4(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]);
        outColor = glsl::vec4(0, 0, 0, 0);
        rgen_main(); /* tonemap.frag main() (renamed by the generator) */
        const float o[4] = {outColor.x, outColor.y, outColor.z, outColor.w};
        for(int k = 0; k < 4; k++) { /* framebuffer store: NaN -> 0, clamp, sRGB encode (rgb), round */
            float x = o[k];
            x = x != x ? 0.0f : fminf(fmaxf(x, 0.0f), 1.0f);
            if(k < 3) x = x <= 0.0031308f ? 12.92f * x : 1.055f * dm_pow(x, 1.0f / 2.4f) - 0.055f;
            out[4 * i + k] = (uint8_t)(x * 255.0f + 0.5f);
        }
    }
}

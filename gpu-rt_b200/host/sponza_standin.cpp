/*
 * sponza_standin.cpp — LABELLED procedural stand-in for media/sponza.
 *
 * The reference snapshot ships Sponza.gltf and its textures but not Sponza.bin
 * (.MISSING_LARGE_BLOBS:3), so the geometry cannot be loaded.  This generator builds an atrium with
 * the properties the traversal benchmark depends on, taken from Sponza.gltf's accessors: exactly
 * 262,267 triangles in 103 objects, object-space extent [-1921,-126,-1183]..[1800,1429,1105],
 * identity model matrices (the glTF node's TRS scale is ignored by the loader), no
 * emitters, a mix of large flat surfaces (floor, walls) and dense small-triangle detail (columns,
 * arches, drapes, vases).  Every result produced on it is reported as "sponza_standin".
 */
#include <cmath>

#include "scene.h"

namespace gpurt {
namespace {

const float X0 = -1921.0f, X1 = 1800.0f, Y0 = -126.0f, Y1 = 1429.0f, Z0 = -1183.0f, Z1 = 1105.0f;
const float TWO_PI = 6.28318530717958647692f;

struct Builder {
    std::vector<Vertex> v;
    std::vector<uint32_t> i;
    uint32_t vert(Vec3 p, Vec3 n, float s, float t) {
        Vertex q;
        q.pos[0] = p.x, q.pos[1] = p.y, q.pos[2] = p.z, q.pos[3] = s;
        q.norm[0] = n.x, q.norm[1] = n.y, q.norm[2] = n.z, q.norm[3] = t;
        q.tang[0] = 1, q.tang[1] = 0, q.tang[2] = 0, q.tang[3] = 1;
        v.push_back(q);
        return (uint32_t)v.size() - 1;
    }
    /* nu x nv quads of a parametric surface f(s,t) -> (pos, normal) */
    template <typename F> void grid(int nu, int nv, F f) {
        uint32_t base = (uint32_t)v.size();
        for(int b = 0; b <= nv; b++)
            for(int a = 0; a <= nu; a++) {
                float s = (float)a / (float)nu, t = (float)b / (float)nv;
                Vec3 p, n;
                f(s, t, p, n);
                vert(p, n, s, t);
            }
        for(int b = 0; b < nv; b++)
            for(int a = 0; a < nu; a++) {
                uint32_t k = base + (uint32_t)(b * (nu + 1) + a);
                uint32_t w = (uint32_t)nu + 1;
                i.insert(i.end(), {k, k + 1, k + w + 1, k, k + w + 1, k + w});
            }
    }
    void fan(Vec3 c, Vec3 n, float r, int ntris, Vec3 ax, Vec3 az) {
        uint32_t cidx = vert(c, n, 0.5f, 0.5f);
        uint32_t first = (uint32_t)v.size();
        for(int k = 0; k <= ntris; k++) {
            float a = TWO_PI * (float)k / (float)ntris;
            vert(c + ax * (r * std::cos(a)) + az * (r * std::sin(a)), n, 0.5f + 0.5f * std::cos(a),
                 0.5f + 0.5f * std::sin(a));
        }
        for(int k = 0; k < ntris; k++) i.insert(i.end(), {cidx, first + (uint32_t)k, first + (uint32_t)k + 1});
    }
};

void emit(Scene& sc, Builder& b, int salt) {
    Object o;
    o.id = sc.reserve_id();
    o.mesh.set(std::move(b.v), std::move(b.i));
    /* deterministic, texture-free materials: albedo from a small hash, GGX-friendly roughness */
    unsigned h = (unsigned)salt * 2654435761u;
    o.material.albedo = Vec3{0.35f + 0.5f * (float)((h >> 8) & 255) / 255.0f,
                             0.35f + 0.5f * (float)((h >> 16) & 255) / 255.0f,
                             0.35f + 0.5f * (float)((h >> 24) & 255) / 255.0f};
    o.material.metal_rough = Vec2{0.0f, 0.25f + 0.5f * (float)(h & 255) / 255.0f};
    sc.add(std::move(o));
    b = Builder();
}

} // namespace

void make_sponza_standin(Scene& sc) {
    sc.clear();
    Builder b;
    int salt = 1;
    size_t tris = 0, objs = 0;
    auto done = [&] {
        tris += b.i.size() / 3;
        objs++;
        emit(sc, b, salt++);
    };
    auto plane_y = [&](float y, float ny, int nu, int nv, float x0, float x1, float z0, float z1) {
        b.grid(nu, nv, [&](float s, float t, Vec3& p, Vec3& n) {
            p = Vec3{x0 + (x1 - x0) * s, y, z0 + (z1 - z0) * t};
            n = Vec3{0, ny, 0};
        });
    };
    /* 1 floor: 60x40 quads = 4,800 tris; 1 roof of the same count, open over the light-well like the
     * real atrium (4 strips of 50x12 quads), so environment light reaches the nave */
    plane_y(Y0, 1.0f, 60, 40, X0, X1, Z0, Z1);
    done();
    plane_y(Y1, -1.0f, 50, 12, X0, X1, Z0, -420.0f);
    plane_y(Y1, -1.0f, 50, 12, X0, X1, 420.0f, Z1);
    plane_y(Y1, -1.0f, 50, 12, X0, -1300.0f, -420.0f, 420.0f);
    plane_y(Y1, -1.0f, 50, 12, 1180.0f, X1, -420.0f, 420.0f);
    done();
    /* 4 outer walls: 40x20 quads = 1,600 tris each */
    for(int w = 0; w < 4; w++) {
        b.grid(40, 20, [&](float s, float t, Vec3& p, Vec3& n) {
            float y = Y0 + (Y1 - Y0) * t;
            if(w == 0) p = Vec3{X0 + (X1 - X0) * s, y, Z0}, n = Vec3{0, 0, 1};
            if(w == 1) p = Vec3{X0 + (X1 - X0) * s, y, Z1}, n = Vec3{0, 0, -1};
            if(w == 2) p = Vec3{X0, y, Z0 + (Z1 - Z0) * s}, n = Vec3{1, 0, 0};
            if(w == 3) p = Vec3{X1, y, Z0 + (Z1 - Z0) * s}, n = Vec3{-1, 0, 0};
        });
        done();
    }
    /* 1 upper gallery slab with a rectangular light-well cut by four strips: 4 x (50x10 quads) */
    {
        float y = Y0 + 560.0f;
        plane_y(y, -1.0f, 50, 10, X0, X1, Z0, -420.0f);
        plane_y(y, -1.0f, 50, 10, X0, X1, 420.0f, Z1);
        plane_y(y, -1.0f, 50, 10, X0, -1300.0f, -420.0f, 420.0f);
        plane_y(y, -1.0f, 50, 10, 1180.0f, X1, -420.0f, 420.0f);
        done();
    }
    /* 48 columns: 2 rows x 12 x 2 storeys; 32 segments x 24 rings + 2 caps = 1,600 tris */
    for(int storey = 0; storey < 2; storey++)
        for(int row = 0; row < 2; row++)
            for(int c = 0; c < 12; c++) {
                float cx = -1300.0f + (float)c * (2480.0f / 11.0f);
                float cz = row ? 420.0f : -420.0f;
                float y0 = Y0 + (storey ? 560.0f : 0.0f), hgt = storey ? 480.0f : 560.0f;
                float r0 = storey ? 42.0f : 58.0f;
                b.grid(32, 24, [&](float s, float t, Vec3& p, Vec3& n) {
                    float a = TWO_PI * s;
                    float r = r0 * (1.0f + 0.18f * std::cos(TWO_PI * t) * (t < 0.12f || t > 0.88f ? 1.0f : 0.15f));
                    n = Vec3{std::cos(a), 0, std::sin(a)};
                    p = Vec3{cx + r * n.x, y0 + hgt * t, cz + r * n.z};
                });
                b.fan(Vec3{cx, y0, cz}, Vec3{0, -1, 0}, r0 * 1.18f, 32, Vec3{1, 0, 0}, Vec3{0, 0, 1});
                b.fan(Vec3{cx, y0 + hgt, cz}, Vec3{0, 1, 0}, r0 * 1.18f, 32, Vec3{1, 0, 0}, Vec3{0, 0, 1});
                done();
            }
    /* 24 arches between ground-storey columns (half tori): 32x16 quads = 1,024 tris */
    for(int row = 0; row < 2; row++)
        for(int c = 0; c < 12; c++) {
            float pitch = 2480.0f / 11.0f;
            float cx = -1300.0f + ((float)c + (c < 11 ? 0.5f : -0.5f)) * pitch;
            float cz = (row ? 420.0f : -420.0f) + (c == 11 ? (row ? 60.0f : -60.0f) : 0.0f);
            float R = 0.5f * pitch - 20.0f, r = 24.0f, y = Y0 + 430.0f;
            b.grid(32, 16, [&](float s, float t, Vec3& p, Vec3& n) {
                float A = 3.14159265f * s, a = TWO_PI * t;
                Vec3 ring{std::cos(A), std::sin(A), 0};
                n = Vec3{ring.x * std::cos(a), ring.y * std::cos(a), std::sin(a)};
                p = Vec3{cx + ring.x * R + n.x * r, y + ring.y * R + n.y * r, cz + n.z * r};
            });
            done();
        }
    /* 12 drapes: wavy sheets hanging in the light-well, 64x64 quads = 8,192 tris */
    for(int d = 0; d < 12; d++) {
        float cx = -1100.0f + (float)(d % 6) * 420.0f;
        float cz = d < 6 ? -300.0f : 300.0f;
        float ph = 0.7f * (float)d;
        b.grid(64, 64, [&](float s, float t, Vec3& p, Vec3& n) {
            float x = cx + 300.0f * (s - 0.5f);
            float y = Y0 + 540.0f + 520.0f * t;
            float wv = 28.0f * std::sin(18.0f * s + ph) * (1.0f - 0.6f * t) + 12.0f * std::sin(7.0f * t + ph);
            float dz = 28.0f * 18.0f / 300.0f * std::cos(18.0f * s + ph) * (1.0f - 0.6f * t);
            p = Vec3{x, y, cz + wv};
            n = Vec3{-dz, 0, 1}.unit();
        });
        done();
    }
    /* 11 vases / busts: UV spheres 48x24 quads = 2,304 tris, squashed profile */
    for(int k = 0; k < 11; k++) {
        float cx = -1500.0f + (float)k * 300.0f, cz = (k & 1) ? 150.0f : -150.0f;
        float R = 55.0f + 6.0f * (float)(k % 4);
        b.grid(48, 24, [&](float s, float t, Vec3& p, Vec3& n) {
            float a = TWO_PI * s, e = 3.14159265f * (t - 0.5f);
            float prof = 1.0f + 0.35f * std::sin(3.0f * e + (float)k);
            n = Vec3{std::cos(e) * std::cos(a), std::sin(e), std::cos(e) * std::sin(a)};
            p = Vec3{cx + R * prof * n.x, Y0 + 1.6f * R + 1.6f * R * n.y, cz + R * prof * n.z};
        });
        done();
    }
    /* 1 floor mosaic: a fan that brings the total to exactly 262,267 triangles */
    {
        const size_t target = 262267;
        int rest = (int)(target - tris); /* 17,243 = 401 * 43: 401 segments x (21 rings of quads + centre fan) */
        const int ns = 401, nr = (rest / ns - 1) / 2;
        const Vec3 c{-60.0f, Y0 + 0.5f, -40.0f};
        const float R = 380.0f, r0 = R / (float)(nr + 1);
        b.fan(c, Vec3{0, 1, 0}, r0, ns, Vec3{1, 0, 0}, Vec3{0, 0, 1});
        b.grid(ns, nr, [&](float s, float t, Vec3& p, Vec3& n) {
            float a = TWO_PI * s, r = r0 + (R - r0) * t;
            p = Vec3{c.x + r * std::cos(a), c.y, c.z + r * std::sin(a)};
            n = Vec3{0, 1, 0};
        });
        done();
    }
    (void)objs;
}

} // namespace gpurt

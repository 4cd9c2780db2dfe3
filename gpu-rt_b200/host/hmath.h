/*
 * hmath.h — host-side fp32 math for the scene / camera front end.
 *
 * The reference computes every model matrix, modelIT, light bbox and camera matrix on the host with
 * its header-only math library; the GPU only ever sees the resulting floats.  To be a drop-in those
 * floats must come out bit-identical, so each routine here restates the reference's evaluation ORDER
 * (cited per function, paths relative to the reference root) in this repo's own types.  Compile
 * without -ffast-math and with -ffp-contract=off.
 */
#pragma once
#include <cfloat>
#include <cmath>
#include <cstring>

namespace gpurt {

constexpr float kPi = 3.14159265358979323846264338327950288f; /* lib/mathlib.h:17 */
inline float radians(float v) { return v * (kPi / 180.0f); }  /* lib/mathlib.h:18 */
inline float degrees(float v) { return v * (180.0f / kPi); }  /* lib/mathlib.h:19 */

struct Vec2 {
    float x = 0, y = 0;
};

struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    explicit Vec3(float s) : x(s), y(s), z(s) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    Vec3 operator+(Vec3 o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vec3 operator-(Vec3 o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vec3 operator-() const { return {-x, -y, -z}; }
    Vec3 operator*(float s) const { return {x * s, y * s, z * s}; }
    Vec3 operator/(float s) const { return {x / s, y / s, z / s}; }
    bool operator==(Vec3 o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(Vec3 o) const { return x != o.x || y != o.y || z != o.z; }
    float norm_squared() const { return x * x + y * y + z * z; } /* lib/vec3.h:155 */
    float norm() const { return std::sqrt(norm_squared()); }
    Vec3 unit() const {
        float n = norm();
        return {x / n, y / n, z / n};
    }
};
inline Vec3 operator*(float s, Vec3 v) { return {v.x * s, v.y * s, v.z * s}; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 l, Vec3 r) { /* lib/vec3.h cross */
    return {l.y * r.z - l.z * r.y, l.z * r.x - l.x * r.z, l.x * r.y - l.y * r.x};
}

struct Vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    Vec4() = default;
    Vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    Vec4(Vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    Vec3 xyz() const { return {x, y, z}; }
};

/* Column-major 4x4 (lib/mat4.h:272-275): c[col][row], data()[4*col+row]. */
struct Mat4 {
    float c[4][4];

    Mat4() { *this = identity(); } /* lib/mat4.h:50-52: default is I */
    static Mat4 zero() {
        Mat4 m(0);
        return m;
    }
    static Mat4 identity() {
        Mat4 m(0);
        m.c[0][0] = m.c[1][1] = m.c[2][2] = m.c[3][3] = 1.0f;
        return m;
    }
    const float* data() const { return &c[0][0]; }
    float* data() { return &c[0][0]; }

    /* lib/mat4.h:136-147: ret[i][j] = sum_k m[i][k] * this[k][j], accumulated from 0 in k order */
    Mat4 operator*(const Mat4& m) const {
        Mat4 r(0);
        for(int i = 0; i < 4; i++)
            for(int j = 0; j < 4; j++) {
                float acc = 0.0f;
                for(int k = 0; k < 4; k++) acc += m.c[i][k] * c[k][j];
                r.c[i][j] = acc;
            }
        return r;
    }
    /* lib/mat4.h:149-151: v0*col0 + v1*col1 + v2*col2 + v3*col3, left to right */
    Vec4 operator*(Vec4 v) const {
        Vec4 r;
        for(int j = 0; j < 4; j++)
            r[j] = ((v[0] * c[0][j] + v[1] * c[1][j]) + v[2] * c[2][j]) + v[3] * c[3][j];
        return r;
    }
    Mat4 T() const { /* lib/mat4.h:323-331 */
        Mat4 r(0);
        for(int i = 0; i < 4; i++)
            for(int j = 0; j < 4; j++) r.c[i][j] = c[j][i];
        return r;
    }

    /* lib/mat4.h:208-233 — 24 signed 4-factor products summed left to right. Entry = sign then
     * factors written as 10*col+row. */
    float det() const {
        static const signed char T[24][5] = {
            {1, 3, 12, 21, 30},   {-1, 2, 13, 21, 30}, {-1, 3, 11, 22, 30}, {1, 1, 13, 22, 30},
            {1, 2, 11, 23, 30},   {-1, 1, 12, 23, 30}, {-1, 3, 12, 20, 31}, {1, 2, 13, 20, 31},
            {1, 3, 10, 22, 31},   {-1, 0, 13, 22, 31}, {-1, 2, 10, 23, 31}, {1, 0, 12, 23, 31},
            {1, 3, 11, 20, 32},   {-1, 1, 13, 20, 32}, {-1, 3, 10, 21, 32}, {1, 0, 13, 21, 32},
            {1, 1, 10, 23, 32},   {-1, 0, 11, 23, 32}, {-1, 2, 11, 20, 33}, {1, 1, 12, 20, 33},
            {1, 2, 10, 21, 33},   {-1, 0, 12, 21, 33}, {-1, 1, 10, 22, 33}, {1, 0, 11, 22, 33}};
        float acc = 0.0f;
        for(int t = 0; t < 24; t++) {
            float p = ((at(T[t][1]) * at(T[t][2])) * at(T[t][3])) * at(T[t][4]);
            acc = t == 0 ? p : (T[t][0] > 0 ? acc + p : acc - p);
        }
        return acc;
    }
    /* lib/mat4.h:333-385 — cofactor expansion, 6 signed 3-factor products per entry in the
     * reference's order, then every entry divided by det(). */
    Mat4 inverse() const {
        static const signed char T[16][6][4] = {
            {{1, 12, 23, 31}, {-1, 13, 22, 31}, {1, 13, 21, 32}, {-1, 11, 23, 32}, {-1, 12, 21, 33}, {1, 11, 22, 33}},
            {{1, 3, 22, 31}, {-1, 2, 23, 31}, {-1, 3, 21, 32}, {1, 1, 23, 32}, {1, 2, 21, 33}, {-1, 1, 22, 33}},
            {{1, 2, 13, 31}, {-1, 3, 12, 31}, {1, 3, 11, 32}, {-1, 1, 13, 32}, {-1, 2, 11, 33}, {1, 1, 12, 33}},
            {{1, 3, 12, 21}, {-1, 2, 13, 21}, {-1, 3, 11, 22}, {1, 1, 13, 22}, {1, 2, 11, 23}, {-1, 1, 12, 23}},
            {{1, 13, 22, 30}, {-1, 12, 23, 30}, {-1, 13, 20, 32}, {1, 10, 23, 32}, {1, 12, 20, 33}, {-1, 10, 22, 33}},
            {{1, 2, 23, 30}, {-1, 3, 22, 30}, {1, 3, 20, 32}, {-1, 0, 23, 32}, {-1, 2, 20, 33}, {1, 0, 22, 33}},
            {{1, 3, 12, 30}, {-1, 2, 13, 30}, {-1, 3, 10, 32}, {1, 0, 13, 32}, {1, 2, 10, 33}, {-1, 0, 12, 33}},
            {{1, 2, 13, 20}, {-1, 3, 12, 20}, {1, 3, 10, 22}, {-1, 0, 13, 22}, {-1, 2, 10, 23}, {1, 0, 12, 23}},
            {{1, 11, 23, 30}, {-1, 13, 21, 30}, {1, 13, 20, 31}, {-1, 10, 23, 31}, {-1, 11, 20, 33}, {1, 10, 21, 33}},
            {{1, 3, 21, 30}, {-1, 1, 23, 30}, {-1, 3, 20, 31}, {1, 0, 23, 31}, {1, 1, 20, 33}, {-1, 0, 21, 33}},
            {{1, 1, 13, 30}, {-1, 3, 11, 30}, {1, 3, 10, 31}, {-1, 0, 13, 31}, {-1, 1, 10, 33}, {1, 0, 11, 33}},
            {{1, 3, 11, 20}, {-1, 1, 13, 20}, {-1, 3, 10, 21}, {1, 0, 13, 21}, {1, 1, 10, 23}, {-1, 0, 11, 23}},
            {{1, 12, 21, 30}, {-1, 11, 22, 30}, {-1, 12, 20, 31}, {1, 10, 22, 31}, {1, 11, 20, 32}, {-1, 10, 21, 32}},
            {{1, 1, 22, 30}, {-1, 2, 21, 30}, {1, 2, 20, 31}, {-1, 0, 22, 31}, {-1, 1, 20, 32}, {1, 0, 21, 32}},
            {{1, 2, 11, 30}, {-1, 1, 12, 30}, {-1, 2, 10, 31}, {1, 0, 12, 31}, {1, 1, 10, 32}, {-1, 0, 11, 32}},
            {{1, 1, 12, 20}, {-1, 2, 11, 20}, {1, 2, 10, 21}, {-1, 0, 12, 21}, {-1, 1, 10, 22}, {1, 0, 11, 22}}};
        Mat4 r(0);
        for(int e = 0; e < 16; e++) {
            float acc = 0.0f;
            for(int t = 0; t < 6; t++) {
                float p = (at(T[e][t][1]) * at(T[e][t][2])) * at(T[e][t][3]);
                acc = t == 0 ? p : (T[e][t][0] > 0 ? acc + p : acc - p);
            }
            r.c[e / 4][e % 4] = acc;
        }
        float d = det();
        for(int i = 0; i < 4; i++)
            for(int j = 0; j < 4; j++) r.c[i][j] = r.c[i][j] / d;
        return r;
    }

    static Mat4 translate(Vec3 t) { /* lib/mat4.h:416-420 */
        Mat4 r;
        r.c[3][0] = t.x, r.c[3][1] = t.y, r.c[3][2] = t.z, r.c[3][3] = 1.0f;
        return r;
    }
    static Mat4 scale(Vec3 s) { /* lib/mat4.h:446-452 */
        Mat4 r;
        r.c[0][0] = s.x, r.c[1][1] = s.y, r.c[2][2] = s.z;
        return r;
    }
    /* lib/mat4.h:428-444, angle in degrees */
    static Mat4 rotate(float t, Vec3 axis) {
        Mat4 r;
        float co = std::cos(radians(t)), si = std::sin(radians(t));
        axis = axis.unit();
        Vec3 tmp = axis * (1.0f - co);
        r.c[0][0] = co + tmp[0] * axis[0];
        r.c[0][1] = tmp[0] * axis[1] + si * axis[2];
        r.c[0][2] = tmp[0] * axis[2] - si * axis[1];
        r.c[1][0] = tmp[1] * axis[0] - si * axis[2];
        r.c[1][1] = co + tmp[1] * axis[1];
        r.c[1][2] = tmp[1] * axis[2] + si * axis[0];
        r.c[2][0] = tmp[2] * axis[0] + si * axis[1];
        r.c[2][1] = tmp[2] * axis[1] - si * axis[0];
        r.c[2][2] = co + tmp[2] * axis[2];
        return r;
    }
    static Mat4 euler(Vec3 a) { /* lib/mat4.h:422-426: Rz * Ry * Rx */
        return rotate(a.z, Vec3{0, 0, 1}) * rotate(a.y, Vec3{0, 1, 0}) * rotate(a.x, Vec3{1, 0, 0});
    }
    /* lib/mat4.h:465-475: reverse-Z, y-flipped, infinite far plane */
    static Mat4 project(float fov, float ar, float n) {
        float f = 1.0f / std::tan(radians(fov) / 2.0f);
        Mat4 r;
        r.c[0][0] = f / ar;
        r.c[1][1] = -f;
        r.c[2][2] = 0.0f;
        r.c[3][3] = 0.0f;
        r.c[3][2] = n;
        r.c[2][3] = -1.0f;
        return r;
    }
    /* lib/mat4.h:387-401 / :403-410 */
    static Mat4 rotate_to(Vec3 dir) {
        dir = dir.unit();
        if(std::abs(dir.y - 1.0f) < 0.00001f) return Mat4();
        if(std::abs(dir.y + 1.0f) < 0.00001f) {
            Mat4 m;
            m.c[1][1] = -1.0f;
            return m;
        }
        Vec3 x = cross(dir, Vec3{0, 1, 0}).unit();
        Vec3 z = cross(x, dir).unit();
        Mat4 m;
        m.set_col(0, Vec4{x, 0}), m.set_col(1, Vec4{dir, 0}), m.set_col(2, Vec4{z, 0});
        return m;
    }
    static Mat4 rotate_z_to(Vec3 dir) {
        Mat4 y = rotate_to(dir);
        Vec4 cy = y.col(1), cz = y.col(2);
        y.set_col(1, cz);
        y.set_col(2, Vec4{-cy.x, -cy.y, -cy.z, -cy.w});
        return y;
    }
    /* lib/mat4.h:158-196 */
    Vec3 to_euler() const {
        static const float sing[12] = {1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0};
        bool single = true;
        for(int i = 0; i < 12 && single; i++)
            single = single && std::abs(data()[i] - sing[i]) < 16.0f * FLT_EPSILON;
        if(single) return Vec3{0.0f, 0.0f, 180.0f};
        Vec3 e1, e2;
        float cy = ::hypotf(c[0][0], c[0][1]);
        if(cy > 16.0f * FLT_EPSILON) {
            e1[0] = std::atan2(c[1][2], c[2][2]);
            e1[1] = std::atan2(-c[0][2], cy);
            e1[2] = std::atan2(c[0][1], c[0][0]);
            e2[0] = std::atan2(-c[1][2], -c[2][2]);
            e2[1] = std::atan2(-c[0][2], -cy);
            e2[2] = std::atan2(-c[0][1], -c[0][0]);
        } else {
            e1[0] = std::atan2(-c[2][1], c[1][1]);
            e1[1] = std::atan2(-c[0][2], cy);
            e1[2] = 0;
            e2 = e1;
        }
        float d1 = std::abs(e1[0]) + std::abs(e1[1]) + std::abs(e1[2]);
        float d2 = std::abs(e2[0]) + std::abs(e2[1]) + std::abs(e2[2]);
        Vec3 e = d1 > d2 ? e2 : e1;
        return Vec3{degrees(e.x), degrees(e.y), degrees(e.z)};
    }
    /* lib/mat4.h:235-270. NB the loader feeds this the TRANSPOSED node matrix
     * (scene/scene.cpp:355), so translation is read from row 3. */
    void decompose(Vec3& pos, Vec3& scl, Vec3& rot) const {
        pos = Vec3{c[0][3], c[1][3], c[2][3]};
        Vec3 v[3] = {Vec3{c[0][0], c[1][0], c[2][0]}, Vec3{c[0][1], c[1][1], c[2][1]},
                     Vec3{c[0][2], c[1][2], c[2][2]}};
        scl = Vec3{v[0].norm(), v[1].norm(), v[2].norm()};
        if(det() < 0) scl = -scl;
        if(scl.x) v[0] = v[0] / scl.x;
        if(scl.y) v[1] = v[1] / scl.y;
        if(scl.z) v[2] = v[2] / scl.z;
        const float eps = 0.00001f;
        rot.y = std::asin(-v[0].z);
        float C = std::cos(rot.y);
        if(std::fabs(C) > eps) {
            float tx = v[2].z / C, ty = v[1].z / C;
            rot.x = std::atan2(ty, tx);
            tx = v[0].x / C, ty = v[0].y / C;
            rot.z = std::atan2(ty, tx);
        } else {
            rot.x = 0;
            float tx = v[1].y, ty = -v[1].x;
            rot.z = std::atan2(ty, tx);
        }
        rot = Vec3{degrees(rot.x), degrees(rot.y), degrees(rot.z)};
    }

    Vec4 col(int i) const { return {c[i][0], c[i][1], c[i][2], c[i][3]}; }
    void set_col(int i, Vec4 v) { c[i][0] = v.x, c[i][1] = v.y, c[i][2] = v.z, c[i][3] = v.w; }

private:
    explicit Mat4(int) { std::memset(c, 0, sizeof(c)); }
    float at(int cr) const { return c[cr / 10][cr % 10]; }
};

struct Quat {
    float x = 0, y = 0, z = 0, w = 1;
    Quat() = default;
    Quat(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    /* lib/quat.h:55-70, XYZ euler in degrees */
    static Quat euler(Vec3 a) {
        if(a == Vec3{0.0f, 0.0f, 180.0f} || a == Vec3{180.0f, 0.0f, 0.0f})
            return Quat{0.0f, 0.0f, -1.0f, 0.0f};
        float c1 = std::cos(radians(a[2] * 0.5f)), c2 = std::cos(radians(a[1] * 0.5f)),
              c3 = std::cos(radians(a[0] * 0.5f));
        float s1 = std::sin(radians(a[2] * 0.5f)), s2 = std::sin(radians(a[1] * 0.5f)),
              s3 = std::sin(radians(a[0] * 0.5f));
        return Quat{c1 * c2 * s3 - s1 * s2 * c3, c1 * s2 * c3 + s1 * c2 * s3,
                    s1 * c2 * c3 - c1 * s2 * s3, c1 * c2 * c3 + s1 * s2 * s3};
    }
    Quat conjugate() const { return {-x, -y, -z, w}; }
    /* lib/quat.h:106-109 */
    Quat operator*(const Quat& r) const {
        return {y * r.z - z * r.y + x * r.w + w * r.x, z * r.x - x * r.z + y * r.w + w * r.y,
                x * r.y - y * r.x + z * r.w + w * r.z, w * r.w - x * r.x - y * r.y - z * r.z};
    }
    /* lib/quat.h:143-145 */
    Vec3 rotate(Vec3 v) const {
        Quat q = ((*this) * Quat{v.x, v.y, v.z, 0}) * conjugate();
        return {q.x, q.y, q.z};
    }
    /* lib/quat.h:134-140 */
    Mat4 to_mat() const {
        Mat4 m;
        m.set_col(0, {1 - 2 * y * y - 2 * z * z, 2 * x * y + 2 * z * w, 2 * x * z - 2 * y * w, 0.0f});
        m.set_col(1, {2 * x * y - 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z + 2 * x * w, 0.0f});
        m.set_col(2, {2 * x * z + 2 * y * w, 2 * y * z - 2 * x * w, 1 - 2 * x * x - 2 * y * y, 0.0f});
        m.set_col(3, {0.0f, 0.0f, 0.0f, 1.0f});
        return m;
    }
};

struct BBox {
    Vec3 min{FLT_MAX}, max{-FLT_MAX}; /* lib/bbox.h:17 */
    void enclose(Vec3 p) {            /* lib/bbox.h:34-37 (hmin/hmax = std::min/max per axis) */
        min = {std::fmin(min.x, p.x), std::fmin(min.y, p.y), std::fmin(min.z, p.z)};
        max = {std::fmax(max.x, p.x), std::fmax(max.y, p.y), std::fmax(max.z, p.z)};
    }
    /* lib/bbox.h:62-78 */
    void transform(const Mat4& t) {
        Vec3 amin = min, amax = max;
        min = max = Vec3{t.c[3][0], t.c[3][1], t.c[3][2]};
        for(int i = 0; i < 3; i++)
            for(int j = 0; j < 3; j++) {
                float a = t.c[j][i] * amin[j], b = t.c[j][i] * amax[j];
                if(a < b) min[i] += a, max[i] += b;
                else min[i] += b, max[i] += a;
            }
    }
};

} // namespace gpurt

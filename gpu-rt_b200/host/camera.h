/*
 * camera.h — orbit camera of the reference (src/util/camera.{h,cpp}); it defines every primary ray
 * through V, P and their inverses (src/vk/rt.cpp:121-127, src/shaders/rt/rt.rgen:551-565).
 */
#pragma once
#include "hmath.h"

namespace gpurt {

class Camera {
public:
    explicit Camera(Vec2 dim) { /* camera.cpp:5-8 */
        reset();
        aspect_ratio = dim.x / dim.y;
    }
    /* camera.cpp:58-71 */
    void reset() {
        vert_fov = 90.0f;
        aspect_ratio = 1.7778f;
        rot = Quat{0.059338f, 0.39328f, 0.025433f, -0.91759f};
        near_plane = 0.01f;
        radius = 2.0f;
        looking_at = Vec3{-2.21737f, -2.33f, 2.7778f};
        update_pos();
    }
    /* camera.cpp:47-56 */
    void look_at(Vec3 cent, Vec3 pos) {
        position = pos;
        looking_at = cent;
        radius = (pos - cent).norm();
        Vec3 front = (looking_at - position).unit();
        if(dot(front, Vec3{0.0f, 1.0f, 0.0f}) == -1.0f) rot = Quat::euler(Vec3{270.0f, 0.0f, 0.0f});
        else rot = Quat::euler(Mat4::rotate_z_to(front).to_euler());
        update_pos();
    }
    void set_fov(float f) { vert_fov = f; }
    void set_ar(float a) { aspect_ratio = a; }
    Mat4 get_view() const { return view; }
    Mat4 get_proj() const { return Mat4::project(vert_fov, aspect_ratio, near_plane); } /* :31-33 */
    Vec3 pos() const { return position; }

private:
    /* camera.cpp:150-155 */
    void update_pos() {
        position = rot.rotate(Vec3{0.0f, 0.0f, 1.0f});
        position = looking_at + radius * position.unit();
        iview = Mat4::translate(position) * rot.to_mat();
        view = iview.inverse();
    }
    Vec3 position, looking_at;
    float vert_fov, aspect_ratio, near_plane, radius;
    Quat rot;
    Mat4 view, iview;
};

} // namespace gpurt

/*
 * ref_api.cpp — C entry points over the UNMODIFIED reference host code (TEST INFRASTRUCTURE).
 * Used only by tests/golden/make_golden.py to pin this repo's host front end (gltf loader, math,
 * camera, Scene_Desc/Scene_Light packing) against the reference's own classes.
 *
 * Packing follows RTPipe::build_desc (src/vk/rt.cpp:26-76) and GPURT::build_accel
 * (src/gpurt.cpp:220-241) using the reference's Mat4/BBox/Pose code; those two functions themselves
 * need a VkDevice so their few lines of arithmetic are restated here against reference types.
 */
#include <scene/scene.h>
#include <util/camera.h>
#include <cstring>
#include <vector>

struct RefObj {
    const VK::Mesh* mesh;
    Mat4 model, modelIT;
    Material mat;
    unsigned int id;
};
struct RefLight {
    Vec4 bmin, bmax;
    unsigned int index, n_triangles;
};
struct RefScene {
    Scene scene;
    std::vector<RefObj> objs;
    std::vector<RefLight> lights;
};

extern "C" {

void* ref_scene_load(const char* path, float scale) {
    RefScene* R = new RefScene;
    Camera cam(Vec2{1280.0f, 720.0f});
    R->scene.load(path, cam);
    R->scene.scale = scale;
    /* rt.cpp:31-45 / gpurt.cpp:228-231 */
    R->scene.for_objs([&](const Object& obj) {
        RefObj o;
        o.mesh = &obj.mesh();
        o.model = Mat4::scale(Vec3{R->scene.scale}) * obj.pose.transform();
        o.modelIT = o.model.inverse().T();
        o.mat = obj.material;
        o.id = obj.id();
        R->objs.push_back(o);
    });
    /* rt.cpp:47-60 */
    unsigned int i = 0;
    R->scene.for_objs([&](const Object& obj) {
        if(obj.material.emissive != Vec3{} || obj.material.emissive_tex != -1) {
            RefLight l;
            l.index = i;
            l.n_triangles = obj.mesh().inds().size() / 3;
            BBox box = obj.mesh().bbox();
            box.transform(R->objs[i].model);
            l.bmin = Vec4{box.min, 0.0f};
            l.bmax = Vec4{box.max, 0.0f};
            R->lights.push_back(l);
        }
        i++;
    });
    return R;
}
void ref_scene_free(void* h) { delete(RefScene*)h; }
int ref_scene_n_objs(void* h) { return (int)((RefScene*)h)->objs.size(); }
int ref_scene_n_lights(void* h) { return (int)((RefScene*)h)->lights.size(); }
int ref_scene_n_textures(void* h) { return (int)((RefScene*)h)->scene.n_textures(); }
/* decoded texture i (tinygltf -> stb_image, RGBA8): sizes, then texels when rgba != NULL */
void ref_texture_get(void* h, int i, unsigned* w, unsigned* ht, unsigned char* rgba) {
    const Util::Image& im = ((RefScene*)h)->scene.images()[i];
    *w = im.w(), *ht = im.h();
    if(rgba) std::memcpy(rgba, im.data(), im.bytes());
}
void ref_obj_counts(void* h, int i, unsigned* nv, unsigned* ni, unsigned* id) {
    RefObj& o = ((RefScene*)h)->objs[i];
    *nv = o.mesh->verts().size();
    *ni = o.mesh->inds().size();
    *id = o.id;
}
/* verts: 12 floats each (mesh.h:16-22), mats: model[16], modelIT[16] column-major,
 * matl: albedo3 emissive3 metal_rough2 ; tex: albedo, emissive, metal_rough, normal */
void ref_obj_get(void* h, int i, float* verts, unsigned* idx, float* mats32, float* matl8, int* tex4) {
    RefObj& o = ((RefScene*)h)->objs[i];
    std::memcpy(verts, o.mesh->verts().data(), o.mesh->verts().size() * sizeof(VK::Mesh::Vertex));
    std::memcpy(idx, o.mesh->inds().data(), o.mesh->inds().size() * 4);
    std::memcpy(mats32, o.model.data, 64);
    std::memcpy(mats32 + 16, o.modelIT.data, 64);
    matl8[0] = o.mat.albedo.x, matl8[1] = o.mat.albedo.y, matl8[2] = o.mat.albedo.z;
    matl8[3] = o.mat.emissive.x, matl8[4] = o.mat.emissive.y, matl8[5] = o.mat.emissive.z;
    matl8[6] = o.mat.metal_rough.x, matl8[7] = o.mat.metal_rough.y;
    tex4[0] = o.mat.albedo_tex, tex4[1] = o.mat.emissive_tex, tex4[2] = o.mat.metal_rough_tex,
    tex4[3] = o.mat.normal_tex;
}
void ref_light_get(void* h, int i, float* bmin4, float* bmax4, unsigned* index, unsigned* ntris) {
    RefLight& l = ((RefScene*)h)->lights[i];
    std::memcpy(bmin4, l.bmin.data, 16);
    std::memcpy(bmax4, l.bmax.data, 16);
    *index = l.index;
    *ntris = l.n_triangles;
}

/* Camera: src/util/camera.cpp. mode 0 = Camera::reset() defaults (camera.cpp:58-71) with aspect
 * w/h; mode 1 = look_at(center,pos) + set_fov. out: V,P,iV,iP (rt.cpp:121-127), 64 floats. */
void ref_camera(int mode, float w, float h, const float* pos3, const float* center3, float vfov,
                float* out64) {
    Camera cam(Vec2{w, h});
    if(mode == 1) {
        cam.look_at(Vec3{center3[0], center3[1], center3[2]}, Vec3{pos3[0], pos3[1], pos3[2]});
        cam.set_fov(vfov);
    }
    Mat4 V = cam.get_view(), P = cam.get_proj();
    Mat4 iV = V.inverse(), iP = P.inverse();
    std::memcpy(out64, V.data, 64);
    std::memcpy(out64 + 16, P.data, 64);
    std::memcpy(out64 + 32, iV.data, 64);
    std::memcpy(out64 + 48, iP.data, 64);
}

void ref_mat4_mul(const float* a, const float* b, float* out) {
    Mat4 A, B;
    std::memcpy(A.data, a, 64);
    std::memcpy(B.data, b, 64);
    Mat4 C = A * B;
    std::memcpy(out, C.data, 64);
}
void ref_mat4_inverse(const float* a, float* out) {
    Mat4 A;
    std::memcpy(A.data, a, 64);
    Mat4 C = A.inverse();
    std::memcpy(out, C.data, 64);
}

} /* extern "C" */

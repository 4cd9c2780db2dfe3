/*
 * app_types.h — TEST INFRASTRUCTURE: stand-ins for the parts of the reference application that surround
 * VK::Accel / VK::RTPipe (Scene, Object, VK::Mesh, Camera, Mat4, Util::Image, VK::Image / ImageView), shaped like the
 * reference's declarations (src/scene/scene.h:12-48, object.h, src/vk/mesh.h:15-45, src/util/camera.h, util/image.h)
 * so that the call sites of src/gpurt.cpp compile against gpu-rt_b200/host/vk_dropin.h without the reference tree.
 * Built on this repo's own host front end (gpu-rt_b200/host/scene.h: same loader semantics and object order).
 * tests/test_dropin.py also builds the same program against the reference's REAL headers when /root/reference exists.
 */
#pragma once
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../gpu-rt_b200/host/camera.h"
#include "../../gpu-rt_b200/host/scene.h"

using gpurt::Mat4;
using gpurt::Vec2;
using gpurt::Vec3;
using gpurt::Vec4;
using gpurt::BBox;
using gpurt::Camera;
using gpurt::Material;
using gpurt::Pose;

namespace VK {
struct Mesh { /* src/vk/mesh.h:15-45 */
    typedef unsigned int Index;
    typedef gpurt::Vertex Vertex;
    Mesh() = default;
    Mesh(std::vector<Vertex>&& v, std::vector<Index>&& i) : _verts(std::move(v)), _idxs(std::move(i)) {}
    const std::vector<Vertex>& verts() const { return _verts; }
    const std::vector<Index>& inds() const { return _idxs; }

private:
    std::vector<Vertex> _verts;
    std::vector<Index> _idxs;
};
struct Image { /* rt_target: only its extent is read at the call site (gpurt.cpp:45) */
    unsigned int w = 0, h = 0;
};
struct ImageView {};
} // namespace VK

namespace Util {
struct Image { /* src/util/image.h:36-42 */
    std::pair<unsigned int, unsigned int> dim() const { return {w, h}; }
    const unsigned char* data() const { return px.data(); }
    unsigned int w = 0, h = 0;
    std::vector<unsigned char> px;
};
} // namespace Util

class Object { /* src/scene/object.h */
public:
    Object(unsigned int id, Pose p, VK::Mesh&& m, Material mat) : pose(p), material(mat), _id(id), _mesh(std::move(m)) {}
    unsigned int id() const { return _id; }
    const VK::Mesh& mesh() const { return _mesh; }
    Pose pose;
    Material material;

private:
    unsigned int _id = 0;
    VK::Mesh _mesh;
};

class Scene { /* src/scene/scene.h:12-48 */
public:
    std::string load(std::string file, Camera&) {
        gpurt::Scene s;
        std::string err;
        if(!s.load(file, err)) return err;
        objs.clear(), textures.clear();
        s.for_objs([&](const gpurt::Object& o) { /* the reference container's iteration order */
            std::vector<VK::Mesh::Vertex> v = o.mesh.verts;
            std::vector<VK::Mesh::Index> i(o.mesh.idx.begin(), o.mesh.idx.end());
            objs.emplace_back(o.id, o.pose, VK::Mesh(std::move(v), std::move(i)), o.material);
        });
        for(const gpurt::Texture& t : s.textures) {
            Util::Image im;
            im.w = t.w, im.h = t.h, im.px = t.rgba;
            textures.push_back(std::move(im));
        }
        return {};
    }
    template <typename F> void for_objs(F&& func) {
        for(auto& o : objs) func(o);
    }
    template <typename F> void for_objs(F&& func) const {
        for(auto& o : objs) func(o);
    }
    Object& get(unsigned int id) {
        for(auto& o : objs)
            if(o.id() == id) return o;
        throw std::runtime_error("no such object");
    }
    size_t size() const { return objs.size(); }
    const std::vector<Util::Image>& images() const { return textures; }
    float scale = 1.0f;

private:
    std::vector<Object> objs;
    std::vector<Util::Image> textures;
};

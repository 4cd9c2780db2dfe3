/*
 * vk_stubs.cpp — link-time stand-ins for the Vulkan resource classes the reference's scene loader
 * touches (TEST INFRASTRUCTURE). The reference's scene.cpp / object.cpp / pose.cpp / camera.cpp and
 * all of src/lib are compiled UNMODIFIED from /root/reference by oracle/Makefile; only the GPU
 * resource objects (which need a live VkDevice) are replaced by no-ops here.
 * Declarations come from the reference's own headers (src/vk/vulkan.h, src/vk/mesh.h).
 */
#include <vk/mesh.h>

namespace VK {

Manager& vk() {
    static Manager singleton;
    return singleton;
}

#define STUB_RESOURCE(T)                                                                           \
    T::~T() {}                                                                                     \
    T::T(T&&) {}                                                                                   \
    T& T::operator=(T&&) { return *this; }

STUB_RESOURCE(Buffer)
STUB_RESOURCE(Image)
STUB_RESOURCE(ImageView)
STUB_RESOURCE(Sampler)
STUB_RESOURCE(Shader)
STUB_RESOURCE(Framebuffer)
STUB_RESOURCE(Accel)
STUB_RESOURCE(PipeData)
Pass::~Pass() {}
Pass::Pass(Pass&&) {}
Pass& Pass::operator=(Pass&&) { return *this; }

void Buffer::recreate(VkDeviceSize, VkBufferUsageFlags, VmaMemoryUsage) {}

/* src/vk/mesh.cpp:37-76 minus the two vbuf/ibuf allocations (no device here): the host vectors
 * and the object-space bbox (mesh.cpp:73-75) are what the hot path consumes. */
Mesh::Mesh(std::vector<Mesh::Vertex>&& vertices, std::vector<Mesh::Index>&& indices) {
    recreate(std::move(vertices), std::move(indices));
}
Mesh::Mesh(Mesh&& src) { *this = std::move(src); }
Mesh& Mesh::operator=(Mesh&& src) {
    _verts = std::move(src._verts);
    _idxs = std::move(src._idxs);
    _bbox = std::move(src._bbox);
    dirty = src.dirty;
    src.dirty = true;
    return *this;
}
void Mesh::recreate(std::vector<Mesh::Vertex>&& vertices, std::vector<Mesh::Index>&& indices) {
    _verts = std::move(vertices);
    _idxs = std::move(indices);
    dirty = true;
    BBox box;
    for(auto& v : _verts) box.enclose(v.pos.xyz());
    _bbox = box;
}

} // namespace VK

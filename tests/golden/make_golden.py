#!/usr/bin/env python
"""Generate tests/golden/ref_host_golden.json from the REFERENCE's own host code.

Runs only in the build container (needs /root/reference and oracle/_ref/libgpurt_ref.so, which is the
reference's scene.cpp / object.cpp / pose.cpp / camera.cpp / lib/*.h compiled unmodified behind VK stubs,
see oracle/Makefile).  The fixture pins this repo's host front end (glTF loader, Mat4 math, camera,
Scene_Desc / Scene_Light packing) bit for bit: floats are stored as uint32 words, geometry as SHA-256.

    make -C oracle ref && python tests/golden/make_golden.py
"""
import base64
import ctypes as C
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_MEDIA = "/root/reference/media"

ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgpurt_ref.so"))
ref.ref_scene_load.restype = C.c_void_p
ref.ref_scene_load.argtypes = [C.c_char_p, C.c_float]


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def words(a):
    return [int(x) for x in np.ascontiguousarray(a).view(np.uint32).reshape(-1)]


def dump_scene(rel, scale, base=REF_MEDIA):
    h = C.c_void_p(ref.ref_scene_load(os.path.join(base, rel).encode(), scale))
    n = ref.ref_scene_n_objs(h)
    out = {"file": rel, "scale": scale, "n_textures": ref.ref_scene_n_textures(h), "objects": [], "lights": []}
    for i in range(n):
        nv, ni, idd = C.c_uint(), C.c_uint(), C.c_uint()
        ref.ref_obj_counts(h, i, C.byref(nv), C.byref(ni), C.byref(idd))
        v = np.empty((nv.value, 12), np.float32)
        ix = np.empty(ni.value, np.uint32)
        m = np.empty(32, np.float32)
        ml = np.empty(8, np.float32)
        tx = np.empty(4, np.int32)
        ref.ref_obj_get(h, i, vp(v), vp(ix), vp(m), vp(ml), vp(tx))
        out["objects"].append({
            "id": idd.value, "n_verts": nv.value, "n_indices": ni.value,
            "verts_sha256": hashlib.sha256(v.tobytes()).hexdigest(),
            "indices_sha256": hashlib.sha256(ix.tobytes()).hexdigest(),
            "model": words(m[:16]), "modelIT": words(m[16:]), "material": words(ml), "textures": [int(t) for t in tx]})
    for i in range(ref.ref_scene_n_lights(h)):
        a, b = np.empty(4, np.float32), np.empty(4, np.float32)
        ix, nt = C.c_uint(), C.c_uint()
        ref.ref_light_get(h, i, vp(a), vp(b), C.byref(ix), C.byref(nt))
        out["lights"].append({"bmin": words(a), "bmax": words(b), "index": ix.value, "n_triangles": nt.value})
    ref.ref_scene_free(h)
    return out


def dump_textures(gltf_path, names):
    """SHA-256 of every texture of `gltf_path` as decoded by the reference (tinygltf -> stb_image, RGBA8)"""
    h = C.c_void_p(ref.ref_scene_load(gltf_path.encode(), 1.0))
    out = []
    for i in range(ref.ref_scene_n_textures(h)):
        w, ht = C.c_uint(), C.c_uint()
        ref.ref_texture_get(h, i, C.byref(w), C.byref(ht), None)
        px = np.zeros((ht.value, w.value, 4), np.uint8)
        ref.ref_texture_get(h, i, C.byref(w), C.byref(ht), vp(px))
        out.append({"file": names[i], "w": w.value, "h": ht.value, "sha256": hashlib.sha256(px.tobytes()).hexdigest()})
    ref.ref_scene_free(h)
    return out


def texture_gltf(directory, files, out_path):
    """one-triangle glTF (the reference's parse_mesh needs a material) whose textures are `files`,
    symlinked next to it (tinygltf resolves image URIs relative to the glTF)"""
    for f in files:
        os.symlink(os.path.join(directory, f), os.path.join(os.path.dirname(out_path), f))
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    blob = pos.tobytes() + np.array([0, 1, 2], np.uint16).tobytes() + b"\0\0"
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
         "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1, "material": 0}]}],
         "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
         "textures": [{"source": i} for i in range(len(files))],
         "images": [{"uri": f} for f in files],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 6}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3", "min": [0, 0, 0], "max": [1, 1, 0]},
                       {"bufferView": 1, "componentType": 5123, "count": 3, "type": "SCALAR"}]}
    with open(out_path, "w") as f:
        json.dump(g, f)


def dump_camera(mode, w, h, pos, center, fov):
    o = np.zeros(64, np.float32)
    p = np.array(pos or (0, 0, 0), np.float32)
    c = np.array(center or (0, 0, 0), np.float32)
    ref.ref_camera(mode, C.c_float(w), C.c_float(h), vp(p), vp(c), C.c_float(fov), vp(o))
    return {"mode": mode, "w": w, "h": h, "pos": pos, "center": center, "vfov": fov, "V_P_iV_iP": words(o)}


def main():
    g = {"about": "generated by tests/golden/make_golden.py from the reference's own host sources (oracle/_ref)",
         "scenes": [dump_scene("cbox/cbox.gltf", 1.0), dump_scene("cbox/cbox.gltf", 0.37),
                    dump_scene("mis_test/mis_test.gltf", 1.0), dump_scene("cube.gltf", 1.0),
                    dump_scene("synth/features.gltf", 1.0, os.path.join(ROOT, "tests", "data")),
                    dump_scene("synth/features.gltf", 2.5, os.path.join(ROOT, "tests", "data")),
                    dump_scene("synth/embedded.glb", 1.0, os.path.join(ROOT, "tests", "data"))],
         "cameras": [dump_camera(0, 1280, 720, None, None, 0.0), dump_camera(0, 1024, 1024, None, None, 0.0),
                     dump_camera(1, 1920, 1080, [-1000.0, 200.0, 0.0], [0.0, 200.0, 0.0], 90.0),
                     dump_camera(1, 1920, 1080, [0.5, 0.6, 2.6], [0.5, 0.45, 0.0], 50.0),
                     dump_camera(1, 800, 600, [0.0, 10.0, 0.0], [0.0, 0.0, 0.0], 70.0),
                     dump_camera(1, 640, 480, [3.0, 2.0, 4.0], [0.0, 0.0, 0.0], 60.0)]}
    # textures: the committed synthetic set, and every texture of media/sponza (the .jpg / .png files are in
    # the reference snapshot even though Sponza.bin is not) through a scratch glTF with absolute image paths
    synth = os.path.join(ROOT, "tests", "data", "synth")
    names = [im["uri"] for im in json.load(open(os.path.join(synth, "textures.gltf")))["images"]]
    g["synth_textures"] = dump_textures(os.path.join(synth, "textures.gltf"), names)
    g["glb_textures"] = dump_textures(os.path.join(synth, "embedded.glb"), ["jprog_420.jpg (bufferView)", "p_pal.png (bufferView)"])
    sponza = os.path.join(REF_MEDIA, "sponza")
    files = sorted(f for f in os.listdir(sponza) if f.endswith((".jpg", ".png")))
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        texture_gltf(sponza, files, os.path.join(tmp, "sponza_textures.gltf"))
        g["sponza_textures"] = dump_textures(os.path.join(tmp, "sponza_textures.gltf"), files)
    with open(os.path.join(HERE, "ref_host_golden.json"), "w") as f:
        json.dump(g, f, indent=0, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "ref_host_golden.json"))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate tests/data/synth/: small JPEG / PNG textures covering the decoder paths (sampling 4:4:4, 4:2:2,
4:2:0, 4:1:1; odd sizes; grayscale; optimised Huffman tables; restart intervals; progressive; PNG colour
types) and a one-triangle glTF that references all of them.  Needs Pillow; the outputs are committed, and
tests/golden/make_golden.py records what the REFERENCE's loader (tinygltf -> stb_image) decodes them to.

    python tests/data/make_synth.py && python tests/golden/make_golden.py
"""
import base64
import json
import os

import numpy as np
from PIL import Image

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth")
rng = np.random.default_rng(11)


def img(w, h, mode="RGB"):
    y, x = np.mgrid[0:h, 0:w]
    a = np.stack([(x * 7 + y * 3) % 256, (x * x // 8 + y * 5) % 256, (x * y // 4) % 256], -1)
    a = (a // 2 + rng.integers(0, 128, (h, w, 3))).astype(np.uint8)
    im = Image.fromarray(a, "RGB")
    return im if mode == "RGB" else im.convert(mode)


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        os.remove(os.path.join(OUT, f))
    j = lambda name: os.path.join(OUT, name)
    img(37, 23).save(j("j444_37x23_q92.jpg"), quality=92, subsampling=0)
    img(37, 23).save(j("j422_37x23_q30.jpg"), quality=30, subsampling=1)
    img(37, 23).save(j("j420_37x23_q92.jpg"), quality=92, subsampling=2)
    img(1, 1).save(j("j420_1x1.jpg"), quality=90, subsampling=2)
    img(17, 9).save(j("j422_17x9.jpg"), quality=85, subsampling=1)
    img(40, 24).save(j("j411_40x24.jpg"), quality=80, subsampling="4:1:1")
    img(50, 40, "L").save(j("jgray_50x40.jpg"), quality=80)
    img(50, 40).save(j("jopt_420.jpg"), quality=85, optimize=True, subsampling=2)
    img(72, 45).save(j("jrst_420.jpg"), quality=85, subsampling=2, restart_marker_blocks=3)
    img(72, 45).save(j("jrst_rows_444.jpg"), quality=85, subsampling=0, restart_marker_rows=1)
    img(33, 31).save(j("jq100_444.jpg"), quality=100, subsampling=0)
    img(33, 31).save(j("jq1_420.jpg"), quality=1, subsampling=2)
    img(49, 35).save(j("jprog_420.jpg"), quality=85, progressive=True, subsampling=2)
    img(24, 40).save(j("jprog_444.jpg"), quality=80, progressive=True, subsampling=0)
    img(64, 64).save(j("jprog_422_q30.jpg"), quality=30, progressive=True, subsampling=1)
    img(40, 30, "L").save(j("jprog_gray.jpg"), quality=80, progressive=True)
    img(60, 41).save(j("jprog_rst.jpg"), quality=80, progressive=True, restart_marker_blocks=5)
    img(20, 12).save(j("p_rgb.png"))
    img(20, 12, "RGBA").save(j("p_rgba.png"))
    img(20, 12, "L").save(j("p_gray.png"))
    img(20, 12, "LA").save(j("p_la.png"))
    img(20, 12, "P").save(j("p_pal.png"))
    files = sorted(f for f in os.listdir(OUT) if f.endswith((".jpg", ".png")))
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    blob = pos.tobytes() + np.array([0, 1, 2], np.uint16).tobytes() + b"\0\0"
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
         "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1, "material": 0}]}],
         "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}],
         "textures": [{"source": i} for i in range(len(files))], "images": [{"uri": f} for f in files],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 6}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3", "min": [0, 0, 0], "max": [1, 1, 0]},
                       {"bufferView": 1, "componentType": 5123, "count": 3, "type": "SCALAR"}]}
    with open(j("textures.gltf"), "w") as f:
        json.dump(g, f, indent=1)
    print(len(files), "images ->", OUT)
    features()
    embedded_glb()


def embedded_glb():
    """embedded.glb: binary glTF whose geometry and two images (a progressive JPEG and a palette PNG) all live in
    the BIN chunk and are referenced through bufferViews (image.bufferView + mimeType)"""
    import struct
    pos = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 2, 1]], np.float32)
    uv = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32)
    idx = np.array([0, 1, 2, 2, 1, 3], np.uint16)
    jpg = open(os.path.join(OUT, "jprog_420.jpg"), "rb").read()
    png = open(os.path.join(OUT, "p_pal.png"), "rb").read()
    chunks, views = [], []
    for raw in (pos.tobytes(), uv.tobytes(), idx.tobytes(), jpg, png):
        off = sum(len(c) for c in chunks)
        views.append({"buffer": 0, "byteOffset": off, "byteLength": len(raw)})
        chunks.append(raw + b"\0" * ((-len(raw)) % 4))
    blob = b"".join(chunks)
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
         "nodes": [{"mesh": 0, "matrix": [1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0.5, 0.25, -1, 1]}],
         "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 0}]}],
         "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 1}},
                        "emissiveFactor": [0.5, 0.25, 0.125]}],
         "textures": [{"source": 0}, {"source": 1}],
         "images": [{"bufferView": 3, "mimeType": "image/jpeg"}, {"bufferView": 4, "mimeType": "image/png"}],
         "buffers": [{"byteLength": len(blob)}], "bufferViews": views,
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3", "min": [0, 0, 0], "max": [2, 2, 1]},
                       {"bufferView": 1, "componentType": 5126, "count": 4, "type": "VEC2"},
                       {"bufferView": 2, "componentType": 5123, "count": 6, "type": "SCALAR"}]}
    js = json.dumps(g).encode()
    js += b" " * ((-len(js)) % 4)
    glb = b"glTF" + struct.pack("<II", 2, 12 + 8 + len(js) + 8 + len(blob)) + struct.pack("<II", len(js), 0x4E4F534A) + js \
        + struct.pack("<II", len(blob), 0x004E4942) + blob
    with open(os.path.join(OUT, "embedded.glb"), "wb") as f:
        f.write(glb)
    print("embedded.glb:", len(glb), "bytes")


def features():
    """features.gltf: nested node matrices (the reference composes them as T * M.T(), scene.cpp:355), a TRS node
    (ignored), one mesh instanced twice, several primitives per mesh, triangle list / strip / fan,
    u8 / u16 / u32 indices, interleaved (byteStride) and tightly packed attributes, NORMAL / TANGENT / TEXCOORD_0
    present or missing, emissive materials (lights), texture references"""
    r = np.random.default_rng(21)
    chunks, views, accs = [], [], []

    def add(arr, ctype, atype, stride=None, target_off=0):
        raw = arr.tobytes()
        off = sum(len(c) for c in chunks)
        pad = (-len(raw)) % 4
        chunks.append(raw + b"\0" * pad)
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(raw)}
        if stride:
            v["byteStride"] = stride
        views.append(v)
        return len(views) - 1

    def acc(view, ctype, count, atype, off=0):
        accs.append({"bufferView": view, "byteOffset": off, "componentType": ctype, "count": count, "type": atype})
        return len(accs) - 1

    def grid(n):   # (n+1)^2 vertices, triangle list
        u = np.linspace(0, 1, n + 1, dtype=np.float32)
        x, y = np.meshgrid(u, u, indexing="ij")
        p = np.stack([x, y, 0.2 * np.sin(5 * x) * np.cos(3 * y)], -1).reshape(-1, 3).astype(np.float32)
        ix = []
        for i in range(n):
            for j in range(n):
                a = i * (n + 1) + j
                ix += [a, a + n + 1, a + 1, a + 1, a + n + 1, a + n + 2]
        return p, np.array(ix)

    # mesh 0 / primitive 0: tightly packed POSITION NORMAL TANGENT TEXCOORD_0, u16 triangle list
    p, ix = grid(6)
    nrm = r.standard_normal((len(p), 3)).astype(np.float32)
    tan = r.standard_normal((len(p), 4)).astype(np.float32)
    uv = r.random((len(p), 2), dtype=np.float32)
    a_p = acc(add(p, 0, 0), 5126, len(p), "VEC3")
    a_n = acc(add(nrm, 0, 0), 5126, len(p), "VEC3")
    a_t = acc(add(tan, 0, 0), 5126, len(p), "VEC4")
    a_uv = acc(add(uv, 0, 0), 5126, len(p), "VEC2")
    a_i = acc(add(ix.astype(np.uint16), 0, 0), 5123, len(ix), "SCALAR")
    prim0 = {"attributes": {"POSITION": a_p, "NORMAL": a_n, "TANGENT": a_t, "TEXCOORD_0": a_uv}, "indices": a_i, "material": 0}
    # mesh 0 / primitive 1: triangle strip, u8 indices, POSITION only
    sp = r.random((9, 3), dtype=np.float32)
    prim1 = {"attributes": {"POSITION": acc(add(sp, 0, 0), 5126, 9, "VEC3")},
             "indices": acc(add(np.arange(9, dtype=np.uint8), 0, 0), 5121, 9, "SCALAR"), "material": 1, "mode": 5}
    # mesh 1: triangle fan, u32 indices, interleaved POSITION + TEXCOORD_0 (stride 20)
    inter = np.zeros((8, 5), np.float32)
    inter[:, :3] = r.random((8, 3), dtype=np.float32) * 2 - 1
    inter[:, 3:] = r.random((8, 2), dtype=np.float32)
    vi = add(inter, 0, 0, stride=20)
    prim2 = {"attributes": {"POSITION": acc(vi, 5126, 8, "VEC3"), "TEXCOORD_0": acc(vi, 5126, 8, "VEC2", off=12)},
             "indices": acc(add(np.array([0, 1, 2, 3, 4, 5, 6, 7], np.uint32), 0, 0), 5125, 8, "SCALAR"), "material": 2, "mode": 6}
    # mesh 2: emissive triangle list with u32 indices and normals only
    p2, ix2 = grid(3)
    prim3 = {"attributes": {"POSITION": acc(add(p2 * np.float32(0.5), 0, 0), 5126, len(p2), "VEC3"),
                            "NORMAL": acc(add(r.standard_normal((len(p2), 3)).astype(np.float32), 0, 0), 5126, len(p2), "VEC3")},
             "indices": acc(add(ix2.astype(np.uint32), 0, 0), 5125, len(ix2), "SCALAR"), "material": 3}

    def mat(angle, axis, scale, t):
        c, s_ = np.cos(angle), np.sin(angle)
        R = np.eye(3)
        i, j = [(1, 2), (0, 2), (0, 1)][axis]
        R[i, i], R[i, j], R[j, i], R[j, j] = c, -s_, s_, c
        M = np.eye(4)
        M[:3, :3] = R * np.array(scale)
        M[:3, 3] = t
        return [float(x) for x in M.T.reshape(-1)]     # glTF stores column-major

    blob = b"".join(chunks)
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0, 2, 3]}],
         "nodes": [{"mesh": 0, "matrix": mat(0.4, 1, (1.5, 1.5, 1.5), (1, 2, 3)), "children": [1]},
                   {"mesh": 1, "matrix": mat(-0.9, 2, (0.5, 2.0, 1.0), (-1, 0.5, 0.25))},
                   {"mesh": 2, "translation": [5, 5, 5], "rotation": [0, 0.7071068, 0, 0.7071068], "scale": [2, 2, 2]},
                   {"mesh": 0, "matrix": mat(1.2, 0, (1, 1, 1), (0, -3, 0.5))}],
         "meshes": [{"primitives": [prim0, prim1]}, {"primitives": [prim2]}, {"primitives": [prim3]}],
         "materials": [{"pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.6, 0.4, 1], "metallicFactor": 0.25, "roughnessFactor": 0.6,
                                                 "baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 1}},
                        "normalTexture": {"index": 2}},
                       {"pbrMetallicRoughness": {"baseColorFactor": [0.1, 0.9, 0.2, 1], "roughnessFactor": 0.0}},
                       {"pbrMetallicRoughness": {}, "emissiveTexture": {"index": 1}},
                       {"pbrMetallicRoughness": {"baseColorFactor": [1, 1, 1, 1]}, "emissiveFactor": [4.0, 3.0, 2.5]}],
         "textures": [{"source": 0}, {"source": 1}, {"source": 0}],
         "images": [{"uri": "p_rgb.png"}, {"uri": "j420_37x23_q92.jpg"}],
         "buffers": [{"byteLength": len(blob), "uri": "features.bin"}], "bufferViews": views, "accessors": accs}
    with open(os.path.join(OUT, "features.bin"), "wb") as f:
        f.write(blob)
    with open(os.path.join(OUT, "features.gltf"), "w") as f:
        json.dump(g, f, indent=1)
    print("features.gltf:", len(accs), "accessors,", len(blob), "bytes")


if __name__ == "__main__":
    main()

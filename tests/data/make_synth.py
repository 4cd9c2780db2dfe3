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


if __name__ == "__main__":
    main()

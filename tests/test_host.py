"""CPU tests of the host front end and the C ABI surface (no GPU needed)."""
import base64
import ctypes as C
import hashlib
import io
import json
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import MEDIA, ROOT

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_host_golden.json")))


def _words(a):
    return [int(x) for x in np.ascontiguousarray(a).view(np.uint32).reshape(-1)]


@pytest.mark.parametrize("g", GOLDEN["scenes"], ids=lambda g: f"{g['file']}@{g['scale']}")
def test_loader_matches_reference_loader(gpurt, g):
    """glTF loader + Mat4/Pose/BBox math + Scene_Desc/Scene_Light packing == the reference's own
    Scene::load / RTPipe::build_desc, bit for bit (fixture from oracle/_ref, tests/golden/make_golden.py)"""
    base = os.path.join(ROOT, "tests", "data") if g["file"].startswith("synth/") else MEDIA
    s = gpurt.Scene(None).load(os.path.join(base, g["file"]), g["scale"])
    descs, lights = s.descs(), s.lights()
    assert len(descs) == len(g["objects"]) and len(lights) == len(g["lights"])
    assert s.counts()["textures"] == g["n_textures"]
    for i, (d, o) in enumerate(zip(descs, g["objects"])):
        v, ix = s.object(i)
        assert (v.shape[0], ix.size) == (o["n_verts"], o["n_indices"])
        assert hashlib.sha256(v.tobytes()).hexdigest() == o["verts_sha256"]
        assert hashlib.sha256(ix.tobytes()).hexdigest() == o["indices_sha256"]
        assert _words(np.array(d.model, np.float32)) == o["model"], f"object {i} model matrix"
        assert _words(np.array(d.modelIT, np.float32)) == o["modelIT"], f"object {i} modelIT"
        mat = list(d.albedo[:3]) + list(d.emissive[:3]) + list(d.metal_rough[:2])
        assert _words(np.array(mat, np.float32)) == o["material"]
        assert [d.albedo_tex, d.emissive_tex, d.metal_rough_tex, d.normal_tex] == o["textures"]
        assert d.index == i
    for l, o in zip(lights, g["lights"]):
        assert _words(np.array(l.bmin, np.float32)) == o["bmin"] and _words(np.array(l.bmax, np.float32)) == o["bmax"]
        assert (l.index, l.n_triangles) == (o["index"], o["n_triangles"])
    # unordered_map iteration order -> descending ids for these scenes
    assert [o["id"] for o in g["objects"]] == sorted([o["id"] for o in g["objects"]], reverse=True)
    offs = s.tri_offsets()
    assert offs[0] == 0 and (np.diff(offs) == [o["n_indices"] // 3 for o in g["objects"]]).all()


@pytest.mark.parametrize("c", GOLDEN["cameras"], ids=lambda c: f"cam{c['mode']}_{c['w']}x{c['h']}")
def test_camera_matches_reference_camera(gpurt, c):
    cam = gpurt.camera(c["mode"], c["w"], c["h"], c["pos"], c["center"], c["vfov"])
    got = _words(np.array(list(cam.V) + list(cam.P) + list(cam.iV) + list(cam.iP), np.float32))
    assert got == c["V_P_iV_iP"]


def test_scene_counts_of_the_named_configs(gpurt):
    assert gpurt.Scene(None).load(os.path.join(MEDIA, "cbox", "cbox.gltf")).counts() == \
        {"objs": 10, "tris": 16732, "lights": 1, "textures": 0}
    assert gpurt.Scene(None).load(os.path.join(MEDIA, "mis_test", "mis_test.gltf")).counts() == \
        {"objs": 7, "tris": 1544, "lights": 3, "textures": 0}
    assert gpurt.Scene(None).load(os.path.join(MEDIA, "cube.gltf")).counts()["tris"] == 12


def test_sponza_standin_is_deterministic_and_sized_like_sponza(gpurt):
    a, b = gpurt.Scene(None).make_sponza_standin(), gpurt.Scene(None).make_sponza_standin()
    assert a.counts() == {"objs": 103, "tris": 262267, "lights": 0, "textures": 0}
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for i in range(103):
        va, ia = a.object(i)
        vb, ib = b.object(i)
        assert (va == vb).all() and (ia == ib).all()
        lo, hi = np.minimum(lo, va[:, :3].min(0)), np.maximum(hi, va[:, :3].max(0))
        assert ia.max() < va.shape[0]
    assert np.allclose(lo, [-1921, -126, -1183]) and np.allclose(hi, [1800, 1429, 1105])


def test_errors_are_codes_not_exits(gpurt, tmp_path):
    with pytest.raises(gpurt.GpurtError) as e:
        gpurt.Scene(None).load(str(tmp_path / "missing.gltf"))
    assert e.value.code == -4
    bad = tmp_path / "bad.gltf"
    bad.write_text("{ not json")
    with pytest.raises(gpurt.GpurtError):
        gpurt.Scene(None).load(str(bad))
    s = gpurt.Scene(None)
    v = np.zeros((3, 12), np.float32)
    with pytest.raises(gpurt.GpurtError):
        s.add_object(v, np.array([0, 1, 3], np.uint32))      # index out of range
    s.add_object(v, np.array([0, 1, 2], np.uint32))
    # no CPU fallback: a scene without a context cannot be built, and without a GPU no context exists
    with pytest.raises(gpurt.GpurtError) as e:
        gpurt.Accel(s)
    assert e.value.code == -2
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(gpurt.GpurtError) as e:
            gpurt.Context(0)
        assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_hostile_gltf_numbers_are_rejected_not_followed(gpurt, tmp_path):
    """untrusted files: a cyclic node graph, accessor counts / offsets that overflow when multiplied, negative and
    non-finite numbers — the loader answers with an error code (or loads nothing), it does not read out of bounds or recurse
    until the stack ends; a material that names a missing texture falls back to its constant factor"""
    import base64
    import json
    buf = np.zeros(36, np.float32).tobytes() + np.arange(3, dtype=np.uint32).tobytes()
    uri = "data:application/octet-stream;base64," + base64.b64encode(buf).decode()

    def doc(**over):
        d = {"asset": {"version": "2.0"}, "buffers": [{"byteLength": len(buf), "uri": uri}],
             "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 144}, {"buffer": 0, "byteOffset": 144, "byteLength": 12}],
             "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                           {"bufferView": 1, "componentType": 5125, "count": 3, "type": "SCALAR"}],
             "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
             "nodes": [{"mesh": 0}], "scenes": [{"nodes": [0]}]}
        d.update(over)
        return d

    def load(d):
        p = tmp_path / "x.gltf"
        p.write_text(json.dumps(d))
        return gpurt.Scene(None).load(str(p))
    assert load(doc()).counts()["tris"] == 1
    with pytest.raises(gpurt.GpurtError):                                     # node 0 -> node 1 -> node 0 -> ...
        load(doc(nodes=[{"mesh": 0, "children": [1]}, {"children": [0]}]))
    for acc in ({"count": 1e30}, {"count": 2 ** 61}, {"byteOffset": -16}, {"byteOffset": 1e300}, {"count": -3}):
        d = doc()
        d["accessors"][0].update(acc)
        try:
            assert load(d).counts()["tris"] <= 1
        except gpurt.GpurtError:
            pass
    d = doc()
    d["bufferViews"][0]["byteStride"] = 2 ** 62
    try:
        load(d)
    except gpurt.GpurtError:
        pass
    s = gpurt.Scene(None)
    m = gpurt.Material()
    m.albedo[:] = (1, 1, 1)
    m.albedo_tex, m.emissive_tex, m.metal_rough_tex, m.normal_tex = 3, -1, 0, -1       # no textures in this scene
    s.add_object(np.zeros((3, 12), np.float32), np.arange(3, dtype=np.uint32), None, m)
    d0 = s.descs()[0]
    assert (d0.albedo_tex, d0.metal_rough_tex) == (-1, -1) and "texture" in gpurt.last_error()
    # nesting without end: the reader gives up at a fixed depth instead of recursing until the stack ends
    for text in ('{"asset":{"version":"2.0"},"x":' + "[" * 200000 + "]" * 200000 + "}", '{"a":' * 100000 + "1" + "}" * 100000):
        p = tmp_path / "deep.gltf"
        p.write_text(text)
        with pytest.raises(gpurt.GpurtError) as e:
            gpurt.Scene(None).load(str(p))
        assert e.value.code == -4
    # an index that is no int (1e30, -1e30) reads as "absent", not through an undefined float -> int conversion
    d = doc(materials=[{"pbrMetallicRoughness": {"baseColorTexture": {"index": 1e30}}, "normalTexture": {"index": -1e30}}])
    d["meshes"][0]["primitives"][0]["material"] = 0
    d0 = load(d).descs()[0]
    assert (d0.albedo_tex, d0.normal_tex) == (-1, -1)


def test_library_exports_every_declared_symbol(gpurt):
    header = open(os.path.join(ROOT, "include", "gpurt.h")).read()
    declared = set(re.findall(r"\b(gpurt_[a-z0-9_]+)\s*\(", header))
    assert declared == set(gpurt.SYMBOLS), declared ^ set(gpurt.SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", gpurt.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gpurt_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported
    for name in declared:
        assert hasattr(gpurt.lib, name)
    assert gpurt.lib.gpurt_version().startswith(b"gpurt-b200")
    # POD layouts promised by the header
    assert C.sizeof(gpurt.SceneDesc) == 208 and C.sizeof(gpurt.SceneLight) == 48 and C.sizeof(gpurt.Constants) == 88
    assert C.sizeof(gpurt.Camera) == 328 and gpurt.RAY_DT.itemsize == 32 and gpurt.HIT_DT.itemsize == 16
    assert gpurt.QUERY_DT.itemsize == 16 and gpurt.CPQ_DT.itemsize == 32


def test_product_does_not_link_or_load_the_oracle(gpurt):
    ldd = subprocess.run(["ldd", gpurt.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "emu" not in ldd
    for root, _, files in os.walk(os.path.join(ROOT, "gpu-rt_b200")):
        if "build" in root:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                src = open(os.path.join(root, f)).read()
                assert "liboracle" not in src and "orc_" not in src and "import orc" not in src, f


def _png(arr):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, format="PNG")
    return buf.getvalue()


def test_gltf_features_embedded_buffers_png_textures_glb_strip_fan(gpurt, tmp_path):
    """data: URIs, PNG decode, byteStride, u8/u16 indices, strip/fan conversion (scene.cpp:71-143), GLB"""
    rng = np.random.default_rng(3)
    rgba = rng.integers(0, 256, (5, 7, 4), dtype=np.uint8)
    rgb = rng.integers(0, 256, (4, 4, 3), dtype=np.uint8)
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 2, 0]], np.float32)
    inter = np.zeros((5, 5), np.float32)          # interleaved pos + uv, stride 20
    inter[:, :3] = pos
    inter[:, 3:] = rng.random((5, 2), dtype=np.float32)
    idx8 = np.array([0, 1, 2, 3, 4], np.uint8)     # strip: 3 triangles, fan: 3 triangles
    blob = inter.tobytes() + idx8.tobytes() + b"\0" * 3
    uri = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()
    def gltf(mode, buffers):
        return {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
                "nodes": [{"mesh": 0, "matrix": [2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0, 1, 2, 3, 1]}],
                "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "TEXCOORD_0": 1}, "indices": 2, "material": 0, "mode": mode}]}],
                "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "roughnessFactor": 0.25},
                               "emissiveTexture": {"index": 1}, "emissiveFactor": [1, 2, 3]}],
                "textures": [{"source": 0}, {"source": 1}],
                "images": [{"uri": "data:image/png;base64," + base64.b64encode(_png(rgba)).decode()},
                           {"uri": "data:image/png;base64," + base64.b64encode(_png(rgb)).decode()}],
                "buffers": buffers,
                "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 100, "byteStride": 20},
                                {"buffer": 0, "byteOffset": 100, "byteLength": 5}],
                "accessors": [{"bufferView": 0, "componentType": 5126, "count": 5, "type": "VEC3"},
                              {"bufferView": 0, "byteOffset": 12, "componentType": 5126, "count": 5, "type": "VEC2"},
                              {"bufferView": 1, "componentType": 5121, "count": 5, "type": "SCALAR"}]}
    expect = {5: [0, 1, 2, 1, 2, 3, 2, 3, 4], 6: [0, 1, 2, 0, 2, 3, 0, 3, 4], 4: [0, 1, 2]}
    for mode in (5, 6):
        p = tmp_path / f"m{mode}.gltf"
        p.write_text(json.dumps(gltf(mode, [{"byteLength": len(blob), "uri": uri}])))
        s = gpurt.Scene(None).load(str(p))
        v, ix = s.object(0)
        assert list(ix) == expect[mode]
        assert (v[:, :3] == pos).all() and (v[:, 3] == inter[:, 3]).all() and (v[:, 7] == inter[:, 4]).all()
        d = s.descs()[0]
        assert s.counts()["textures"] == 2 and (d.albedo_tex, d.emissive_tex, d.normal_tex) == (0, 1, -1)
        assert list(d.emissive[:3]) == [1, 2, 3] and d.metal_rough[1] == 0.25 and len(s.lights()) == 1
        m = np.array(d.model, np.float32).reshape(4, 4).T
        assert np.allclose(m[:3, 3], [1, 2, 3], atol=1e-5) and np.allclose(np.diag(m)[:3], 2, atol=1e-5)
    # GLB container with the binary chunk as buffer 0
    js = json.dumps(gltf(5, [{"byteLength": len(blob)}])).encode()
    js += b" " * (-len(js) % 4)
    bin_chunk = blob + b"\0" * (-len(blob) % 4)
    glb = b"glTF" + struct.pack("<II", 2, 12 + 8 + len(js) + 8 + len(bin_chunk)) + struct.pack("<II", len(js), 0x4E4F534A) + js \
        + struct.pack("<II", len(bin_chunk), 0x004E4942) + bin_chunk
    p = tmp_path / "m.glb"
    p.write_bytes(glb)
    s = gpurt.Scene(None).load(str(p))
    assert list(s.object(0)[1]) == expect[5] and s.counts()["textures"] == 2


def _texture_hashes(scene):
    return [(t.shape[1], t.shape[0], hashlib.sha256(t.tobytes()).hexdigest())
            for t in (scene.texture(i) for i in range(scene.counts()["textures"]))]


def test_texture_decode_matches_reference_decoder(gpurt):
    """JPEG (baseline + progressive, 4:4:4 / 4:2:2 / 4:2:0 / 4:1:1, gray, restart intervals, optimised tables) and
    PNG (RGB, RGBA, gray, gray+alpha, palette) decode to the bytes the reference's tinygltf -> stb_image path
    produced for the same files (tests/golden/make_golden.py)"""
    s = gpurt.Scene(None).load(os.path.join(ROOT, "tests", "data", "synth", "textures.gltf"))
    assert "[warn]" not in gpurt.last_error()
    got = _texture_hashes(s)
    assert len(got) == len(GOLDEN["synth_textures"]) == 22
    for g, (w, h, sha) in zip(GOLDEN["synth_textures"], got):
        assert (g["w"], g["h"], g["sha256"]) == (w, h, sha), g["file"]


def test_glb_embedded_images_match_reference_decoder(gpurt):
    """binary glTF with geometry and images (JPEG + PNG) in the BIN chunk, referenced through bufferViews"""
    s = gpurt.Scene(None).load(os.path.join(ROOT, "tests", "data", "synth", "embedded.glb"))
    assert "[warn]" not in gpurt.last_error()
    got = _texture_hashes(s)
    assert [(g["w"], g["h"], g["sha256"]) for g in GOLDEN["glb_textures"]] == got


@pytest.mark.skipif(not os.path.isdir("/root/reference/media/sponza"), reason="reference media not present")
def test_sponza_textures_match_reference_decoder(gpurt, tmp_path):
    """all 69 textures of media/sponza (65 baseline JPEG, 4 PNG; up to 2048^2) decode bit-identically to the reference"""
    src = "/root/reference/media/sponza"
    files = [g["file"] for g in GOLDEN["sponza_textures"]]
    for f in files:
        os.symlink(os.path.join(src, f), tmp_path / f)
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    blob = pos.tobytes() + np.array([0, 1, 2], np.uint16).tobytes() + b"\0\0"
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
         "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
         "textures": [{"source": i} for i in range(len(files))], "images": [{"uri": f} for f in files],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 6}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                       {"bufferView": 1, "componentType": 5123, "count": 3, "type": "SCALAR"}]}
    (tmp_path / "t.gltf").write_text(json.dumps(g))
    s = gpurt.Scene(None).load(str(tmp_path / "t.gltf"))
    got = _texture_hashes(s)
    assert len(got) == 69
    for gt, (w, h, sha) in zip(GOLDEN["sponza_textures"], got):
        assert (gt["w"], gt["h"], gt["sha256"]) == (w, h, sha), gt["file"]


def test_corrupt_images_fall_back_to_a_placeholder_with_a_warning(gpurt, tmp_path):
    """truncated / garbage JPEG and PNG never crash the loader: 1x1 white placeholder + [warn] in last_error"""
    good = open(os.path.join(ROOT, "tests", "data", "synth", "j420_37x23_q92.jpg"), "rb").read()
    variants = {"trunc.jpg": good[: len(good) // 2], "head.jpg": good[:40], "noise.jpg": good[:200] + bytes(range(256)) * 4,
                "soi_only.jpg": b"\xff\xd8\xff", "bad.png": b"\x89PNG\r\n\x1a\n" + b"\0" * 30}
    rng = np.random.default_rng(5)
    for k in range(12):      # random byte corruption inside the entropy-coded data
        b = bytearray(good)
        for p in rng.integers(300, len(b) - 2, 6):
            b[p] = int(rng.integers(0, 256))
        variants[f"fuzz{k}.jpg"] = bytes(b)
    files = sorted(variants)
    for f in files:
        (tmp_path / f).write_bytes(variants[f])
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    blob = pos.tobytes() + np.array([0, 1, 2], np.uint16).tobytes() + b"\0\0"
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
         "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
         "textures": [{"source": i} for i in range(len(files))], "images": [{"uri": f} for f in files],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 6}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                       {"bufferView": 1, "componentType": 5123, "count": 3, "type": "SCALAR"}]}
    (tmp_path / "t.gltf").write_text(json.dumps(g))
    s = gpurt.Scene(None).load(str(tmp_path / "t.gltf"))
    assert s.counts()["textures"] == len(files)
    assert "[warn]" in gpurt.last_error()
    for i, f in enumerate(files):
        t = s.texture(i)
        if f in ("head.jpg", "soi_only.jpg", "bad.png"):    # no image data at all
            assert t.shape == (1, 1, 4) and (t == 255).all(), f
        else:                                                # damaged data decodes to something of the right size, or not at all
            assert t.shape in ((1, 1, 4), (23, 37, 4)), f


def test_png_header_that_promises_more_than_its_data_is_refused_before_allocating(gpurt, tmp_path):
    """a PNG whose IHDR names 60000 x 60000 pixels over a few hundred IDAT bytes (found by tools/fuzz_loader.sh under
    AddressSanitizer: a 24 GB allocation) decodes to the placeholder; a short IHDR chunk is not read past its end; an object
    index of 2^32 - 1 is out of range, not index 0 after a wrap"""
    import struct
    import zlib
    good = open(os.path.join(ROOT, "tests", "data", "synth", "p_rgba.png"), "rb").read()

    def chunks(b):
        o = 8
        while o + 12 <= len(b):
            n = struct.unpack(">I", b[o:o + 4])[0]
            yield b[o + 4:o + 8], b[o + 8:o + 8 + n]
            o += 12 + n

    def png(parts):
        out = good[:8]
        for tag, data in parts:
            out += struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))
        return out
    parts = list(chunks(good))
    huge = [(t, struct.pack(">II", 60000, 60000) + d[8:]) if t == b"IHDR" else (t, d) for t, d in parts]
    short = [(t, d[:6]) if t == b"IHDR" else (t, d) for t, d in parts]
    files = {"good.png": good, "huge.png": png(huge), "short.png": png(short)[:8 + 12 + 6]}
    for f, b in files.items():
        (tmp_path / f).write_bytes(b)
    names = sorted(files)
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    blob = pos.tobytes() + np.array([0, 1, 2], np.uint32).tobytes()
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
         "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "indices": 1}]}],
         "textures": [{"source": i} for i in range(len(names))], "images": [{"uri": f} for f in names],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 12}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                       {"bufferView": 1, "componentType": 5125, "count": 3, "type": "SCALAR"}]}
    (tmp_path / "t.gltf").write_text(json.dumps(g))
    s = gpurt.Scene(None).load(str(tmp_path / "t.gltf"))
    shapes = {f: s.texture(i).shape for i, f in enumerate(names)}
    assert shapes["huge.png"] == (1, 1, 4) and shapes["short.png"] == (1, 1, 4) and shapes["good.png"] != (1, 1, 4), shapes
    import ctypes as C
    nv, ni = C.c_uint32(), C.c_uint32()
    assert gpurt.lib.gpurt_scene_object_sizes(s.h, 0, C.byref(nv), C.byref(ni)) == 0 and (nv.value, ni.value) == (3, 3)
    assert gpurt.lib.gpurt_scene_object_sizes(s.h, 0xFFFFFFFF, C.byref(nv), C.byref(ni)) == -1
    assert gpurt.lib.gpurt_scene_get_object(s.h, 0xFFFFFFFF, None, None) == -1


def test_header_is_plain_c_and_links(gpurt, tmp_path):
    """include/gpurt.h compiles as C99 with -pedantic, and a C program linked against libgpurt.so can call the
    host-only entry points (what a cgo / JNI / ctypes binding of the reference's maintainers would do)"""
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "gpurt.h"
int main(void) {
    GpurtPipeParams p;
    GpurtCamera cam;
    float pos[3] = {0, 0, 5}, at[3] = {0, 0, 0};
    gpurt_scene* s = 0;
    unsigned int no = 9, nt = 9, nl = 9, nx = 9;
    if(sizeof(GpurtRay) != 32 || sizeof(GpurtHit) != 16 || sizeof(GpurtQuery) != 16 || sizeof(GpurtClosestPoint) != 32) return 2;
    if(sizeof(GpurtSceneDesc) != 208 || sizeof(GpurtSceneLight) != 48 || sizeof(GpurtConstants) != 88 || sizeof(GpurtCamera) != 328) return 3;
    if(gpurt_pipe_params_default(&p) != GPURT_OK || p.max_frames != 256 || p.samples_per_frame != 8) return 4;
    if(gpurt_camera_make(1, 640.0f, 480.0f, pos, at, 60.0f, &cam) != GPURT_OK) return 5;
    if(gpurt_scene_create(0, &s) != GPURT_OK || gpurt_scene_make_sponza_standin(s) != GPURT_OK) return 6;
    if(gpurt_scene_counts(s, &no, &nt, &nl, &nx) != GPURT_OK || no != 103 || nt != 262267) return 7;
    if(gpurt_trace_closest(0, 0, 0, 0, GPURT_MEM_HOST) != GPURT_E_INVALID || !strlen(gpurt_last_error())) return 8;
    gpurt_scene_destroy(s);
    printf("%s\n", gpurt_version());
    return 0;
}
''')
    exe = tmp_path / "abi"
    lib_dir = os.path.join(ROOT, "gpu-rt_b200")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", lib_dir, "-lgpurt", f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "gpurt-b200" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_non_triangle_primitives_keep_their_slot_but_no_triangles(gpurt, tmp_path):
    """points / lines primitives: the object (and with it ids and object order) stays, with zero triangles and a
    warning — the reference keeps indices next to an empty vertex array and dies in the Vulkan upload"""
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [2, 1, 0], [2, 2, 0]], np.float32)
    blob = pos.tobytes() + np.arange(6, dtype=np.uint16).tobytes()
    prim = lambda mode: {"attributes": {"POSITION": 0}, "indices": 1, "material": 0, "mode": mode}
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
         "meshes": [{"primitives": [prim(0), prim(1), prim(4)]}], "materials": [{"pbrMetallicRoughness": {}}],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 72}, {"buffer": 0, "byteOffset": 72, "byteLength": 12}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 6, "type": "VEC3"},
                       {"bufferView": 1, "componentType": 5123, "count": 6, "type": "SCALAR"}]}
    p = tmp_path / "pl.gltf"
    p.write_text(json.dumps(g))
    s = gpurt.Scene(None).load(str(p))
    assert s.counts()["objs"] == 3 and s.counts()["tris"] == 2 and "[warn]" in gpurt.last_error()
    assert sorted(s.object(i)[1].size for i in range(3)) == [0, 0, 6]
    assert list(s.tri_offsets()) in ([0, 2, 2, 2], [0, 0, 2, 2], [0, 0, 0, 2])


def test_exr_writer_round_trip(gpurt, tmp_path):
    from exr_reader import read_exr as _read_exr
    """gpurt_write_exr (SURVEY §8f rank 1: PNG / EXR writer): the bits written are the bits read back, incl. inf / nan /
    denormals; errors are codes"""
    rng = np.random.default_rng(3)
    img = rng.standard_normal((37, 53, 4)).astype(np.float32) * 100
    img[0, 0] = [np.inf, -np.inf, np.nan, 1e-42]
    path = str(tmp_path / "a.exr")
    gpurt.write_exr(path, img)
    attrs, ch = _read_exr(path)
    assert sorted(ch) == ["A", "B", "G", "R"]
    for k, n in enumerate("RGBA"):
        assert (ch[n].view(np.uint32) == img[..., k].view(np.uint32)).all(), n
    assert os.path.getsize(path) < img.nbytes + 37 * 16 + 512
    with pytest.raises(gpurt.GpurtError):
        gpurt.write_exr(str(tmp_path / "no_such_dir" / "a.exr"), img)
    assert gpurt.lib.gpurt_write_exr(None, None, 0, 0) == -1

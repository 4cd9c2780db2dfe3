"""Drop-in boundary: gpu-rt_b200/host/vk_dropin.h gives VK::Accel / VK::RTPipe the reference's own
signatures (src/vk/vulkan.h:256-284, src/vk/rt.h:13-53), and tests/dropin/dropin_app.cpp holds GPURT's call sites
(src/gpurt.cpp:39-45, :216-241) as they stand in the reference.

CPU: the program compiles and links (a) against stand-in application types and (b), when /root/reference is mounted,
against the reference's REAL Scene / Object / VK::Mesh / Camera headers and code (oracle/_ref/libgpurt_ref.so); without a
GPU it must stop with GPURT_E_NO_DEVICE (no CPU fallback).  GPU: both programs render the Cornell box through those call
sites — including a pose edit that goes through TLAS.drop() + TLAS->recreate(BLAS, BLAS_T) — and the images equal the
ones rendered through the C ABI / Python binding bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from conftest import MEDIA, ROOT

PKG = os.path.join(ROOT, "gpu-rt_b200")
SRC = os.path.join(ROOT, "tests", "dropin", "dropin_app.cpp")
REF = "/root/reference"


def _build_stub_app(tmp):
    out = os.path.join(tmp, "dropin_app")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", out, SRC, "-I" + os.path.join(ROOT, "include"),
                           "-L" + PKG, "-lgpurt", "-Wl,-rpath," + PKG])
    return out


def _ref_app():
    """oracle/_ref/dropin_ref_app (reference's own host code + call sites); rebuilt here when the reference is mounted"""
    path = os.path.join(ROOT, "oracle", "_ref", "dropin_ref_app")
    if os.path.isdir(os.path.join(REF, "src")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref/dropin_ref_app"], stdout=subprocess.DEVNULL)
    return path if os.path.exists(path) else None


def _run(app, out, w, h, frames, integ, brdf, spp, depth, seed, edit=(-1, 0.0)):
    cmd = [app, os.path.join(MEDIA, "cbox", "cbox.gltf"), out] + [str(x) for x in (w, h, frames, integ, brdf, spp, depth, seed, edit[0], edit[1])]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_call_sites_compile_unchanged_and_fail_loudly_without_a_device(built, tmp_path):
    import torch
    apps = [_build_stub_app(str(tmp_path))]
    ref = _ref_app()
    if os.path.isdir(os.path.join(REF, "src")):
        assert ref, "the reference is mounted but oracle/_ref/dropin_ref_app did not build"
    if ref:
        apps.append(ref)
    for app in apps:
        r = _run(app, str(tmp_path / "o.f32"), 32, 18, 1, 0, 0, 1, 2, 1)
        if torch.cuda.is_available():
            assert r.returncode == 0, r.stderr
        else:
            assert r.returncode == 42 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


def test_the_header_takes_its_signatures_from_the_reference():
    """every member the reference declares on VK::Accel / VK::RTPipe (vulkan.h:256-271, rt.h:15-53) is declared in
    vk_dropin.h with the same spelling (checked textually, so it also runs where the reference is not mounted)"""
    text = open(os.path.join(PKG, "host", "vk_dropin.h")).read()
    for decl in ("Accel(const Mesh& mesh)", "Accel(const std::vector<Drop<Accel>>& blas, const std::vector<Mat4>& inst)",
                 "void recreate(const Mesh& mesh)", "void recreate(const std::vector<Drop<Accel>>& blas, const std::vector<Mat4>& inst)",
                 "void recreate(const Drop<Accel>& blas, Mat4 inst)", "void destroy()",
                 "RTPipe(const Scene& scene)", "void recreate(const Scene& scene)", "void recreate_swap(const Scene&",
                 "void update_uniforms(const Camera& cam)", "void use_accel(const Accel& tlas)", "void use_image(const View&",
                 "void reset_frame()", "bool trace(const Camera& cam, VkCommandBuffer&", "VkExtent2D ext)",
                 "int max_frames = 256;", "int samples_per_frame = 8;", "int max_depth = 8;", "Vec3 clear = Vec3{0.3f};",
                 "Vec3 env = Vec3{1.0f};", "float env_scale = 0.0f;", "bool use_normal_map = false;", "bool use_rr = true;",
                 "bool use_metalness = false;", "bool use_qmc = false;", "bool use_temporal = true;", "int integrator = 0;",
                 "int temporal_scale = 16;", "int brdf = 0;", "int debug_view = 0;", "int res_samples = 4;"):
        assert decl in text, f"vk_dropin.h lacks `{decl}`"
    if os.path.isdir(os.path.join(REF, "src")):
        rt = open(os.path.join(REF, "src", "vk", "rt.h")).read()
        for line in rt.split("Drop<PipeData> pipe;")[1].split("Drop<Image> pos_image")[0].strip().splitlines():
            assert line.strip() in text, f"tunable `{line.strip()}` of rt.h:38-53 is not in vk_dropin.h"


@pytest.mark.gpu
def test_reference_call_sites_render_the_same_image_as_the_c_abi(gpurt, ctx, tmp_path):
    w, h, frames = 160, 90, 3
    kw = dict(integrator=2, brdf=1, samples_per_frame=2, max_depth=4, seed=11, max_frames=1 << 20)
    edit_obj, dx = 4, 0.3
    # the same sequence through the Python binding: 3 calls, pose edit (set_transform + update + reset_frame), 3 calls
    scene = gpurt.Scene(ctx).load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    prm, cam = gpurt.pipe_params(**kw), gpurt.camera(0, w, h)
    for _ in range(frames):
        assert pipe.render_frame(prm, cam, w, h) == 0
    want = [pipe.read_image().copy()]
    m = np.array(list(scene.descs()[edit_obj].model), np.float32)
    m[12] = np.float32(m[12]) + np.float32(dx)
    scene.set_transform(edit_obj, m)
    accel.update()
    pipe.reset_frame()
    for _ in range(frames):
        assert pipe.render_frame(prm, cam, w, h) == 0
    want.append(pipe.read_image().copy())
    pipe.close(), accel.close(), scene.close()

    apps = {"stand-in application types": _build_stub_app(str(tmp_path))}
    if _ref_app():
        apps["the reference's own Scene / Camera / Mesh code"] = _ref_app()
    for what, app in apps.items():
        for k, edit in enumerate(((-1, 0.0), (edit_obj, dx))):
            out = str(tmp_path / f"o{k}.f32")
            r = _run(app, out, w, h, frames, kw["integrator"], kw["brdf"], kw["samples_per_frame"], kw["max_depth"], kw["seed"], edit)
            assert r.returncode == 0, r.stderr
            got = np.fromfile(out, np.float32).reshape(h, w, 4)
            assert (got.view(np.uint32) == want[k].view(np.uint32)).all(), f"{what}: image {k} differs from the C ABI render"

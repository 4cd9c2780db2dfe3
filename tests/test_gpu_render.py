"""Integrator parity: the wavefront CUDA integrator (through the C ABI) vs the CPU restatement of
rt.rgen, frame by frame, at equal seed and spp.

The oracle is fed the exact push constants / UBO words the product used for each frame
(gpurt_pipe_last_uniforms), so this checks the device code; the host-side frame logic
(RTPipe::update_uniforms / trace) is checked separately below.

Stated tolerance: images must agree to per-pixel RMSE <= 1e-6 relative to the mean radiance and at
most 1e-4 of the pixels may differ at all.  (With the shared deterministic sin/cos/pow and the fp32
contract N8 the expectation is bit-exact; the tolerance only leaves room for fp ties in the
traversal)
"""
import numpy as np
import pytest

from scenes import load_scene

pytestmark = pytest.mark.gpu

MAX_DIFF_FRAC = 1e-4
MAX_REL_RMSE = 1e-6


def _compare(name, got, ref):
    got, ref = np.asarray(got, np.float32), np.asarray(ref, np.float32)
    same = (got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))
    diff_px = (~same.reshape(-1, got.shape[-1]).all(axis=1)).mean()
    d = np.nan_to_num(got.astype(np.float64) - ref.astype(np.float64))
    rmse = np.sqrt((d ** 2).mean()) / max(1e-12, np.abs(np.nan_to_num(ref)).mean())
    assert diff_px <= MAX_DIFF_FRAC and rmse <= MAX_REL_RMSE, f"{name}: {diff_px:.2e} of pixels differ, rel RMSE {rmse:.2e}"
    return diff_px


def _run(gpurt, orc, ctx, scene_name, w, h, frames, cam=None, textures=(), scene=None, **params):
    scene = scene or load_scene(gpurt, ctx, scene_name)
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    if not len(textures):   # a loaded scene brings its own (media/sponza through GPURT_SPONZA_GLTF)
        textures = [scene.texture(i) for i in range(scene.counts()["textures"])]
    rs = orc.RenderScene(scene, textures)
    st = orc.FrameState(w, h)
    prm = gpurt.pipe_params(**params)
    spatial = (params.get("spatial_samples", 0), params.get("spatial_radius", 16.0))
    cam = cam or gpurt.camera(0, w, h)
    worst = 0.0
    for f in range(frames):
        assert pipe.render_frame(prm, cam, w, h) == 0
        consts, ubo, seed_word = pipe.last_uniforms()
        # frame counter: 0, 0, 1, 2, ... (the second call resets: old_cam was never assigned, rt.cpp:132-135)
        assert consts[8] == max(0, f - 1)
        counts = orc.render_frame(rs, st, consts, ubo, seed_word ^ int(consts[8]), spatial=spatial,
                                  light_sampling=params.get("light_sampling", 0))  # oracle xors the frame itself
        worst = max(worst, _compare(f"{scene_name} frame {f} image", pipe.read_image(), st.image))
        for g in range(3):
            _compare(f"{scene_name} frame {f} gbuffer {g}", pipe.read_gbuffer(g), st.gb[st.parity ^ 1][g])
        if params.get("integrator", 0) in (3, 4):
            _compare(f"{scene_name} frame {f} reservoirs", pipe.read_reservoirs().view(np.float32),
                     st.res[st.parity ^ 1].view(np.float32))
        assert pipe.ray_counts() == (int(counts[0]), int(counts[1])), "ray counts differ from the oracle"
    img = pipe.read_image()
    assert np.isfinite(img[..., :3]).mean() > 0.999
    pipe.close(), accel.close()
    return img


@pytest.mark.parametrize("integrator", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("brdf", [0, 1])
def test_cbox_all_integrators(gpurt, orc, ctx, integrator, brdf):
    img = _run(gpurt, orc, ctx, "cbox", 192, 108, 3, integrator=integrator, brdf=brdf, samples_per_frame=2,
               max_depth=4, seed=1234 + integrator)
    assert img[..., :3].mean() > 0.01


def test_mis_test_scene_mis_and_restir(gpurt, orc, ctx):
    # scene spans x[-0.14,1.10] y[0,1.16] z[-1,1]
    cam = gpurt.camera(1, 160, 90, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    _run(gpurt, orc, ctx, "mis_test", 160, 90, 2, cam=cam, integrator=2, brdf=1, samples_per_frame=1, max_depth=4, seed=7)
    _run(gpurt, orc, ctx, "mis_test", 160, 90, 8, cam=cam, integrator=3, brdf=0, samples_per_frame=1, max_depth=4,
         res_samples=4, use_temporal=1, temporal_scale=16, seed=8)
    _run(gpurt, orc, ctx, "mis_test", 160, 90, 4, cam=cam, integrator=4, brdf=1, samples_per_frame=2, max_depth=4, seed=9)


def test_sponza_standin_config2(gpurt, orc, ctx):
    """BASELINE config 2 at reduced resolution: integrator 1, GGX, depth 2, 1 spp, env light, no RR"""
    cam = gpurt.camera(1, 320, 180, (-1000.0, 200.0, 0.0), (0.0, 200.0, 0.0), 90.0)
    _run(gpurt, orc, ctx, "sponza", 320, 180, 2, cam=cam, integrator=1, brdf=1, samples_per_frame=1,
         max_depth=2, use_rr=0, env_scale=1.0, seed=3)


def test_shadow_queue_stage_is_bit_identical(gpurt, orc, ctx, monkeypatch):
    """GPURT_SHADOW_QUEUE=1 (opt-in; measured slower than the inline trace): integrator 0 hands its shadow segments to a
    queue, k_shadow_resolve traces them and finishes the pixels — same frames, same ray counts"""
    monkeypatch.setenv("GPURT_SHADOW_QUEUE", "1")
    for brdf in (0, 1):
        _run(gpurt, orc, ctx, "cbox", 192, 108, 3, integrator=0, brdf=brdf, samples_per_frame=3, max_depth=4, seed=77)
    _run(gpurt, orc, ctx, "cbox", 128, 72, 2, integrator=0, brdf=0, samples_per_frame=1, max_depth=1, use_rr=0, seed=12)


def test_config2_full_size_1080p(gpurt, orc, ctx):
    """BASELINE config 2 at its full size: 1920x1080, integrator 1 (Material), GGX, depth 2, 1 spp, env light,
    no RR — image, G-buffers and ray counts of the frame against the oracle's rt.rgen restatement (rt.rgen:567-677)"""
    cam = gpurt.camera(1, 1920, 1080, (-1000.0, 200.0, 0.0), (0.0, 200.0, 0.0), 90.0)
    _run(gpurt, orc, ctx, "sponza", 1920, 1080, 2, cam=cam, integrator=1, brdf=1, samples_per_frame=1,
         max_depth=2, use_rr=0, env_scale=1.0, seed=3)


def test_config3_full_size_1080p(gpurt, orc, ctx):
    """BASELINE config 3 at its full size: mis_test at 1920x1080, depth 4, 1 spp — MIS (integrator 2), then 8 frames of
    ReSTIR direct (3) and 4 of ReSTIR (4) with temporal reuse (res_samples 4, temporal_scale 16); image, G-buffers,
    reservoirs and ray counts of every frame against the oracle"""
    cam = gpurt.camera(1, 1920, 1080, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    _run(gpurt, orc, ctx, "mis_test", 1920, 1080, 2, cam=cam, integrator=2, brdf=1, samples_per_frame=1, max_depth=4, seed=7)
    _run(gpurt, orc, ctx, "mis_test", 1920, 1080, 9, cam=cam, integrator=3, brdf=1, samples_per_frame=1, max_depth=4,
         res_samples=4, use_temporal=1, temporal_scale=16, seed=8)
    _run(gpurt, orc, ctx, "mis_test", 1920, 1080, 5, cam=cam, integrator=4, brdf=1, samples_per_frame=1, max_depth=4,
         res_samples=4, use_temporal=1, temporal_scale=16, seed=9)


def test_restir_spatial_reuse_extension(gpurt, orc, ctx):
    """GpurtPipeParams::spatial_samples / spatial_radius (extension, off by default): CUDA frames ==
    the oracle's restatement of the extension over frames with temporal + spatial reuse (image, G-buffers, reservoirs,
    ray counts); and the estimator it changes stays an estimator of the same image: against a converged render of the
    reference's estimator (MIS, 256 spp) the mean of 8 frames with spatial reuse is not further away than the
    reference's ReSTIR mean by more than 10 % RMSE, and its mean radiance agrees within 3 %"""
    for integ in (3, 4):
        _run(gpurt, orc, ctx, "cbox", 160, 90, 5, integrator=integ, brdf=1, samples_per_frame=2, max_depth=3, res_samples=4,
             use_temporal=1, temporal_scale=16, seed=40 + integ, spatial_samples=4, spatial_radius=8.0)
    cam = gpurt.camera(1, 160, 90, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    _run(gpurt, orc, ctx, "mis_test", 160, 90, 4, cam=cam, integrator=3, brdf=0, samples_per_frame=1, max_depth=4,
         res_samples=4, use_temporal=1, temporal_scale=16, seed=8, spatial_samples=3, spatial_radius=20.0)

    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    w, h = 160, 90

    def mean_image(frames, **params):
        pipe = gpurt.RTPipe(scene, accel)
        prm = gpurt.pipe_params(**params)
        for _ in range(frames + 1):   # the pipe's start-up renders frame 0 twice (rt.cpp:132-135): accumulation restarts
            assert pipe.render_frame(prm, gpurt.camera(0, w, h), w, h) == 0
        img = pipe.read_image()[..., :3].astype(np.float64)
        pipe.close()
        return img

    ref = mean_image(32, integrator=2, brdf=0, samples_per_frame=8, max_depth=1, seed=5)
    base = mean_image(8, integrator=3, brdf=0, samples_per_frame=1, max_depth=1, res_samples=4, use_temporal=1, temporal_scale=16, seed=6)
    spat = mean_image(8, integrator=3, brdf=0, samples_per_frame=1, max_depth=1, res_samples=4, use_temporal=1, temporal_scale=16, seed=6,
                      spatial_samples=4, spatial_radius=8.0)
    accel.close()
    rm = lambda a: float(np.sqrt(((a - ref) ** 2).mean()))
    assert np.isfinite(spat).all() and not np.array_equal(base, spat)
    assert abs(spat.mean() - ref.mean()) <= 0.03 * ref.mean(), (spat.mean(), base.mean(), ref.mean())
    assert rm(spat) <= 1.10 * rm(base), (rm(spat), rm(base))


def test_light_sampling_extension(gpurt, orc, ctx):
    """GpurtPipeParams::light_sampling = 1 (extension, off by default): light triangles chosen in
    proportion to area x luma(emissive), light_pdf weighted to match.  CUDA frames == the oracle's restatement for the
    direct, MIS and ReSTIR integrators (image, G-buffers, reservoirs, ray counts); and it is an estimator of the same
    image: against a converged MIS render of the reference's estimator, the 8-frame mean of the direct integrator with the
    flag has a mean radiance within 3 % (MIS with the weighted light_pdf: 2 %) and noise of the same order"""
    cam = gpurt.camera(1, 160, 90, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    for integ, frames in ((0, 2), (2, 2), (3, 4), (4, 3)):
        _run(gpurt, orc, ctx, "mis_test", 160, 90, frames, cam=cam, integrator=integ, brdf=1, samples_per_frame=2, max_depth=3,
             res_samples=4, use_temporal=1, temporal_scale=16, seed=50 + integ, light_sampling=1)
    _run(gpurt, orc, ctx, "cbox", 160, 90, 2, integrator=2, brdf=0, samples_per_frame=1, max_depth=4, seed=3, light_sampling=1)

    scene = load_scene(gpurt, ctx, "mis_test")
    accel = gpurt.Accel(scene)
    w, h = 160, 90

    def mean_image(frames, **params):
        pipe = gpurt.RTPipe(scene, accel)
        prm = gpurt.pipe_params(**params)
        for _ in range(frames + 1):   # the pipe's start-up renders frame 0 twice (rt.cpp:132-135): accumulation restarts
            assert pipe.render_frame(prm, cam, w, h) == 0
        img = pipe.read_image()[..., :3].astype(np.float64)
        pipe.close()
        return img

    ref = mean_image(64, integrator=2, brdf=0, samples_per_frame=8, max_depth=1, seed=5)
    base = mean_image(8, integrator=0, brdf=0, samples_per_frame=1, max_depth=1, seed=6)
    powr = mean_image(8, integrator=0, brdf=0, samples_per_frame=1, max_depth=1, seed=6, light_sampling=1)
    mis1 = mean_image(64, integrator=2, brdf=0, samples_per_frame=8, max_depth=1, seed=7, light_sampling=1)
    accel.close()
    rm = lambda a: float(np.sqrt(((a - ref) ** 2).mean()))
    assert np.isfinite(powr).all() and not np.array_equal(base, powr)
    assert abs(powr.mean() - ref.mean()) <= 0.03 * ref.mean(), (powr.mean(), base.mean(), ref.mean())
    assert abs(mis1.mean() - ref.mean()) <= 0.02 * ref.mean(), (mis1.mean(), ref.mean())   # MIS with the weighted light_pdf
    # power-proportional sampling is not uniformly better (it starves small emitters next to the surfaces they light); what
    # must hold is that it stays an estimator of the same image with noise of the same order
    print(f"light_sampling RMSE vs converged reference: uniform {rm(base):.5f}, by power {rm(powr):.5f}; means {base.mean():.5f} {powr.mean():.5f} {ref.mean():.5f}")
    assert rm(powr) <= 2.0 * rm(base), (rm(powr), rm(base))


def test_options_qmc_metalness_rr_off_depth1(gpurt, orc, ctx):
    _run(gpurt, orc, ctx, "cbox", 128, 72, 3, integrator=1, brdf=1, samples_per_frame=3, max_depth=5, use_qmc=1,
         use_metalness=1, seed=11)
    _run(gpurt, orc, ctx, "cbox", 128, 72, 2, integrator=0, brdf=0, samples_per_frame=1, max_depth=1, use_rr=0, seed=12)
    _run(gpurt, orc, ctx, "cbox", 128, 72, 3, integrator=4, brdf=0, samples_per_frame=2, max_depth=3, debug_view=2, seed=13)
    _run(gpurt, orc, ctx, "cbox", 128, 72, 2, integrator=3, brdf=1, samples_per_frame=2, max_depth=3, use_temporal=0, seed=14)


def test_textured_scene(gpurt, orc, ctx):
    """albedo / emissive / metal-rough / normal textures, sRGB decode, bilinear + repeat"""
    rng = np.random.default_rng(5)
    texs = [rng.integers(0, 256, (16, 16, 4), dtype=np.uint8), rng.integers(0, 256, (8, 32, 4), dtype=np.uint8),
            rng.integers(64, 256, (4, 4, 4), dtype=np.uint8), rng.integers(100, 156, (32, 32, 4), dtype=np.uint8)]
    texs[3][..., 2] = 250  # normals mostly +z
    scene = gpurt.Scene(ctx)
    for t in texs:
        scene.add_texture(t)

    def quad(z, size, mat, uvscale=3.0):
        v = np.zeros((4, 12), np.float32)
        v[:, 0:3] = [[-size, -size, z], [size, -size, z], [size, size, z], [-size, size, z]]
        v[:, 3] = np.array([0, 1, 1, 0]) * uvscale - 0.7   # u (exercises REPEAT with negatives)
        v[:, 7] = np.array([0, 0, 1, 1]) * uvscale - 0.3   # v
        v[:, 4:7] = [0, 0, 1]
        v[:, 8:12] = [1, 0, 0, 1]
        scene.add_object(v, np.array([0, 1, 2, 0, 2, 3], np.uint32), None, mat)

    m = gpurt.Material()
    m.albedo[:] = (1, 1, 1)
    m.albedo_tex, m.emissive_tex, m.metal_rough_tex, m.normal_tex = 0, -1, 2, 3
    m.metal_rough[:] = (0.5, 0.5)
    quad(0.0, 2.0, m)
    e = gpurt.Material()
    e.albedo[:] = (1, 1, 1)
    e.emissive[:] = (4, 4, 4)
    e.albedo_tex, e.emissive_tex, e.metal_rough_tex, e.normal_tex = -1, 1, -1, -1
    e.metal_rough[:] = (0, 1)
    quad(3.0, 0.8, e, uvscale=1.0)
    cam = gpurt.camera(1, 128, 96, (1.5, 1.0, 2.5), (0.0, 0.0, 0.5), 70.0)
    for integ in (0, 1, 2, 4):
        _run(gpurt, orc, ctx, "textured", 128, 96, 2, cam=cam, textures=texs, scene=scene, integrator=integ, brdf=1,
             samples_per_frame=2, max_depth=3, use_normal_map=1, use_metalness=1, seed=20 + integ)
    scene.close()


def test_rtpipe_frame_logic(gpurt, ctx):
    """RTPipe::update_uniforms / trace host semantics (rt.cpp:121-138, :346-398)"""
    scene = load_scene(gpurt, ctx, "cube")
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    w, h = 64, 48
    prm = gpurt.pipe_params(max_frames=3, samples_per_frame=1, max_depth=2, integrator=1, env_scale=1.0)
    cam = gpurt.camera(1, w, h, (3, 2, 4), (0, 0, 0), 60.0)
    assert pipe.frame_index() == -1
    assert pipe.render_frame(prm, cam, w, h) == 0 and pipe.frame_index() == 0
    c, ubo, _ = pipe.last_uniforms()
    ident = np.eye(4, dtype=np.float32).reshape(-1)
    assert (ubo[64:80].view(np.float32) == ident).all(), "prev_PV is identity until old_cam is first assigned"
    # second call: camera differs from old_cam (never assigned) -> reset_frame, frame 0 again (rt.cpp:132-135)
    assert pipe.render_frame(prm, cam, w, h) == 0 and pipe.frame_index() == 0
    assert pipe.render_frame(prm, cam, w, h) == 0 and pipe.frame_index() == 1
    _, ubo, _ = pipe.last_uniforms()
    P, V = np.array(cam.P, np.float64).reshape(4, 4).T, np.array(cam.V, np.float64).reshape(4, 4).T
    assert np.allclose(ubo[64:80].view(np.float32).reshape(4, 4).T, P @ V, rtol=1e-5, atol=1e-5)
    assert pipe.render_frame(prm, cam, w, h) == 0 and pipe.frame_index() == 2
    assert pipe.render_frame(prm, cam, w, h) == 0 and pipe.frame_index() == 3
    before = pipe.read_image().copy()
    assert pipe.render_frame(prm, cam, w, h) == 1, "frame >= max_frames: trace() returns false, nothing rendered"
    assert (pipe.read_image() == before).all()
    cam2 = gpurt.camera(1, w, h, (3, 2.5, 4), (0, 0, 0), 60.0)
    assert pipe.render_frame(prm, cam2, w, h) == 0 and pipe.frame_index() == 0, "camera change resets accumulation"
    pipe.reset_frame()
    assert pipe.frame_index() == -1
    # progressive accumulation converges: mean of frames == running mix (rt.rgen:638-645)
    prm2 = gpurt.pipe_params(max_frames=64, samples_per_frame=1, max_depth=2, integrator=1, env_scale=1.0)
    pipe.render_frame(prm2, cam2, w, h)
    imgs = []
    for _ in range(4):
        pipe.render_frame(prm2, cam2, w, h)
        imgs.append(pipe.read_image().copy())
    assert np.isfinite(imgs[-1]).all()
    rgba8 = pipe.tonemap(1, 1.0, 2.2)
    assert rgba8.shape == (h, w, 4) and rgba8[..., 3].min() == 255
    pipe.close(), accel.close(), scene.close()


def test_sharded_frames_compose_to_the_unsharded_frame(gpurt, ctx):
    """interleaved row bands per rank; RNG keyed by the global pixel -> rank-count invariant"""
    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    w, h = 160, 100   # 100 rows: the last 16-row band is partial
    cam = gpurt.camera(0, w, h)
    prm = gpurt.pipe_params(integrator=2, brdf=1, samples_per_frame=2, max_depth=4, seed=5)

    def frames(pipe, n=3):
        for _ in range(n):
            pipe.render_frame(prm, cam, w, h)
        return pipe.read_image(), [pipe.read_gbuffer(g) for g in range(3)]

    whole = gpurt.RTPipe(scene, accel)
    ref_img, ref_gb = frames(whole)
    for world in (2, 3, 8):
        img = np.zeros_like(ref_img)
        gb = [np.zeros_like(g) for g in ref_gb]
        rows = np.arange(h)
        for rank in range(world):
            p = gpurt.RTPipe(scene, accel)
            p.set_shard(16, world, rank)
            i, g = frames(p)
            mine = (rows // 16) % world == rank
            img[mine] = i[mine]
            for k in range(3):
                gb[k][mine] = g[k][mine]
            assert (i[~mine] == 0).all(), "a shard must not touch other ranks' rows"
            p.close()
        assert (img.view(np.uint32) == ref_img.view(np.uint32)).all(), f"{world} shards"
        for k in range(3):
            assert (gb[k].view(np.uint32) == ref_gb[k].view(np.uint32)).all()
    whole.set_shard(0, 1, 0)
    whole.close(), accel.close(), scene.close()


def _restir_sharded(*argv, nproc=0):
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    tool = os.path.join(ROOT, "tools", "restir_sharded.py")
    cmd = [sys.executable, tool] if not nproc else [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                                                    str(nproc), "--master-addr", "127.0.0.1", "--master-port", "29741", tool]
    out = subprocess.run(cmd + ["--verify", "--reps", "1"] + list(argv), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-2500:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["flag_wait_timeouts"] == 0 and all(res["verified"].values()), res
    return res


def test_restir_sharded_frame_with_history_exchange(gpurt, built):
    """SURVEY §8e, ReSTIR: shards store their rows of the finished frame into each other's previous-frame blocks
    (k_history_push / signal / wait); image, G-buffers and reservoirs after every frame == the unsharded render, with a
    moving camera (whole rows to everyone) and with a halo (static camera, spatial reuse inside the halo).  One process,
    the shards' pipes on one GPU — the two-GPU test below runs the same thing over NVLink."""
    r = _restir_sharded("--shards", "3", "--size", "320", "180", "--frames", "5", "--bands", "interleaved", "--orbit", "--integrator", "4")
    assert r["n_shards"] == 3 and r["bytes_pushed_per_frame_shard0"] > 0
    _restir_sharded("--shards", "4", "--size", "320", "200", "--frames", "4", "--halo", "10", "--spatial", "3", "--radius", "8")
    _restir_sharded("--shards", "2", "--size", "1920", "1080", "--frames", "3", "--orbit")


def test_restir_sharded_frame_two_gpus(gpurt, built):
    """the same over NVLink: one rank per GPU, blocks mapped through gpurt_shared_open (needs 2 GPUs)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _restir_sharded("--frames", "6", "--orbit", nproc=2)
    _restir_sharded("--frames", "6", "--halo", "16", "--spatial", "2", "--radius", "8", nproc=2)


def test_tonemap_matches_oracle(gpurt, orc, ctx):
    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    w, h = 96, 54
    pipe.render_frame(gpurt.pipe_params(samples_per_frame=2, max_depth=3, integrator=0), gpurt.camera(0, w, h), w, h)
    img = pipe.read_image()
    for op, exposure, gamma in [(0, 1.0, 2.2), (1, 1.0, 2.2), (1, 2.5, 1.8), (2, 1.0, 1.0)]:
        assert (pipe.tonemap(op, exposure, gamma) == orc.tonemap(img, op, exposure, gamma)).all(), (op, exposure, gamma)
    pipe.close(), accel.close(), scene.close()


def test_gltf_feature_scene_with_decoded_textures(gpurt, orc, ctx):
    """tests/data/synth/features.gltf end to end: loader (nested matrices, strip / fan, instancing), PNG + JPEG
    texture decode, albedo / metal-rough / normal / emissive textures, an emissive light — rendered by every
    integrator and compared with the oracle, which is given the decoded texels of the product's loader
    (those are pinned against the reference's decoder in tests/test_host.py)"""
    import os
    from conftest import ROOT
    path = os.path.join(ROOT, "tests", "data", "synth", "features.gltf")
    cam = gpurt.camera(1, 160, 120, (4.0, 3.0, 6.0), (1.0, 1.0, 2.0), 60.0)
    for integ in (0, 1, 2, 4):
        scene = gpurt.Scene(ctx).load(path)
        texs = [scene.texture(i) for i in range(scene.counts()["textures"])]
        assert len(texs) == 3 and scene.counts()["lights"] == 2
        img = _run(gpurt, orc, ctx, "features", 160, 120, 2, cam=cam, textures=texs, scene=scene, integrator=integ,
                   brdf=integ % 2, samples_per_frame=2, max_depth=3, use_normal_map=1, use_metalness=1, env_scale=0.5,
                   seed=77 + integ)
        assert img[..., :3].mean() > 0.001
        scene.close()


def test_headless_cli_writes_the_same_image_as_the_api(gpurt, orc, ctx, tmp_path):
    """gpurt_render (the headless replacement of GPURT::loop + save_rt, src/gpurt.cpp:258-262) renders cbox until
    trace() reports convergence and writes a PNG; its pixels equal the tonemapped image of the same sequence of
    frames driven through the Python binding, which equals the oracle's tonemap of the oracle's image"""
    import os
    import subprocess
    from conftest import MEDIA, ROOT
    Image = pytest.importorskip("PIL.Image")
    exe = os.path.join(ROOT, "gpu-rt_b200", "gpurt_render")
    out = str(tmp_path / "cli.png")
    w, h, frames, spp = 160, 96, 3, 2
    r = subprocess.run([exe, "-s", os.path.join(MEDIA, "cbox", "cbox.gltf"), "-o", out, "--size", str(w), str(h), "--frames", str(frames),
                        "--spp", str(spp), "--depth", "4", "--integrator", "2", "--brdf", "1", "--seed", "9", "--tonemap", "1",
                        "--exposure", "1.5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "16732 tris" in r.stdout
    cli = np.array(Image.open(out).convert("RGBA"))
    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    rs, st = orc.RenderScene(scene), orc.FrameState(w, h)
    prm = gpurt.pipe_params(max_frames=frames, samples_per_frame=spp, max_depth=4, integrator=2, brdf=1, seed=9)
    cam = gpurt.camera(0, w, h)
    n = 0
    while pipe.render_frame(prm, cam, w, h) == 0:
        consts, ubo, seed_word = pipe.last_uniforms()
        orc.render_frame(rs, st, consts, ubo, seed_word ^ int(consts[8]))
        n += 1
    assert n == frames + 2                 # frames 0, 0, 1, ..., max_frames
    api = pipe.tonemap(1, 1.5, 2.2)
    assert cli.shape == api.shape and (cli == api).all()
    assert (api == orc.tonemap(st.image, 1, 1.5, 2.2)).all()
    # -o *.exr: the same frames as linear radiance, bit for bit the oracle's rt_target
    from exr_reader import read_exr as _read_exr
    out_exr = str(tmp_path / "cli.exr")
    r = subprocess.run([exe, "-s", os.path.join(MEDIA, "cbox", "cbox.gltf"), "-o", out_exr, "--size", str(w), str(h), "--frames", str(frames),
                        "--spp", str(spp), "--depth", "4", "--integrator", "2", "--brdf", "1", "--seed", "9"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    _, ch = _read_exr(out_exr)
    lin = np.stack([ch[c] for c in "RGBA"], axis=-1)
    assert (lin.view(np.uint32) == pipe.read_image().view(np.uint32)).all()
    _compare("cli exr vs oracle", lin, st.image)
    pipe.close(), accel.close(), scene.close()


def test_frame_parallel_equals_sequential(gpurt, ctx):
    """gpurt_pipe_render_frame_mean + gpurt_pipe_accumulate_mean (frame-parallel multi-GPU sharding): frames rendered
    out of order into separate buffers and folded in order give the bits of the ordinary progressive loop"""
    import torch
    w, h, frames = 200, 120, 5
    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    prm = gpurt.pipe_params(max_frames=frames - 1, samples_per_frame=3, max_depth=4, integrator=2, brdf=1, seed=21)
    cam = gpurt.camera(0, w, h)
    seq = gpurt.RTPipe(scene, accel)
    n = 0
    while seq.render_frame(prm, cam, w, h) == 0:
        n += 1
    assert n == frames + 1                      # 0, 0, 1, ..., frames-1
    want = seq.read_image()
    par = gpurt.RTPipe(scene, accel)
    other = gpurt.RTPipe(scene, accel)          # stands in for another GPU
    means = torch.zeros((frames, h, w, 4), dtype=torch.float32, device="cuda")
    for f in (3, 0, 4, 2, 1):                   # any order, any pipe
        (par if f % 2 else other).render_frame_mean(prm, cam, w, h, f, means[f])
    for f in range(frames):
        par.accumulate_mean(means[f], f, w, h)
    torch.cuda.synchronize()
    got = par.read_image()
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert par.frame_index() == frames - 1
    with pytest.raises(gpurt.GpurtError):       # ReSTIR frames depend on the previous frame
        par.render_frame_mean(gpurt.pipe_params(integrator=3), cam, w, h, 0, means[0])
    for o in (seq, par, other, accel, scene):
        o.close()


def test_async_image_read_back_is_a_snapshot_of_its_frame(gpurt, ctx):
    """gpurt_pipe_read_image_async / _wait: every queued copy holds the image as of the frames rendered before the call,
    although the following frames were queued (and trace / shade) while it crossed PCIe"""
    import torch
    w, h, frames = 640, 360, 6
    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    prm = gpurt.pipe_params(max_frames=64, samples_per_frame=2, max_depth=4, integrator=2, brdf=1, seed=5)
    cam = gpurt.camera(0, w, h)
    seq = gpurt.RTPipe(scene, accel)
    want = []
    for _ in range(frames):
        assert seq.render_frame(prm, cam, w, h) == 0
        want.append(seq.read_image().copy())
    pipe = gpurt.RTPipe(scene, accel)
    pinned = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(frames)]
    for f in range(frames):                     # no host synchronisation inside this loop
        assert pipe.render_frame(prm, cam, w, h) == 0
        pipe.read_image_async(pinned[f])
    pipe.read_image_wait()
    for f in range(frames):
        assert (pinned[f].numpy().view(np.uint32) == want[f].view(np.uint32)).all(), f"snapshot of frame {f}"
    assert (pipe.read_image().view(np.uint32) == want[-1].view(np.uint32)).all()
    pipe.read_image_wait()                      # nothing pending: returns at once
    for o in (seq, pipe, accel, scene):
        o.close()


def test_cuda_frames_equal_the_reference_shader_digests(gpurt, orc, ctx):
    """the wavefront CUDA integrator against tests/golden/glsl_frames_golden.json — the digests of whole frames rendered
    by the REFERENCE'S OWN rt.rgen compiled as C++ (tests/test_oracle.py): every case, all five integrators, debug views.
    A new pipe renders frame 0 twice like the reference (rt.cpp:121-138); for integrators 0-2 the second call reproduces
    frame 0, for ReSTIR (3 / 4) the goldens hold that very call sequence under case/pipeK (make_glsl_golden.py).
    NaNs are compared as NaNs (buffer_digest), everything else bit for bit."""
    import importlib.util
    import json
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("make_glsl_golden", os.path.join(ROOT, "tests", "golden", "make_glsl_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "glsl_frames_golden.json")))
    checked, bad = 0, []
    for name, scene0, texs, w, h, frames, cam, kw in mg.frame_cases(gpurt):
        restir = kw.get("integrator", 0) in (3, 4)
        # the golden scenes were built without a context; rebuild the same scene on the device, object i = object i
        scene = gpurt.Scene(ctx).set_ordered()
        for t in texs:
            scene.add_texture(t)
        for i, d in enumerate(scene0.descs()):
            v, idx = scene0.object(i)
            m = gpurt.Material()
            m.albedo[:], m.emissive[:], m.metal_rough[:] = d.albedo[:3], d.emissive[:3], d.metal_rough[:2]
            m.albedo_tex, m.emissive_tex, m.metal_rough_tex, m.normal_tex = d.albedo_tex, d.emissive_tex, d.metal_rough_tex, d.normal_tex
            scene.add_object(v, idx, np.array(list(d.model), np.float32), m)
        assert [bytes(a)[:192] for a in scene.descs()] == [bytes(a)[:192] for a in scene0.descs()]
        accel = gpurt.Accel(scene)
        pipe = gpurt.RTPipe(scene, accel)
        prm = gpurt.pipe_params(**kw)
        cam = cam or gpurt.camera(0, w, h)
        for call in range(frames + 1):
            assert pipe.render_frame(prm, cam, w, h) == 0
            key = f"{name}/pipe{call}" if restir else f"{name}/frame{max(0, call - 1)}"
            if call == 0 and not restir:
                continue
            bufs = {"image": pipe.read_image(), "pos": pipe.read_gbuffer(0), "norm": pipe.read_gbuffer(1), "albedo": pipe.read_gbuffer(2)}
            if restir:
                bufs["reservoirs"] = pipe.read_reservoirs()
            for b, arr in bufs.items():
                if mg.buffer_digest(arr) != want[f"{key}/{b}"]:
                    bad.append(f"{key}/{b}")
                checked += 1
            c = pipe.ray_counts()
            if f"{c[0]},{c[1]}" != want[f"{key}/rays"]:
                bad.append(f"{key}/rays")
        pipe.close(), accel.close(), scene.close()
    assert not bad, f"{len(bad)} of {checked} buffers differ from the reference shader's: {bad[:12]}"
    assert checked > 250

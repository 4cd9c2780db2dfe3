#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native GPU-RT hot path.

Metric (BASELINE.json): Mrays/s closest-hit on Sponza (plus M closest-point queries/s in `cpq`).
Workload (BASELINE.json configs[1]): Sponza 1080p, integrator 1 (Material),
GGX, max_depth 2, 1 spp, frame 0 -> the frame's closest-hit ray set: 2,073,600 primary rays plus the
surviving 1-bounce rays.  Sponza.bin is missing from the reference snapshot, so unless
GPURT_SPONZA_GLTF points at a complete copy the scene is the labelled procedural stand-in with the
same triangle count / object count / extent (`config.scene`).

One "step" = one pass of the hot path (closest-hit traversal) over the whole ray set.
  value  : rays / device time, ray buffers resident in HBM, L2 flushed between steps
  e2e    : same rays through the C ABI with HOST (pinned) buffers: H2D rays + kernel + D2H hits
  roofline: trace kernel alone, algorithmic bytes per SURVEY §8d (48 B stream + 80 B x nodes visited
            + 48 B x triangles tested per ray, counters from the instrumented kernel)
  cpu_baseline: the oracle's CPU BVH traversal (restated reference semantics) on a bounded sample
`--impl reference` times that CPU path alone (the reference has no host implementation of this
path: it runs inside the Vulkan driver — see DESIGN.md §6).

Launch: python bench.py --gpus N --steps K --warmup W   (N>1 under torchrun, one rank per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))

import numpy as np  # noqa: E402

W, H = 1920, 1080
CAM_POS, CAM_AT, VFOV = (-1000.0, 200.0, 0.0), (0.0, 200.0, 0.0), 90.0
BYTES_NODE, BYTES_TRI, BYTES_STREAM = 80, 48, 48
CPQ_BYTES_STREAM = 48


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc, self.t_rows = index, [], None, []
        self.window = [0.0, 1e30]

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.t_rows.append((time.time(), [x.strip() for x in line.strip().split(",")]))
        except Exception:
            pass

    def summary(self):
        if self.proc:
            self.proc.terminate()
        time.sleep(0.05)
        self.rows = [r for t, r in self.t_rows if self.window[0] - 0.05 <= t <= self.window[1] + 0.05]
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def tea(v0, v1):
    """rtcommon.glsl:99-109, vectorised (bench-side generator, independent of the oracle)"""
    v0 = np.asarray(v0, np.uint32).copy()
    v1 = np.asarray(v1, np.uint32).copy() if np.ndim(v1) else np.full_like(v0, v1)
    s0 = np.uint32(0)
    with np.errstate(over="ignore"):
        for _ in range(16):
            s0 = np.uint32(s0 + np.uint32(0x9E3779B9))
            v0 += ((v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ (v1 + s0) ^ ((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))
            v1 += ((v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ (v0 + s0) ^ ((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))
    return v0


def lcg_randf(state):
    with np.errstate(over="ignore"):
        state *= np.uint32(1664525)
        state += np.uint32(1013904223)
    return (state & np.uint32(0x00FFFFFF)).astype(np.float32) / np.float32(0x01000000)


def primary_rays(gpurt, rank):
    """make_camera_ray (rt.rgen:551-565) at frame 0 (jitter 0.5), camera shifted per rank (weak scaling)."""
    pos = (CAM_POS[0] + 40.0 * rank, CAM_POS[1], CAM_POS[2] + 25.0 * rank)
    cam = gpurt.camera(1, W, H, pos, CAM_AT, VFOV)
    iP = np.array(cam.iP, np.float32).reshape(4, 4).T  # column-major -> math matrix
    iV = np.array(cam.iV, np.float32).reshape(4, 4).T
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
    u = (xs + 0.5) / W * 2.0 - 1.0
    v = (ys + 0.5) / H * 2.0 - 1.0
    ndc = np.stack([u, v, np.zeros_like(u), np.ones_like(u)], -1).reshape(-1, 4)
    target = ndc @ iP.T
    d = np.concatenate([target[:, :3], np.zeros((target.shape[0], 1), np.float32)], 1) @ iV.T
    d = d[:, :3] / np.linalg.norm(d[:, :3], axis=1, keepdims=True)
    o = (iV @ np.array([0, 0, 0, 1], np.float32))[:3]
    rays = np.zeros((W * H, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, 1e-5, d, 1e7
    return rays, cam


def build_scene(gpurt, ctx):
    scene = gpurt.Scene(ctx)
    path = os.environ.get("GPURT_SPONZA_GLTF")
    if path and os.path.exists(path):
        scene.load(path)
        return scene, "sponza (GPURT_SPONZA_GLTF)"
    scene.make_sponza_standin()
    return scene, "sponza_standin (Sponza.bin missing from the reference snapshot)"


def scene_world_tris(scene, orc):
    """world-space triangles for the CPU arm (the oracle's own flattening, contract N1)"""
    return np.concatenate([orc.flatten(*scene.object(i), np.array(d.model, np.float32))
                           for i, d in enumerate(scene.descs())])


def reference_arm(args, rank, world):
    """--impl reference: CPU traversal (oracle restatement of the reference semantics) on all host
    cores, on a bounded sample of the same workload: the config-2 frame is rendered once on the CPU
    (untimed) with the oracle's restatement of rt.rgen, the closest-hit rays it traces (primary + bounce)
    are recorded, and every step traces a 1-in-8 sample of them through the CPU BVH."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gpurt
    import orc
    scene, label = build_scene(gpurt, None)
    t0 = time.time()
    rs = orc.RenderScene(scene)            # flattens the scene and builds the CPU LBVH
    build_s = time.time() - t0
    cam = gpurt.camera(1, W, H, CAM_POS, CAM_AT, VFOV)
    c = gpurt.Constants()
    c.clear_col[:] = [0.3, 0.3, 0.3, 1.0]
    c.env_light[:] = [1.0, 1.0, 1.0, 1.0]
    c.frame, c.samples, c.max_frame, c.max_depth, c.integrator, c.brdf, c.use_rr = 0, 1, 1, 2, 1, 1, 0
    c.use_temporal, c.n_lights, c.n_objs = 1, scene.counts()["lights"], scene.counts()["objs"]
    rays = orc.frame_rays(rs, W, H, np.frombuffer(bytes(c), np.uint32), np.frombuffer(bytes(cam), np.uint32), 0, 2 * W * H)
    stride = 8
    sample = rays[::stride].copy()
    cores = orc.lib.orc_hw_threads()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        rs.bvh.closest_hit(sample, threads=cores)
        if i >= args.warmup:
            times.append(time.time() - t0)
    dt = float(np.mean(times))
    val = sample.shape[0] / dt / 1e6
    what = f"every {stride}th of the {rays.shape[0]} closest-hit rays of the frame ({sample.shape[0]} rays per step)"
    line = {"impl": "reference", "metric": "Mrays/s closest-hit on Sponza", "value": val, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "sponza 1080p config-2 ray set (2,073,600 primary + 1-bounce rays), closest-hit",
                       "scene": label, "sample": what},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": what + f"; CPU LBVH build {build_s:.2f} s not included"},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    import gpurt

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gpurt.Context(local)
    scene, label = build_scene(gpurt, ctx)
    gpurt.Accel(scene).close()      # first build warms the allocation pool
    accel = gpurt.Accel(scene)
    info = accel.info()
    # pose edit + in-place rebuild (GPURT::build_accel after edit_scene): same matrix, so the scene is unchanged
    scene.set_transform(0, np.array(list(scene.descs()[0].model), np.float32))
    accel.update()
    update_ms = accel.info().build_ms

    # ---- the frame's ray set, resident in HBM ------------------------------------------------
    # Rendered by the wavefront integrator itself (config 2: integrator 1 = Material, GGX, depth 2,
    # 1 spp, frame 0, no RR, env light on); the closest-hit rays it traced are re-used as the step.
    pos = (CAM_POS[0] + 40.0 * rank, CAM_POS[1], CAM_POS[2] + 25.0 * rank)   # weak scaling: one view per rank
    cam = gpurt.camera(1, W, H, pos, CAM_AT, VFOV)
    pipe = gpurt.RTPipe(scene, accel)
    prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=2, samples_per_frame=1, max_frames=1, use_rr=0,
                            env_scale=1.0, seed=rank)
    ctx.use_torch_stream()
    frame_ms = []
    for _ in range(4):
        pipe.reset_frame()
        assert pipe.render_frame(prm, cam, W, H) == 0
        frame_ms.append(pipe.time_ms())
    frame_rays = pipe.ray_counts()
    d_rays = torch.cat([pipe.bounce_rays(0), pipe.bounce_rays(1)]).clone()
    rays_np = d_rays.cpu().numpy()
    prim = rays_np[: W * H]
    n_rays = rays_np.shape[0]
    d_hits = torch.empty((n_rays, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step():
        accel.trace_closest(d_rays, d_hits)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.time()
    sampler.window[0] = wall0
    for a, b in ev:
        flush.zero_()          # L2 flush between timed iterations (not timed)
        a.record()
        step()
        b.record()
    barrier()
    wall = time.time() - wall0
    sampler.window[1] = wall0 + wall
    ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.summary()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    tot = torch.tensor([float(n_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_rays = float(tot.item())
    value = total_rays * args.steps / (ms_max * 1e-3) / 1e6

    # ---- roofline of the trace kernel (rank 0's numbers) ---------------------------------------
    st = accel.trace_stats(d_rays, d_hits)
    nodes_per_ray = st.nodes_visited / st.rays
    tris_per_ray = st.tris_tested / st.rays
    bytes_per_ray = BYTES_STREAM + BYTES_NODE * nodes_per_ray + BYTES_TRI * tris_per_ray
    kernel_ms = ms / args.steps
    achieved = bytes_per_ray * n_rays / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = peaks()

    # primary-only pass (SURVEY §8d headline) and closest-point queries on the hit points
    n_p = prim.shape[0]
    flush.zero_()
    accel.trace_closest(d_rays[:n_p], d_hits[:n_p])
    prim_ms = []
    for _ in range(5):
        flush.zero_()
        accel.trace_closest(d_rays[:n_p], d_hits[:n_p])
        prim_ms.append(ctx.last_kernel_ms())
    q_np = np.zeros((n_p, 4), np.float32)
    hp = d_hits[:n_p].cpu().numpy().view(gpurt.HIT_DT).reshape(-1)
    tt = np.where(np.isfinite(hp["t"]), hp["t"], 100.0).astype(np.float32)
    jit = (lcg_randf(tea(np.arange(n_p, dtype=np.uint32), np.uint32(0xD00D + rank)))[:, None] - 0.5) * 60.0
    q_np[:, :3] = prim[:, 0:3] + 0.8 * tt[:, None] * prim[:, 4:7] + jit
    q_np[:, 3] = np.inf
    d_q = torch.from_numpy(q_np).cuda()
    d_cp = accel.closest_points(d_q)
    cpq_ms = []
    for _ in range(5):
        flush.zero_()
        accel.closest_points(d_q, d_cp)
        cpq_ms.append(ctx.last_kernel_ms())

    # ---- e2e: C ABI with HOST buffers (pinned), copies inside the timed region ------------------
    h_rays = torch.from_numpy(rays_np).pin_memory()
    h_hits = torch.empty((n_rays, 4), dtype=torch.float32).pin_memory()
    hr, hh = h_rays.numpy(), h_hits.numpy()
    for _ in range(2):
        accel.trace_closest(hr, hh)
    barrier()
    t0 = time.time()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        accel.trace_closest(hr, hh)     # H2D + kernel + D2H + sync inside the call
    barrier()
    e2e_s = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = total_rays * e2e_steps / float(e2e_s.item()) / 1e6
    assert np.array_equal(hh.view(np.uint32), d_hits.cpu().numpy().view(np.uint32)), "host and device paths disagree"

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) -------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc
        ob = orc.Bvh(scene_world_tris(scene, orc))
        cores = orc.lib.orc_hw_threads()
        stride = 1   # the whole step: ~13 core-seconds of CPU traversal
        sample = rays_np[::stride].copy()
        t0 = time.time()
        ref = ob.closest_hit(sample, threads=cores)
        dt = time.time() - t0
        got = np.ascontiguousarray(d_hits.cpu().numpy()[::stride])
        assert np.array_equal(ref.view(np.uint32).reshape(-1, 4), got.view(np.uint32)), "GPU result differs from the CPU oracle"
        cpu = {"value": sample.shape[0] / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
               "sample": f"all {sample.shape[0]} rays of one step ({dt:.1f} s wall on {cores} threads); "
                         "results compared bit-exactly with the GPU's"}

    # physical DRAM bytes of one launch: dram__bytes_read.sum + dram__bytes_write.sum of k_trace_closest from the
    # committed ncu capture of this same workload, scaled per ray (never measured under the profiler here)
    traffic, traffic_src, secondary = None, None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["dram_bytes_per_ray"] * n_rays, tj["source"]
        secondary = tj.get("secondary_ceilings")   # L2 / issue-slot ceilings of the same ncu capture
    if rank == 0:
        line = {
            "metric": "Mrays/s closest-hit on Sponza", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "sponza 1080p config-2 ray set (2,073,600 primary + 1-bounce rays), closest-hit",
                       "scene": label, "rays_per_gpu": n_rays, "tris": info.n_tris, "wide_nodes": info.n_wide_nodes,
                       "wide_depth": info.wide_depth, "bvh_build_ms": info.build_ms, "bvh_update_ms": update_ms,
                       "bvh_build_mtris_s": info.n_tris / (info.build_ms * 1e-3) / 1e6,
                       "l2": "flushed between timed steps (256 MiB memset)", "sharding": "rays sharded per rank, scene replicated, no collective in the timed region"},
            "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": int(n_rays * 32), "d2h_bytes_per_step": int(n_rays * 16)},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "k_trace_closest<false>",
                         "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                         "stream_floor_bytes_per_ray": BYTES_STREAM, "kernel_ms": kernel_ms,
                         "secondary_ceilings": secondary},
            "cpu_baseline": cpu,
            "primary_only": {"value": n_p / (float(np.median(prim_ms)) * 1e-3) / 1e6, "unit": "Mrays/s", "rays": n_p},
            "cpq": {"value": n_p / (float(np.median(cpq_ms)) * 1e-3) / 1e6, "unit": "Mqueries/s", "queries": n_p,
                    "what": "closest-point queries near the primary hit points"},
            "render": {"what": "whole config-2 frame through gpurt_pipe_render_frame (gen + trace + shade + accumulate)",
                       "ms_per_frame": float(np.median(frame_ms)), "closest_rays": frame_rays[0], "any_rays": frame_rays[1],
                       "mrays_s": frame_rays[0] / (float(np.median(frame_ms)) * 1e-3) / 1e6,
                       "mpaths_s": W * H / (float(np.median(frame_ms)) * 1e-3) / 1e6},
            "clocks": clocks, "wall_s": wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

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

At N > 1 (one rank per GPU, scene replicated, one view per rank = weak scaling) the timed step includes result
placement: every rank's trace kernel stores its hits straight into rank 0's result array over NVLink
(gpurt_shared_alloc / gpurt_shared_open), so a step ends with all N result sets in rank 0's HBM.  The line also carries
`placement` (the same step with results left local, and an NCCL gather of the same bytes, for comparison) and `strong`
(BASELINE configs 4 and 5 at the same N: fixed total work sharded over the ranks, results placed on rank 0).
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


def build_irregular_scene(gpurt, ctx, copies=16):
    """A Sponza-sized scene of REAL meshes (the stand-in is a regular procedural grid, on which a Morton order is as good as
    any tree): `copies` instances of every object of media/cbox (16,732 triangles: the bunny / dragon-class meshes the
    reference ships; the side walls, the ceiling and the back wall are left out so that the instances see each other) under
    rotations, scales and offsets drawn from the reference RNG -> 96 objects, 266,944 triangles."""
    src = gpurt.Scene(None).load(os.path.join(ROOT, "tests", "data", "media", "cbox", "cbox.gltf"))
    descs = src.descs()
    keep = [i for i in range(len(descs)) if i not in (5, 6, 7, 9)]
    objs = [src.object(i) for i in keep]
    descs = [descs[i] for i in keep]
    scene = gpurt.Scene(ctx).set_ordered()
    side = int(np.ceil(np.sqrt(copies)))
    for k in range(copies):
        st = tea(np.array([k], np.uint32), np.uint32(0x5EED))
        r = [float(lcg_randf(st)[0]) for _ in range(6)]
        ang, tilt, sc = 2 * np.pi * r[0], 0.5 * (r[1] - 0.5), 0.7 + 0.6 * r[2]
        cy, sy, cx, sx = np.cos(ang), np.sin(ang), np.cos(tilt), np.sin(tilt)
        R = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        T = np.eye(4)
        T[:3, :3] = R * sc
        T[:3, 3] = [13.0 * (k % side) + 3.0 * (r[3] - 0.5), 4.0 * (r[4] - 0.5), 13.0 * (k // side) + 3.0 * (r[5] - 0.5)]
        for (v, idx), d in zip(objs, descs):
            m = gpurt.Material()
            m.albedo[:], m.emissive[:], m.metal_rough[:] = d.albedo[:3], d.emissive[:3], d.metal_rough[:2]
            m.albedo_tex = m.emissive_tex = m.metal_rough_tex = m.normal_tex = -1
            model = T @ np.array(list(d.model), np.float64).reshape(4, 4).T
            scene.add_object(v, idx, np.ascontiguousarray(model.T, np.float32).reshape(16), m)
    ext = 13.0 * (side - 1)
    cam = dict(pos=(-9.0, 9.0, -9.0), at=(ext * 0.5, -1.0, ext * 0.5), vfov=70.0)
    return scene, f"cbox_instances ({copies} x media/cbox under random transforms)", cam


def tree_report(gpurt, torch, ctx, scene, cam_args, flush):
    """closest-hit / closest-point throughput of one scene under the default (binned-SAH) build and the Morton LBVH:
    the frame's config-2 ray set (primary + 1 bounce at 1080p) is recorded once and traced through both trees"""
    out = {}
    cam = gpurt.camera(1, W, H, cam_args["pos"], cam_args["at"], cam_args["vfov"])
    prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=2, samples_per_frame=1, max_frames=1, use_rr=0, env_scale=1.0, seed=0)
    rays = None
    for name, flags in (("default_sah", 0), ("lbvh", gpurt.BUILD_LBVH)):
        gpurt.Accel(scene, flags).close()
        accel = gpurt.Accel(scene, flags)
        if rays is None:
            pipe = gpurt.RTPipe(scene, accel)
            ctx.use_torch_stream()
            pipe.render_frame(prm, cam, W, H)
            rays = torch.cat([pipe.bounce_rays(0), pipe.bounce_rays(1)]).clone()
            pipe.close()
        n, n_p = rays.shape[0], W * H
        hits = torch.empty((n, 4), dtype=torch.float32, device=rays.device)

        def med(fn, reps=7):
            ms = []
            flush.zero_()
            fn()
            for _ in range(reps):
                flush.zero_()
                fn()
                ms.append(ctx.last_kernel_ms())
            return float(np.median(ms))
        all_ms = med(lambda: accel.trace_closest(rays, hits))
        prim_ms = med(lambda: accel.trace_closest(rays[:n_p], hits[:n_p]))
        st = accel.trace_stats(rays, hits)
        t = hits[:n_p, 0]
        q = torch.empty((n_p, 4), dtype=torch.float32, device=rays.device)
        q[:, :3] = rays[:n_p, 0:3] + (0.8 * torch.where(torch.isfinite(t), t, torch.full_like(t, 10.0)))[:, None] * rays[:n_p, 4:7]
        q[:, 3] = float("inf")
        cp = accel.closest_points(q)
        cpq_ms = med(lambda: accel.closest_points(q, cp))
        info = accel.info()
        out[name] = {"mrays_s": n / (all_ms * 1e-3) / 1e6, "primary_mrays_s": n_p / (prim_ms * 1e-3) / 1e6,
                     "bounce_mrays_s": (n - n_p) / max(1e-9, (all_ms - prim_ms) * 1e-3) / 1e6 if n > n_p else None,
                     "cpq_mqueries_s": n_p / (cpq_ms * 1e-3) / 1e6, "rays": n, "nodes_per_ray": st.nodes_visited / st.rays,
                     "tris_per_ray": st.tris_tested / st.rays, "build_ms": info.build_ms, "wide_nodes": info.n_wide_nodes,
                     "binary_tree_sah_cost": info.tree_cost}
        accel.close()
    out["tris"] = scene.counts()["tris"]
    return out


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
    # the CPU arm never maps the CUDA library: the scene is constructed by the host half alone (libgpurt_host.so = glTF
    # loader / stand-in generator / camera, g++ only), everything timed is oracle/
    os.environ["GPURT_LIB"] = os.path.join(ROOT, "gpu-rt_b200", "libgpurt_host.so")
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
    cores = orc.lib.orc_hw_threads()
    # every ray of the step (the same config as the GPU arm) unless the whole run would then pass ~2 minutes: a probe on
    # 1/16 of the rays sizes the sample
    probe = rays[::16].copy()
    t0 = time.time()
    rs.bvh.closest_hit(probe, threads=cores)
    est_full = (time.time() - t0) * 16.0
    stride = max(1, int(np.ceil(est_full * (args.warmup + args.steps) / 120.0)))
    sample = rays[::stride].copy()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        rs.bvh.closest_hit(sample, threads=cores)
        if i >= args.warmup:
            times.append(time.time() - t0)
    dt = float(np.mean(times))
    val = sample.shape[0] / dt / 1e6
    what = (f"all {rays.shape[0]} closest-hit rays of the frame per step" if stride == 1 else
            f"every {stride}th of the {rays.shape[0]} closest-hit rays of the frame ({sample.shape[0]} rays per step)")
    line = {"impl": "reference", "metric": "Mrays/s closest-hit on Sponza", "value": val, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "sponza 1080p config-2 ray set (2,073,600 primary + 1-bounce rays), closest-hit",
                       "scene": label, "rays_per_gpu": int(sample.shape[0]), "sample": what},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": what + f"; CPU LBVH build {build_s:.2f} s not included"},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa(local):
    """Pin this rank's host threads (and, by first touch, its pinned staging memory) to the NUMA node the GPU hangs
    off, when the box exposes one (SCALE_r01: e2e scaled 2.4x on 8 GPUs with every rank's staging memory wherever the
    scheduler put it).  Returns what was found, for the JSON line."""
    info = {"numa_node": None, "cpus": len(os.sched_getaffinity(0)), "bound": False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{dev}"
        node = int(open(f"{base}/numa_node").read())
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        info["numa_node"] = node
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus and len(cpus) < info["cpus"]:
            os.sched_setaffinity(0, cpus)
            info["cpus"], info["bound"] = len(cpus), True
    except Exception as e:  # noqa: BLE001 — diagnostics only
        info["error"] = str(e)[:80]
    return info


def device_max(dist, world, x, dev, op="max"):
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def roofline_obj(kernel, n, ms, nodes, tris, stream_bytes, peak, peak_src, ncu=None):
    """logical bytes = stream + 80 B x nodes visited + 48 B x triangles tested per element, over the
    kernel's device time; `ncu` adds the physical side of the same kernel from the committed capture"""
    bpe = stream_bytes + BYTES_NODE * nodes + BYTES_TRI * tris
    achieved = bpe * n / (ms * 1e-3) / 1e9
    o = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
         "peak_source": peak_src, "kernel": kernel, "elements": int(n), "kernel_ms": ms, "bytes_per_element": bpe,
         "nodes_per_element": nodes, "tris_per_element": tris, "stream_floor_bytes_per_element": stream_bytes,
         "m_elements_s": n / (ms * 1e-3) / 1e6,
         "frac_is": "LOGICAL bytes (the BVH is L2-resident): read dram_frac and issue_lane_frac for what the hardware sees"}
    if ncu:
        o["traffic"] = ncu["dram_bytes_per_element"] * n
        o["traffic_source"] = ncu["source"]
        o["dram_frac"] = o["traffic"] / (ms * 1e-3) / 1e9 / peak
        sc = ncu.get("secondary_ceilings") or {}
        if "issue_slots_active_pct" in sc and "lanes_per_instruction" in sc:
            o["issue_lane_frac"] = sc["issue_slots_active_pct"] / 100.0 * sc["lanes_per_instruction"] / 32.0
        o["secondary_ceilings"] = sc
    return o


def ncu_facts():
    """physical-side numbers of the committed ncu captures (never measured under the profiler here), newest round first"""
    for name in ("r05_traffic.json", "r04_traffic.json", "r03_traffic.json", "r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            tj = json.load(open(p))
            if "kernels" not in tj:  # round-1 layout: the closest-hit kernel only
                tj = {"kernels": {"closest": {"dram_bytes_per_element": tj["dram_bytes_per_ray"], "source": tj["source"],
                                              "secondary_ceilings": tj.get("secondary_ceilings")}}}
            return tj["kernels"]
    return {}


# ---- strong-scaling sub-benchmarks (BASELINE configs 4 and 5 at the same N) -----------------------------------------
def strong_config4(gpurt, torch, dist, ctx, rank, world, dev, n_tris, n_queries, check, schedule=None):
    """config 4: 10 M-triangle soup, 100 M closest-point queries in contiguous ranges per rank, all 3.2 GB of results
    placed in rank 0's memory inside the timed region, through gpurt_gather_* (csrc/gather.cu): every rank makes ONE
    gpurt_closest_points call on its range with its part of rank 0's array as the result pointer.  The library sorts the
    range once, traverses it in slices of the processing order, and after each slice the copy engine moves the slice's
    results + storage indices into an inbox on rank 0 and raises a flag; kernels rank 0 queued on a side stream wait for
    the flags and scatter the slices into the array while rank 0 traverses its own range.  No collective, no host
    synchronisation inside the round.  GPURT_CONFIG4_MODE=chunks runs round 2's caller-side chunked copies instead."""
    if os.environ.get("GPURT_CONFIG4_MODE", "gather") == "chunks" or schedule is not None:
        return strong_config4_chunks(gpurt, torch, dist, ctx, rank, world, dev, n_tris, n_queries, check, schedule)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from config4_cpq import make_queries, make_soup
    from gpurt.dist import shard_range
    tris_h = make_soup(n_tris, dev).cpu().numpy()
    scene = gpurt.Scene(ctx)
    scene.add_triangles(tris_h)
    accel = gpurt.Accel(scene)
    info = accel.info()
    # rank 0 also receives and scatters everybody else's results (~6 GB of memory traffic next to its own traversal): it gets
    # `share` of an equal part of the queries, the rest is spread over the other ranks (still contiguous ascending ranges).
    # Measured at 8 GPUs (gpurun_out/r03o): share 1.0 -> 7921 Mq/s, 0.85 -> 8734, 0.7 -> 9489.
    # With k = cost of scattering one foreign record in units of one query (0.049, fitted at 8 GPUs) the two finish together
    # when rank 0 answers s0 = (1/(N-1) - k) / (1 - k + 1/(N-1)) of the queries: 0.975 / 0.886 / 0.69 of an equal part at 2 / 4 / 8.
    k_scatter = 0.049
    model = world * (1.0 / (world - 1) - k_scatter) / (1.0 - k_scatter + 1.0 / (world - 1)) if world > 1 else 1.0
    share = float(os.environ.get("GPURT_CONFIG4_OWNER_SHARE", model)) if world > 1 else 1.0
    n0 = int(n_queries / world * share) // 128 * 128
    first = [0] + [n0 + shard_range(n_queries - n0, r, world - 1)[0] for r in range(max(0, world - 1))] + [n_queries]
    if world == 1:
        first = [0, n_queries]
    a, b = first[rank], first[rank + 1]
    nq = b - a
    q = torch.empty((nq, 4), dtype=torch.float32, device=dev)
    for c0 in range(0, nq, 12_500_000):
        c1 = min(nq, c0 + 12_500_000)
        q[c0:c1] = make_queries(a + c0, a + c1, dev)
    box = [None]
    if rank == 0:
        g = gpurt.Gather.create(ctx, n_queries, 32, first)
        box[0] = g.handle
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    if rank != 0:
        g = gpurt.Gather.open(ctx, n_queries, 32, first, rank, handle=box[0])
    if world > 1:
        dist.barrier()
    local = torch.empty((nq, 8), dtype=torch.float32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run():
        e0.record()
        if rank == 0:
            g.begin()
        accel.closest_points(q, g.mine())
        if rank == 0:
            g.end()
        e1.record()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    run()                                                          # warm-up (arenas, sort scratch, streams)
    sync()
    t0 = time.time()
    run()
    sync()
    wall = time.time() - t0
    ms = device_max(dist, world, e0.elapsed_time(e1), dev)
    timeouts = g.end(sync=True) if rank == 0 else 0
    # the same queries with the results left on each rank (no placement): what the kernels alone take
    e0.record()
    accel.closest_points(q, local)
    e1.record()
    sync()
    ms_local = device_max(dist, world, e0.elapsed_time(e1), dev)
    # throughput depends on how dense a batch is (the batch is traversed in Morton order of its points: the more points per
    # triangle, the more the lanes of a warp share): the same range in calls of 12.5 M points — what one rank of 8 gets
    ms_small = None
    if world == 1 and nq > 12_500_000:
        e0.record()
        for c0 in range(0, nq, 12_500_000):
            c1 = min(nq, c0 + 12_500_000)
            accel.closest_points(q[c0:c1], local[c0:c1])
        e1.record()
        sync()
        ms_small = device_max(dist, world, e0.elapsed_time(e1), dev)
    out = {"config": "4: synthetic 10 M-triangle soup, 100 M closest-point queries", "tris": info.n_tris, "queries": n_queries,
           "n_gpus": world, "mqueries_s": n_queries / (ms * 1e-3) / 1e6, "ms": ms, "wall_ms_barrier_to_barrier": wall * 1e3,
           "mqueries_s_in_calls_of_12_5M_results_left_local": (n_queries / (ms_small * 1e-3) / 1e6) if ms_small else None,
           "results": "all results in rank 0's array at the end of the timed region (gpurt_gather_*): one call per rank, slices of the "
                      "sorted batch copied into rank 0's inbox while the next slice is traversed, scattered there by rank 0's side "
                      "stream while rank 0 traverses its own range; no collective",
           "bytes_into_rank0": int((n_queries - first[1]) * 36), "queries_per_call": nq, "owner_share_of_equal_part": share, "flag_wait_timeouts": timeouts,
           "mqueries_s_results_left_local": n_queries / (ms_local * 1e-3) / 1e6, "bvh_build_ms": info.build_ms}
    if rank == 0 and check:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc
        remote = g.tensor().view(torch.float32).view(-1, 8)
        c0 = first[world - 1]                                      # a slice the LAST rank placed
        got = remote[c0:c0 + check].cpu().numpy().view(np.uint32)
        ref = orc.Bvh(tris_h).closest_point(make_queries(c0, c0 + check, dev).cpu().numpy()).view(np.uint32).reshape(-1, 8)
        out["check"] = {"queries": check, "placed_by_rank": world - 1,
                        "bit_exact_vs_oracle": bool((got[:, [0, 1, 2, 3, 4, 6, 7]] == ref[:, [0, 1, 2, 3, 4, 6, 7]]).all())}
    if world > 1:
        dist.barrier()
    g.close()
    accel.close(), scene.close()
    return out


def strong_config4_chunks(gpurt, torch, dist, ctx, rank, world, dev, n_tris, n_queries, check, schedule=None):
    """config 4: 10 M-triangle soup, 100 M closest-point queries in contiguous ranges per rank, all 3.2 GB of results
    placed in rank 0's memory inside the timed region: rank 0's kernel writes there directly; the other ranks compute
    chunk k into a local buffer while chunk k-1 crosses NVLink on a second stream (copy engine, no collective).
    (Measured alternative, profiles/r03_summary.md: one call per rank with rank 0's buffer as the result pointer and
    GPURT_PLACE_SLICES=8 — the library sorts the 12.5 M batch once and stores slice k of the processing order to rank 0 while
    slice k + 1 is traversed.  Kernels run at the single-GPU rate (7.95x at 8 GPUs with results left local) but 87.5 M
    scattered 32-byte stores from 7 senders into one GPU arrive at only 175 GB/s: 6039 Mq/s against 8105 for the chunked copies.)"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from config4_cpq import make_queries, make_soup
    from gpurt.dist import shard_range, shared_result_buffer
    tris_h = make_soup(n_tris, dev).cpu().numpy()
    scene = gpurt.Scene(ctx)
    scene.add_triangles(tris_h)
    accel = gpurt.Accel(scene)
    info = accel.info()
    a, b = shard_range(n_queries, rank, world)
    nq = b - a
    # Chunks by storage position.  Each chunk is sorted and traversed on its own, and small chunks are less coherent
    # (3.1 M points: 1096 Mq/s per GPU, 12.5 M: 1328), but only the last chunk's copy is exposed: so a large first chunk
    # and a short tail — fractions of a 12.5 M block, overridable with GPURT_CONFIG4_SCHEDULE="0.75,0.17,0.08".
    if schedule is None:
        schedule = [float(x) for x in os.environ.get("GPURT_CONFIG4_SCHEDULE", "0.75,0.17,0.08").split(",")]
    bounds = [0]
    for blk0 in range(0, nq, 12_500_000):
        blk = min(12_500_000, nq - blk0)
        if rank == 0 or world == 1 or blk < (4 << 20):
            bounds.append(blk0 + blk)                               # rank 0 writes in place: nothing to overlap
            continue
        acc = 0.0
        for f in schedule[:-1]:
            acc += f
            bounds.append(blk0 + min(blk, int(blk * acc) // 128 * 128))
        bounds.append(blk0 + blk)
    chunks = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i]]
    chunk = max(c1 - c0 for c0, c1 in chunks)
    q = torch.empty((nq, 4), dtype=torch.float32, device=dev)
    for c0 in range(0, nq, 12_500_000):
        c1 = min(nq, c0 + 12_500_000)
        q[c0:c1] = make_queries(a + c0, a + c1, dev)
    shared = shared_result_buffer(ctx, n_queries * 32)
    remote = shared.tensor().view(torch.float32).view(-1, 8)       # rank 0: its own memory; others: NVLink mapping
    local = [torch.empty((chunk, 8), dtype=torch.float32, device=dev) for _ in range(2)] if rank else None
    main_s, copy_s = torch.cuda.current_stream(), torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    free_ev = [None, None]

    def run():
        e0.record()
        for k, (c0, c1) in enumerate(chunks):
            if rank == 0:
                accel.closest_points(q[c0:c1], shared.at((a + c0) * 32))
                continue
            buf = local[k & 1]
            if free_ev[k & 1] is not None:
                main_s.wait_event(free_ev[k & 1])                  # the copy that last read this buffer is done
            accel.closest_points(q[c0:c1], buf[: c1 - c0])
            done = torch.cuda.Event()
            done.record(main_s)
            copy_s.wait_event(done)
            with torch.cuda.stream(copy_s):
                remote[a + c0:a + c1].copy_(buf[: c1 - c0], non_blocking=True)
                free_ev[k & 1] = torch.cuda.Event()
                free_ev[k & 1].record(copy_s)
        main_s.wait_stream(copy_s)
        e1.record()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    run()                                                          # warm-up (communicators, arenas, sort scratch)
    sync()
    t0 = time.time()
    run()
    sync()
    wall = time.time() - t0
    ms = device_max(dist, world, e0.elapsed_time(e1), dev)
    # the same queries with the results left on each rank (no placement): what the kernels alone take
    e0.record()
    for c0, c1 in chunks:
        accel.closest_points(q[c0:c1], local[0][: c1 - c0] if rank else shared.at((a + c0) * 32))
    e1.record()
    sync()
    ms_local = device_max(dist, world, e0.elapsed_time(e1), dev)
    out = {"config": "4: synthetic 10 M-triangle soup, 100 M closest-point queries", "tris": info.n_tris, "queries": n_queries,
           "n_gpus": world, "mqueries_s": n_queries / (ms * 1e-3) / 1e6, "ms": ms, "wall_ms_barrier_to_barrier": wall * 1e3,
           "results": "all results in rank 0's memory at the end of the timed region (rank 0: direct; others: chunk k computed "
                      "while chunk k-1 is copied over NVLink, no collective)",
           "bytes_into_rank0": int((n_queries - (shard_range(n_queries, 0, world)[1])) * 32), "chunk_queries": [c1 - c0 for c0, c1 in chunks] if len(chunks) <= 16 else chunk,
           "mqueries_s_results_left_local": n_queries / (ms_local * 1e-3) / 1e6, "bvh_build_ms": info.build_ms}
    if rank == 0 and check:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc
        c0 = shard_range(n_queries, world - 1, world)[0]           # a slice the LAST rank placed
        got = remote[c0:c0 + check].cpu().numpy().view(np.uint32)
        ref = orc.Bvh(tris_h).closest_point(make_queries(c0, c0 + check, dev).cpu().numpy()).view(np.uint32).reshape(-1, 8)
        out["check"] = {"queries": check, "placed_by_rank": world - 1,
                        "bit_exact_vs_oracle": bool((got[:, [0, 1, 2, 3, 4, 6, 7]] == ref[:, [0, 1, 2, 3, 4, 6, 7]]).all())}
    if world > 1:
        dist.barrier()
    shared.close()
    accel.close(), scene.close()
    return out


def strong_config3_restir(gpurt, torch, dist, ctx, rank, world, dev):
    """config 3's ReSTIR frames (mis_test 1920x1080, integrator 3, depth 4, 1 spp, res_samples 4, temporal reuse) with the frame
    sharded over the ranks in contiguous row bands: each rank's frame-end stores its rows of the G-buffers + reservoirs
    (those within 16 rows of another rank's band; the camera is static) into the other ranks' previous-frame blocks over
    NVLink and raises a flag; the next frame's first kernel waits for the flags (gpurt_pipe_history_peers) — no host
    synchronisation or collective per frame.  The composite image on rank 0 is compared with the unsharded render."""
    from gpurt.dist import gather_to_rank0, share_history
    w, h, frames, halo = 1920, 1080, 9, 16
    scene = gpurt.Scene(ctx).load(os.path.join(ROOT, "tests", "data", "media", "mis_test", "mis_test.gltf"))
    accel = gpurt.Accel(scene)
    cam = gpurt.camera(1, w, h, (0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0)
    prm = gpurt.pipe_params(integrator=3, brdf=1, samples_per_frame=1, max_depth=4, res_samples=4, use_temporal=1, temporal_scale=16, seed=8)
    band = (h + world - 1) // world
    pipe = gpurt.RTPipe(scene, accel)
    maps = []
    if world > 1:
        pipe.set_shard(band, world, rank)
        maps = share_history(pipe, ctx, w, h, halo)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def loop(p):
        p.reset_frame()
        for _ in range(frames):
            p.render_frame(prm, cam, w, h)

    best = None
    for _ in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        loop(pipe)
        e1.record()
        torch.cuda.synchronize()
        ms = device_max(dist, world, e0.elapsed_time(e1), dev)
        best = ms if best is None else min(best, ms)
    timeouts = int(device_max(dist, world, float(pipe.history_status()[1]), dev))
    identical = None
    rows = torch.arange(h, device=dev)
    mine = pipe.device_image()[(rows // band) % world == rank] if world > 1 else pipe.device_image()
    full = gather_to_rank0(mine.contiguous().view(-1, 4)) if world > 1 else mine     # contiguous bands: rank order = row order
    if rank == 0:
        ref = gpurt.RTPipe(scene, accel)
        for _ in range(4):      # the same call sequence as the sharded pipes (reservoir history carries across reset_frame)
            loop(ref)
        identical = bool(torch.equal(ref.device_image().view(torch.int32).view(-1, 4), full.view(torch.int32).view(-1, 4)))
        ref.close()
    out = {"config": "3 (ReSTIR direct): mis_test 1920x1080, depth 4, 1 spp, res_samples 4, temporal reuse, 9 frames", "n_gpus": world,
           "ms_per_frame": best / frames, "mpaths_s": w * h * frames / (best * 1e-3) / 1e6,
           "sharding": (f"{band}-row contiguous bands; rows within {halo} rows of another band pushed into that rank's previous-frame "
                        "block by the frame-end kernels over NVLink, flag wait at the next frame start") if world > 1 else "one GPU",
           "flag_wait_timeouts": timeouts, "composite_bit_identical_to_unsharded": identical}
    if world > 1:
        dist.barrier()
    pipe.close()
    for m in maps:
        m.close()
    accel.close(), scene.close()
    return out


def strong_config5(gpurt, torch, dist, ctx, scene, accel, rank, world, dev, label, verify=True):
    """config 5: 3840x2160, 64 spp as 8 samples x 8 progressive frames, integrator 1, depth 8, RR on.  Frame f is rendered
    at full resolution by rank f mod N; its frame-end kernel stores the per-pixel mean straight into rank 0's buffer over
    NVLink; rank 0 folds the means in frame order (rt.rgen:640-645) — bit-identical to the sequential loop."""
    from gpurt.dist import shared_result_buffer
    w, h, F, spp = 3840, 2160, 8, 8
    pipe = gpurt.RTPipe(scene, accel)
    cam = gpurt.camera(1, w, h, CAM_POS, CAM_AT, VFOV)
    prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=8, samples_per_frame=spp, max_frames=F - 1, use_rr=1,
                            env_scale=1.0, seed=7)
    img_bytes = w * h * 16
    means = shared_result_buffer(ctx, F * img_bytes)
    mine = list(range(rank, F, world))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run():
        e0.record()
        for f in mine:
            pipe.render_frame_mean(prm, cam, w, h, f, means.at(f * img_bytes))
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()                                         # every mean is in rank 0's memory
        fold = 0.0
        if rank == 0:
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for f in range(F):
                pipe.accumulate_mean(means.at(f * img_bytes), f, w, h)
            f1.record()
            torch.cuda.synchronize()
            fold = f0.elapsed_time(f1)
        return fold

    run()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    fold_ms = run()
    wall = time.time() - t0
    ms = device_max(dist, world, e0.elapsed_time(e1) if mine else 0.0, dev)
    total_s = (ms + fold_ms) * 1e-3
    out = None
    if rank == 0:
        identical = None
        if verify:
            ref = gpurt.RTPipe(scene, accel)
            while ref.render_frame(prm, cam, w, h) == 0:
                pass
            identical = bool(torch.equal(ref.device_image().view(torch.int32), pipe.device_image().view(torch.int32)))
            ref.close()
        out = {"config": "5: Sponza 3840x2160, 64 spp (8 x 8 progressive frames), integrator 1, depth 8, RR", "scene": label,
               "n_gpus": world, "s_per_image": total_s, "mpaths_s": w * h * spp * F / total_s / 1e6,
               "render_ms_max_rank": ms, "fold_ms_rank0": fold_ms, "wall_s_incl_barriers": wall,
               "results": "frame means stored into rank 0's buffer by the frame-end kernel over NVLink, folded there in frame order",
               "bytes_into_rank0": int(img_bytes * (F - len(range(0, F, world)))), "bit_identical_to_sequential_render": identical}
    if world > 1:
        dist.barrier()
    means.close()
    pipe.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the config-4 / config-5 strong-scaling sub-benchmarks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    host = bind_to_gpu_numa(local)
    import torch
    import torch.distributed as dist
    import gpurt
    from gpurt.dist import gather_to_rank0, shared_result_buffer, warmup as dist_warmup

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist_warmup(dev)
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

    # ---- result placement (N > 1): every rank's hits go into rank 0's array, written by the trace kernel itself ----
    counts = [n_rays]
    if world > 1:
        t = torch.tensor([n_rays], dtype=torch.int64, device=dev)
        ts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        counts = [int(x.item()) for x in ts]
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    shared = shared_result_buffer(ctx, int(offs[-1]) * 16) if world > 1 else None
    dst = shared.at(int(offs[rank]) * 16) if world > 1 else d_hits

    def step():
        accel.trace_closest(d_rays, dst)

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
    barrier()                  # every rank's stores have landed in rank 0's memory
    wall = time.time() - wall0
    sampler.window[1] = wall0 + wall
    ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.summary()
    ms_max = device_max(dist, world, ms, dev)
    total_rays = device_max(dist, world, n_rays, dev, "sum")
    value = total_rays * args.steps / (ms_max * 1e-3) / 1e6

    # ---- placement: verification + what it costs (results left local; NCCL gather of the same bytes) --------------
    placement = None
    accel.trace_closest(d_rays, d_hits)
    if world > 1:
        def checksum(x):
            return int(x.contiguous().view(torch.int32).to(torch.int64).sum().item())
        mine = torch.tensor([checksum(d_hits)], dtype=torch.int64, device=dev)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        ok = None
        if rank == 0:
            full = shared.tensor().view(torch.float32).view(-1, 4)
            ok = all(checksum(full[offs[r]:offs[r + 1]]) == int(sums[r].item()) for r in range(world))
            assert ok, "hits placed in rank 0's array differ from the ranks' local results"
        k_local = max(5, min(args.steps, 30))
        evl = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k_local)]
        for a, b in evl:
            flush.zero_()
            a.record()
            accel.trace_closest(d_rays, d_hits)
            b.record()
        barrier()
        local_ms = device_max(dist, world, sum(a.elapsed_time(b) for a, b in evl) / k_local, dev)
        gather_to_rank0(d_hits)
        g = []
        for _ in range(3):
            barrier()
            t0 = time.time()
            gather_to_rank0(d_hits)
            torch.cuda.synchronize()
            g.append(device_max(dist, world, (time.time() - t0) * 1e3, dev))
        placement = {"how": "k_trace_closest stores its hits into rank 0's array over NVLink (gpurt_shared_*); no collective",
                     "bytes_into_rank0_per_step": int((offs[-1] - offs[1]) * 16), "verified_on_rank0": ok,
                     "ms_per_step_with_placement": ms_max / args.steps, "ms_per_step_results_left_local": local_ms,
                     "collective_ms": 0.0, "nccl_gather_ms_same_bytes": float(np.median(g)),
                     "note": "compute-then-NCCL-gather would cost results_left_local + nccl_gather per step"}

    # ---- rooflines (rank 0's numbers): the frame's ray set, primary rays only, closest-point queries ----------------
    peak, peak_src = peaks()
    ncu = ncu_facts()
    st = accel.trace_stats(d_rays, d_hits)
    kernel_ms = ms / args.steps
    roof = roofline_obj("k_trace_closest<false,false>", n_rays, kernel_ms, st.nodes_visited / st.rays, st.tris_tested / st.rays,
                        BYTES_STREAM, peak, peak_src, ncu.get("closest"))
    roof.update({"bytes_per_ray": roof["bytes_per_element"], "nodes_per_ray": roof["nodes_per_element"],
                 "tris_per_ray": roof["tris_per_element"]})
    n_p = prim.shape[0]

    def median_ms(fn, reps=7):
        out = []
        flush.zero_()
        fn()
        for _ in range(reps):
            flush.zero_()
            fn()
            out.append(ctx.last_kernel_ms())
        return float(np.median(out))

    prim_ms = median_ms(lambda: accel.trace_closest(d_rays[:n_p], d_hits[:n_p]))
    stp = accel.trace_stats(d_rays[:n_p], d_hits[:n_p])
    roof_primary = roofline_obj("k_trace_closest<false,false>", n_p, prim_ms, stp.nodes_visited / stp.rays, stp.tris_tested / stp.rays,
                                BYTES_STREAM, peak, peak_src, ncu.get("primary"))
    bounce_ms = median_ms(lambda: accel.trace_closest(d_rays[n_p:], d_hits[n_p:])) if n_rays > n_p else None
    q_np = np.zeros((n_p, 4), np.float32)
    hp = d_hits[:n_p].cpu().numpy().view(gpurt.HIT_DT).reshape(-1)
    tt = np.where(np.isfinite(hp["t"]), hp["t"], 100.0).astype(np.float32)
    jit = (lcg_randf(tea(np.arange(n_p, dtype=np.uint32), np.uint32(0xD00D + rank)))[:, None] - 0.5) * 60.0
    q_np[:, :3] = prim[:, 0:3] + 0.8 * tt[:, None] * prim[:, 4:7] + jit
    q_np[:, 3] = np.inf
    d_q = torch.from_numpy(q_np).cuda()
    d_cp = accel.closest_points(d_q)
    cpq_ms = median_ms(lambda: accel.closest_points(d_q, d_cp))
    stq = accel.closest_points_stats(d_q)
    # the timed span is the whole call: probe + Morton keys + 4 sort passes + k_closest_points<64,true> through the sorted index
    # when the probe finds the batch incoherent (these queries: it does), k_closest_points<64,false> alone otherwise
    roof_cpq = roofline_obj("k_closest_points<64,ordered> incl. probe / keys / sort of the call", n_p, cpq_ms, stq.nodes_visited / stq.rays, stq.tris_tested / stq.rays,
                            CPQ_BYTES_STREAM, peak, peak_src, ncu.get("cpq"))
    accel.trace_closest(d_rays, d_hits)

    # ---- e2e: C ABI with HOST buffers (pinned), copies inside the timed region ------------------
    # rays: written once by the application, read by the GPU -> write-combined pinned memory (no cache snoops on the DMA
    # reads; with the hits coming back at the same time this box moves 52 instead of 49 GB/s, tools/wc_probe.py)
    ray_mem, hr = "pinned", None
    try:
        import ctypes as C
        rt = C.CDLL("libcudart.so.12")
        wc_ptr = C.c_void_p()
        if rt.cudaHostAlloc(C.byref(wc_ptr), C.c_size_t(rays_np.nbytes), C.c_uint(4)) == 0:   # cudaHostAllocWriteCombined
            hr = np.frombuffer((C.c_uint8 * rays_np.nbytes).from_address(wc_ptr.value), np.float32).reshape(rays_np.shape)
            hr[:] = rays_np
            ray_mem = "pinned, write-combined"
    except OSError:
        pass
    if hr is None:
        h_rays = torch.from_numpy(rays_np).pin_memory()
        hr = h_rays.numpy()
    h_hits = torch.empty((n_rays, 4), dtype=torch.float32).pin_memory()
    hh = h_hits.numpy()
    for _ in range(3):
        accel.trace_closest(hr, hh)
    e2e_steps = max(3, min(args.steps, 10))
    blocks = []                         # 5 blocks of e2e_steps calls, each bracketed by barriers; the median block counts
    for _ in range(5):
        barrier()
        t0 = time.time()
        for _ in range(e2e_steps):
            accel.trace_closest(hr, hh)     # rays in over PCIe + kernel + hits out + sync inside the call
        barrier()
        blocks.append(device_max(dist, world, time.time() - t0, dev))
    e2e_val = total_rays * e2e_steps / float(np.median(blocks)) / 1e6
    assert np.array_equal(hh.view(np.uint32), d_hits.cpu().numpy().view(np.uint32)), "host and device paths disagree"
    # e2e of the reference-facing call of this path (RTPipe::trace -> rt_target read back): camera uniforms in, the frame's
    # rays generated, traced and shaded on the device, RGBA32F image out to pinned host memory every frame
    h_img = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    himg = h_img.numpy()
    for _ in range(2):
        pipe.reset_frame()
        pipe.render_frame(prm, cam, W, H)
        pipe.read_image(himg)
    blocks = []
    for _ in range(5):
        barrier()
        t0 = time.time()
        for _ in range(e2e_steps):
            pipe.reset_frame()
            pipe.render_frame(prm, cam, W, H)
            pipe.read_image(himg)
        barrier()
        blocks.append(device_max(dist, world, time.time() - t0, dev))
    e2e_render_s = float(np.median(blocks))
    # the same loop with the read-back queued behind each frame (gpurt_pipe_read_image_async): every frame's image still
    # lands in pinned host memory inside the timed region, but the copy of frame f overlaps the tracing / shading of f + 1
    h_img2 = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    blocks = []
    for _ in range(5):
        barrier()
        t0 = time.time()
        for i in range(e2e_steps):
            pipe.reset_frame()
            pipe.render_frame(prm, cam, W, H)
            pipe.read_image_async(h_img2[i & 1])
        pipe.read_image_wait()
        barrier()
        blocks.append(device_max(dist, world, time.time() - t0, dev))
    e2e_render_async_s = float(np.median(blocks))
    assert np.array_equal(h_img2[(e2e_steps - 1) & 1].numpy().view(np.uint32), himg.view(np.uint32)), "async read-back differs"
    frame_rays_total = device_max(dist, world, frame_rays[0], dev, "sum")

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

    # ---- tree quality: default (binned SAH) vs Morton LBVH, on the headline scene and on real meshes (N = 1) -------
    tree = None
    if world == 1:
        irr, irr_label, irr_cam = build_irregular_scene(gpurt, ctx)
        tree = {"headline_scene": tree_report(gpurt, torch, ctx, scene, dict(pos=CAM_POS, at=CAM_AT, vfov=VFOV), flush),
                "irregular_scene": dict(tree_report(gpurt, torch, ctx, irr, irr_cam, flush), scene=irr_label),
                "note": "the headline `value` uses the default build; `lbvh` = GPURT_BUILD_LBVH (round 1's only build)"}
        irr.close()

    # ---- BASELINE configs 1 (cbox: random / coherent / shadow rays, closest points) and 3 (mis_test 1080p: MIS, ReSTIR) ---
    configs13 = None
    if world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import config13
        configs13 = config13.run(ctx, flush, local)

    # ---- strong scaling at this N (configs 4 and 5) ---------------------------------------------
    strong = None
    if not args.no_strong:
        pipe.close()
        strong = {"config5": strong_config5(gpurt, torch, dist, ctx, scene, accel, rank, world, dev, label)}
        strong["config3_restir"] = strong_config3_restir(gpurt, torch, dist, ctx, rank, world, dev)
        accel.close()
        del d_rays, d_hits, flush, d_q, d_cp
        torch.cuda.empty_cache()
        strong["config4"] = strong_config4(gpurt, torch, dist, ctx, rank, world, dev, 10_000_000, 100_000_000, 100_000)
        strong["note"] = ("fixed total work sharded over the ranks of this run; divide by the N=1 line's numbers for the "
                          "strong-scaling factor")

    if rank == 0:
        sharding = ("one view per rank, scene replicated" + (", every rank's hits stored into rank 0's array by the trace kernel "
                    "inside the timed region (no collective)" if world > 1 else ", no collective in the timed region"))
        line = {
            "metric": "Mrays/s closest-hit on Sponza", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "sponza 1080p config-2 ray set (2,073,600 primary + 1-bounce rays), closest-hit",
                       "scene": label, "rays_per_gpu": n_rays, "tris": info.n_tris, "wide_nodes": info.n_wide_nodes,
                       "wide_depth": info.wide_depth, "bvh_build_ms": info.build_ms, "bvh_update_ms": update_ms,
                       "bvh_build_mtris_s": info.n_tris / (info.build_ms * 1e-3) / 1e6,
                       "l2": "flushed between timed steps (256 MiB memset)", "sharding": sharding},
            "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": int(n_rays * 32), "d2h_bytes_per_step": int(n_rays * 16),
                    "what": f"gpurt_trace_closest with HOST ray ({ray_mem}) / hit (pinned) arrays: the kernel reads the rays and "
                            "stores the hits in place over PCIe (32 B in + 16 B out per ray inside the timed call); median of 5 blocks of "
                            f"{e2e_steps} calls"},
            "e2e_render": {"value": frame_rays_total * e2e_steps / e2e_render_s / 1e6, "unit": "Mrays/s",
                           "ms_per_frame": e2e_render_s / e2e_steps * 1e3, "h2d_bytes_per_step": 416, "d2h_bytes_per_step": W * H * 16,
                           "what": "gpurt_pipe_render_frame + gpurt_pipe_read_image into pinned host memory (the reference-facing "
                                   "RTPipe::trace call: rays generated, traced and shaded on the device; closest-hit rays of the frame / wall time)",
                           "overlapped": {"value": frame_rays_total * e2e_steps / e2e_render_async_s / 1e6, "unit": "Mrays/s",
                                          "ms_per_frame": e2e_render_async_s / e2e_steps * 1e3,
                                          "what": "the same frames and the same bytes per frame with gpurt_pipe_read_image_async: the copy of "
                                                  "frame f runs on the pipe's copy stream while frame f + 1 traces and shades"}},
            "gpu_launches": args.steps,
            "roofline": roof, "roofline_primary": roof_primary, "roofline_cpq": roof_cpq,
            "cpu_baseline": cpu,
            "primary_only": {"value": n_p / (prim_ms * 1e-3) / 1e6, "unit": "Mrays/s", "rays": n_p},
            "bounce_only": {"value": (n_rays - n_p) / (bounce_ms * 1e-3) / 1e6, "unit": "Mrays/s", "rays": n_rays - n_p} if bounce_ms else None,
            "cpq": {"value": n_p / (cpq_ms * 1e-3) / 1e6, "unit": "Mqueries/s", "queries": n_p,
                    "what": "closest-point queries near the primary hit points"},
            "render": {"what": "whole config-2 frame through gpurt_pipe_render_frame (gen + trace + shade + accumulate)",
                       "ms_per_frame": float(np.median(frame_ms)), "closest_rays": frame_rays[0], "any_rays": frame_rays[1],
                       "mrays_s": frame_rays[0] / (float(np.median(frame_ms)) * 1e-3) / 1e6,
                       "mpaths_s": W * H / (float(np.median(frame_ms)) * 1e-3) / 1e6},
            "tree": tree, "configs_1_and_3": configs13, "placement": placement, "strong": strong, "host": host,
            "clocks": clocks, "wall_s": wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    if shared is not None:
        shared.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

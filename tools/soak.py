#!/usr/bin/env python
"""One-off large parity soak (development aid, uses the oracle like the tests do): tens of millions of random rays,
shadow segments and closest-point queries on the stand-in, the Cornell box and a 1 M-triangle soup, GPU vs the CPU
oracle BVH, bit for bit.  Prints mismatch counts."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpurt  # noqa: E402
import orc  # noqa: E402
from scenes import load_scene, soup, world_tris  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
ctx = gpurt.Context(0)
total_bad = 0
for name in ("sponza_standin", "cbox", "soup1m"):
    if name == "soup1m":
        tris = soup(1_000_000, seed=99, ext=0.02)
        scene = gpurt.Scene(ctx)
        scene.add_triangles(tris)
    else:
        scene = load_scene(gpurt, ctx, name)
        tris = world_tris(orc, scene)
    accel = gpurt.Accel(scene)
    ob = orc.Bvh(tris, sah=True)
    assert (accel.prim_order() == ob.prim_order()).all()
    box = ob.scene_box()
    for seed in (1, 2):
        t0 = time.time()
        rays = orc.gen_random_rays(N, 1000 + seed, box, frac=0.1 if seed == 1 else -0.05)
        bad_c = int((accel.trace_closest(rays).view(np.uint32).reshape(-1, 4) != ob.closest_hit(rays).view(np.uint32).reshape(-1, 4)).any(1).sum())
        seg = rays.copy()
        seg[:, 7] = np.random.default_rng(seed).random(N, dtype=np.float32) * np.float32(np.abs(box).max())
        bad_a = int((accel.trace_any(seg) != ob.any_hit(seg)).sum())
        q = orc.gen_random_points(N // 2, 2000 + seed, box, frac=0.25, r2=np.inf if seed == 1 else float((np.abs(box).max() * 0.05) ** 2))
        cp, cref = accel.closest_points(q), ob.closest_point(q)
        bad_q = int(((cp["prim"] != cref["gid"]) | (cp["dist"].view(np.uint32) != cref["dist"].view(np.uint32))
                     | (cp["p"].view(np.uint32).reshape(-1, 3) != cref["p"].view(np.uint32).reshape(-1, 3)).any(1)).sum())
        total_bad += bad_c + bad_a + bad_q
        print(f"{name:15s} seed {seed}: {N} rays closest {bad_c} / any {bad_a} mismatches, {N // 2} queries {bad_q} mismatches ({time.time() - t0:.0f} s)", flush=True)
    accel.close(), scene.close()
print("TOTAL MISMATCHES", total_bad)

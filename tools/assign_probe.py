#!/usr/bin/env python
"""CPU experiment (no GPU): does the assignment of a wide node's children to octant slots (csrc/bvh8.cuh collapse_assign)
explain why the SAH-built tree visits more nodes per ray than the Morton tree on the regular stand-in?

The replay library (tests/emu) is compiled once per variant (-DGPURT_ASSIGN_VARIANT=k: 0 greedy = shipped, 1 sum-optimal
assignment, 2 greedy on extent-normalised offsets, 3 both) and runs the product's collapse + traversal over the product's two
binary trees (oracle: Morton / binned SAH).  Reports wide nodes and node visits / triangle tests per ray.
usage: python tools/assign_probe.py [--scenes sponza_standin cbox] [--variants 0 1 2 3]"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-rt_b200"), ROOT, os.path.join(ROOT, "tools")]
import gpurt  # noqa: E402
import orc  # noqa: E402
from sah_probe import camera_rays, vp  # noqa: E402
from scenes import load_scene, world_tris  # noqa: E402


def emu_lib(variant):
    out = f"/tmp/libemu_assign_v{variant}.so"
    src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-w", "-shared",
                    f"-DGPURT_ASSIGN_VARIANT={variant}", "-o", out, src, "-lpthread"], check=True, cwd=os.path.dirname(src))
    lib = C.CDLL(out)
    lib.emu_build.restype = C.c_void_p
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", nargs="*", default=["sponza_standin", "cbox"])
    ap.add_argument("--variants", type=int, nargs="*", default=[0, 1, 2, 3])
    ap.add_argument("--rays", type=int, default=200000)
    args = ap.parse_args()
    cams = {"cbox": None, "mis_test": ((0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0),
            "sponza_standin": ((-1000.0, 200.0, 0.0), (0.0, 200.0, 0.0), 90.0)}
    out = {}
    for name in args.scenes:
        scene = load_scene(gpurt, None, name)
        tris = world_tris(orc, scene)
        n = len(tris)
        trees = {}
        for sah in (False, True):
            b = orc.Bvh(tris, sah=sah)
            l, r, bx = b.bvh2()
            trees["sah" if sah else "morton"] = (b.prim_order(), l, r, np.ascontiguousarray(bx))
        inflate = np.float32(max(1e-30, np.abs(b.scene_box()).max()) * 2.0 ** -19)
        w, h = 640, 360
        c = cams[name]
        cam = gpurt.camera(0, w, h) if c is None else gpurt.camera(1, w, h, c[0], c[1], c[2])
        sets = {"random": orc.gen_random_rays(args.rays, 0xC0FFEE, b.scene_box()), "primary": camera_rays(cam, w, h)}
        ph = b.closest_hit(sets["primary"])
        ok = ph["gid"] != 0xFFFFFFFF
        pr = sets["primary"][ok]
        rng = np.random.default_rng(1)
        d = rng.standard_normal((len(pr), 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        br = np.zeros((len(pr), 8), np.float32)
        br[:, 0:3] = pr[:, 0:3] + pr[:, 4:7] * ph["t"][ok][:, None]
        br[:, 3], br[:, 4:7], br[:, 7] = 1e-5, d, 1e7
        sets["bounce"] = br
        res, ref = {}, {}
        for v in args.variants:
            emu = emu_lib(v)
            for tname, (o, tl, tr, tb) in trees.items():
                hnd = C.c_void_p(emu.emu_build(vp(tris), n, vp(o), vp(tl), vp(tr), vp(tb), C.c_float(inflate)))
                e = {"wide_nodes": int(emu.emu_n_nodes(hnd))}
                for sname, rays in sets.items():
                    m = len(rays)
                    hits, cnt = np.zeros((m, 4), np.uint32), np.zeros(4, np.uint64)
                    emu.emu_trace(hnd, vp(rays), C.c_ulonglong(m), vp(hits), None, vp(cnt))
                    e[sname] = [round(float(cnt[0]) / m, 3), round(float(cnt[1]) / m, 3)]
                    if sname in ref:
                        assert (ref[sname] == hits).all(), "hits changed"
                    ref[sname] = hits
                emu.emu_free(hnd)
                res[f"v{v}_{tname}"] = e
                print(name, f"v{v}_{tname}", e, file=sys.stderr, flush=True)
        out[name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""CPU experiment (no GPU): what would a binned-SAH binary tree buy the shipped collapse + traversal?

For each scene the product's wide-BVH code (csrc/bvh8.cuh collapse + node encoding, csrc/traverse.cuh traversal,
replayed by tests/emu/libemu.so) is run on two binary trees over the same triangles: the Morton LBVH the build ships
(from the oracle) and a top-down binned-SAH tree (tests/emu/emu.cpp emu_sah_bvh2).  Reported per ray set: wide nodes,
node visits and triangle tests per ray (the instrumented traversal), and that both trees give bit-identical hits.
usage: python tools/sah_probe.py [--bins 16] [--rays 200000] [--scenes cbox mis_test sponza_standin]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-rt_b200"), ROOT]
import gpurt  # noqa: E402
import orc  # noqa: E402
from scenes import load_scene, world_tris  # noqa: E402


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def camera_rays(cam, w, h):
    iP = np.array(cam.iP, np.float32).reshape(4, 4).T
    iV = np.array(cam.iV, np.float32).reshape(4, 4).T
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    ndc = np.stack([(xs + 0.5) / w * 2 - 1, (ys + 0.5) / h * 2 - 1, np.zeros_like(xs), np.ones_like(xs)], -1).reshape(-1, 4)
    t = ndc @ iP.T
    d = np.concatenate([t[:, :3], np.zeros((len(t), 1), np.float32)], 1) @ iV.T
    d = d[:, :3] / np.linalg.norm(d[:, :3], axis=1, keepdims=True)
    rays = np.zeros((w * h, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = (iV @ np.array([0, 0, 0, 1], np.float32))[:3], 1e-5, d, 1e7
    return rays


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bins", type=int, default=16)
    ap.add_argument("--rays", type=int, default=200000)
    ap.add_argument("--scenes", nargs="*", default=["cbox", "mis_test", "sponza_standin"])
    ap.add_argument("--clusters", type=int, nargs="*", default=[8, 64])
    ap.add_argument("--ploc", type=int, nargs="*", default=[8, 32])
    args = ap.parse_args()
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libemu.so"))
    emu.emu_build.restype = C.c_void_p
    cams = {"cbox": None, "mis_test": ((0.5, 0.6, 2.6), (0.5, 0.45, 0.0), 50.0),
            "sponza_standin": ((-1000.0, 200.0, 0.0), (0.0, 200.0, 0.0), 90.0)}
    out = {}
    for name in args.scenes:
        scene = load_scene(gpurt, None, name)
        tris = world_tris(orc, scene)
        n = len(tris)
        b = orc.Bvh(tris)
        inflate = np.float32(max(1e-30, np.abs(b.scene_box()).max()) * 2.0 ** -19)
        trees = {}
        l, r, bx = b.bvh2()
        trees["lbvh"] = (b.prim_order(), l, r, bx)
        order = np.zeros(n, np.uint32)
        sl, sr, sbx = np.zeros(n - 1, np.int32), np.zeros(n - 1, np.int32), np.zeros((n - 1, 6), np.float32)
        emu.emu_sah_bvh2(vp(tris), n, args.bins, vp(order), vp(sl), vp(sr), vp(sbx))
        assert sorted(order.tolist()) == list(range(n))
        trees["binned_sah"] = (order, sl, sr, sbx)
        for radius in args.ploc:        # PLOC over the Morton order
            po = np.zeros(n, np.uint32)
            pl, pr_, pb = np.zeros(n - 1, np.int32), np.zeros(n - 1, np.int32), np.zeros((n - 1, 6), np.float32)
            rounds = emu.emu_ploc_bvh2(vp(tris), n, vp(trees["lbvh"][0]), radius, vp(po), vp(pl), vp(pr_), vp(pb))
            assert sorted(po.tolist()) == list(range(n))
            trees[f"ploc_r{radius}"] = (po, pl, pr_, pb)
            print(f"{name}: PLOC radius {radius}: {rounds} rounds", file=sys.stderr)
        for cluster in args.clusters:   # Morton subtrees of <= cluster triangles under a binned-SAH top tree
            ho = np.zeros(n, np.uint32)
            hl, hr, hb = np.zeros(n - 1, np.int32), np.zeros(n - 1, np.int32), np.zeros((n - 1, 6), np.float32)
            k = emu.emu_hybrid_bvh2(vp(tris), n, vp(trees["lbvh"][0]), vp(l), vp(r), vp(np.ascontiguousarray(bx)), cluster,
                                    args.bins, vp(ho), vp(hl), vp(hr), vp(hb))
            assert sorted(ho.tolist()) == list(range(n))
            trees[f"hybrid_{cluster}"] = (ho, hl, hr, hb)
            print(f"{name}: {k} clusters of <= {cluster} triangles", file=sys.stderr)
        w, h = 640, 360
        c = cams[name]
        cam = gpurt.camera(0, w, h) if c is None else gpurt.camera(1, w, h, c[0], c[1], c[2])
        sets = {"random": orc.gen_random_rays(args.rays, 0xC0FFEE, b.scene_box()), "primary": camera_rays(cam, w, h)}
        # bounce rays: from the primary hit points into random directions
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
        res = {"tris": n}
        ref = {}
        variants = [(t, 1) for t in trees] + [(t, 0) for t in ("lbvh", "binned_sah")]
        for tname, greedy in variants:
            o, tl, tr, tb = trees[tname]
            emu.emu_set_greedy(greedy)   # 0: the SAH-optimal wide collapse (GPURT_BUILD_SAH_COLLAPSE)
            tname = tname if greedy else tname + "+dp_collapse"
            hnd = C.c_void_p(emu.emu_build(vp(tris), n, vp(o), vp(tl), vp(tr), vp(np.ascontiguousarray(tb)), C.c_float(inflate)))
            assert emu.emu_depth(hnd) < 60, "too deep for the traversal stack"
            e = {"wide_nodes": int(emu.emu_n_nodes(hnd)), "wide_depth": int(emu.emu_depth(hnd))}
            for sname, rays in sets.items():
                m = len(rays)
                hits, cnt = np.zeros((m, 4), np.uint32), np.zeros(4, np.uint64)
                emu.emu_trace(hnd, vp(rays), C.c_ulonglong(m), vp(hits), None, vp(cnt))
                e[sname] = {"rays": m, "nodes_per_ray": float(cnt[0]) / m, "tris_per_ray": float(cnt[1]) / m}
                if sname in ref:
                    assert (ref[sname] == hits).all(), "the two trees disagree on a hit"
                ref[sname] = hits
            q = orc.gen_random_points(min(args.rays, 100000), 0xFACADE, b.scene_box())
            cres, ccnt = np.zeros((len(q), 8), np.uint32), np.zeros(2, np.uint64)
            emu.emu_cpq_counts(vp(ccnt))                       # reset
            emu.emu_cpq(hnd, vp(q), C.c_ulonglong(len(q)), vp(cres))
            emu.emu_cpq_counts(vp(ccnt))
            e["closest_point"] = {"queries": len(q), "nodes_per_query": float(ccnt[0]) / len(q), "tris_per_query": float(ccnt[1]) / len(q)}
            if "cpq" in ref:
                assert (ref["cpq"][:, 3] == cres[:, 3]).all(), "the two trees disagree on a closest-point distance"
            ref["cpq"] = cres
            emu.emu_free(hnd)
            emu.emu_set_greedy(1)
            res[tname] = e
        out[name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""CPU replay of the product's host+device BVH8 code (collapse, node encoding, octant-ordered
traversal, closest-point descent) against the oracle — catches logic errors without a GPU.
tests/emu/libemu.so is test-only; the product never links it."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import MEDIA, ROOT
from scenes import soup, world_tris


@pytest.fixture(scope="module")
def emu(built):
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libemu.so"))
    lib.emu_build.restype = C.c_void_p
    return lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(emu, orc, tris, n=20000, box=None):
    b = orc.Bvh(tris)
    order = b.prim_order()
    l, r, bx = b.bvh2()
    inflate = np.float32(max(1e-30, np.abs(b.scene_box()).max()) * 2.0 ** -19)   # N7 global part, as csrc/build.cu
    h = C.c_void_p(emu.emu_build(_vp(tris), len(tris), _vp(order), _vp(l), _vp(r), _vp(bx), C.c_float(inflate)))
    assert emu.emu_depth(h) < 60
    box = b.scene_box() if box is None else box
    rays = orc.gen_random_rays(n, 0xC0FFEE, box)
    hits, occ, cnt = np.zeros((n, 4), np.uint32), np.zeros(n, np.uint8), np.zeros(4, np.uint64)
    emu.emu_trace(h, _vp(rays), C.c_ulonglong(n), _vp(hits), _vp(occ), _vp(cnt))
    assert (hits == b.closest_hit(rays).view(np.uint32).reshape(-1, 4)).all()
    assert (occ == b.any_hit(rays)).all()
    for r2 in (np.inf, 0.003):
        q = orc.gen_random_points(n, 0xFACADE, box, r2=r2)
        res = np.zeros((n, 8), np.uint32)
        emu.emu_cpq(h, _vp(q), C.c_ulonglong(n), _vp(res))
        assert (res == b.closest_point(q).view(np.uint32).reshape(-1, 8)).all()
    nodes = emu.emu_n_nodes(h)
    emu.emu_free(h)
    return nodes, cnt[0] / n, cnt[1] / n


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 100, 5000, 60000])
def test_soups(emu, orc, n):
    tris = soup(n, seed=n + 1)
    if n >= 100:
        tris[10:14] = tris[10]
    _check(emu, orc, tris, 8000)


def test_sah_optimal_collapse_same_results_fewer_nodes(emu, orc, gpurt):
    """GPURT_BUILD_SAH_COLLAPSE (dp_node / collapse_node_dp): identical hits and closest points, fewer wide nodes"""
    s = gpurt.Scene(None)
    s.load(os.path.join(MEDIA, "cbox/cbox.gltf"))
    tris = world_tris(orc, s)
    try:
        emu.emu_set_greedy(1)
        greedy_nodes, g_npr, _ = _check(emu, orc, tris, 20000)
        emu.emu_set_greedy(0)
        sah_nodes, s_npr, _ = _check(emu, orc, tris, 20000)
        _check(emu, orc, soup(5000, seed=3), 8000)
    finally:
        emu.emu_set_greedy(1)
    assert sah_nodes < greedy_nodes and s_npr < g_npr * 1.02


@pytest.mark.parametrize("name", ["cube", "mis_test", "cbox"])
def test_reference_scenes(emu, orc, gpurt, name):
    s = gpurt.Scene(None)
    s.load(os.path.join(MEDIA, {"cube": "cube.gltf", "mis_test": "mis_test/mis_test.gltf", "cbox": "cbox/cbox.gltf"}[name]))
    tris = world_tris(orc, s)
    nodes, npr, tpr = _check(emu, orc, tris, 20000)
    assert nodes <= len(tris) // 2 + 2 and npr < 64


def test_flat_grid_with_duplicate_layer(emu, orc):
    g = []
    for _ in range(2):
        for i in range(20):
            for j in range(20):
                x0, x1, y0, y1 = i / 20, (i + 1) / 20, j / 20, (j + 1) / 20
                g += [[x0, y0, 0.5, x1, y0, 0.5, x1, y1, 0.5], [x0, y0, 0.5, x1, y1, 0.5, x0, y1, 0.5]]
    _check(emu, orc, np.array(g, np.float32), 20000, box=np.array([0, 0, 0, 1, 1, 1], np.float32))


def test_adversarial_inputs(emu, orc):
    """degenerate triangles, axis-aligned / in-plane / zero / NaN rays, on-surface and NaN queries:
    the BVH8 code, the oracle BVH and brute force agree bit for bit"""
    from scenes import adversarial_points, adversarial_rays, adversarial_scene
    tris = adversarial_scene()
    b = orc.Bvh(tris)
    order = b.prim_order()
    l, r, bx = b.bvh2()
    inflate = np.float32(max(1e-30, np.abs(b.scene_box()).max()) * 2.0 ** -19)
    h = C.c_void_p(emu.emu_build(_vp(tris), len(tris), _vp(order), _vp(l), _vp(r), _vp(bx), C.c_float(inflate)))
    rays = adversarial_rays(tris)
    n = len(rays)
    hits, occ, cnt = np.zeros((n, 4), np.uint32), np.zeros(n, np.uint8), np.zeros(4, np.uint64)
    emu.emu_trace(h, _vp(rays), C.c_ulonglong(n), _vp(hits), _vp(occ), _vp(cnt))
    ref = b.closest_hit(rays).view(np.uint32).reshape(-1, 4)
    brute = orc.closest_hit_brute(tris, rays).view(np.uint32).reshape(-1, 4)
    assert (ref == brute).all(), f"oracle BVH vs brute force: rays {np.nonzero((ref != brute).any(1))[0][:10]}"
    assert (hits == ref).all(), f"BVH8 vs oracle: rays {np.nonzero((hits != ref).any(1))[0][:10]}"
    assert (occ == b.any_hit(rays)).all()
    q = adversarial_points(tris)
    res = np.zeros((len(q), 8), np.uint32)
    emu.emu_cpq(h, _vp(q), C.c_ulonglong(len(q)), _vp(res))
    cref = b.closest_point(q).view(np.uint32).reshape(-1, 8)
    cbrute = orc.closest_point_brute(tris, q).view(np.uint32).reshape(-1, 8)
    assert (cref == cbrute).all(), f"oracle BVH vs brute force: queries {np.nonzero((cref != cbrute).any(1))[0][:10]}"
    assert (res == cref).all(), f"BVH8 vs oracle: queries {np.nonzero((res != cref).any(1))[0][:10]}"
    emu.emu_free(h)


def test_experimental_trees_give_the_same_hits(emu, orc):
    """tools/sah_probe.py's binary trees (binned SAH, PLOC, SAH top over Morton clusters) through the product's
    collapse + traversal: any valid tree must give the oracle's hits and closest points"""
    tris = soup(3000, seed=77, ext=0.08)
    n = len(tris)
    b = orc.Bvh(tris)
    l, r, bx = b.bvh2()
    inflate = np.float32(max(1e-30, np.abs(b.scene_box()).max()) * 2.0 ** -19)
    rays = orc.gen_random_rays(20000, 0xC0FFEE, b.scene_box())
    want = b.closest_hit(rays).view(np.uint32).reshape(-1, 4)
    q = orc.gen_random_points(5000, 0xFACADE, b.scene_box())
    want_cp = b.closest_point(q).view(np.uint32).reshape(-1, 8)

    def arrays():
        return np.zeros(n, np.uint32), np.zeros(n - 1, np.int32), np.zeros(n - 1, np.int32), np.zeros((n - 1, 6), np.float32)
    trees = []
    o, tl, tr, tb = arrays()
    emu.emu_sah_bvh2(_vp(tris), n, 16, _vp(o), _vp(tl), _vp(tr), _vp(tb))
    trees.append((o, tl, tr, tb))
    o, tl, tr, tb = arrays()
    emu.emu_ploc_bvh2(_vp(tris), n, _vp(b.prim_order()), 8, _vp(o), _vp(tl), _vp(tr), _vp(tb))
    trees.append((o, tl, tr, tb))
    o, tl, tr, tb = arrays()
    emu.emu_hybrid_bvh2(_vp(tris), n, _vp(b.prim_order()), _vp(l), _vp(r), _vp(np.ascontiguousarray(bx)), 64, 16, _vp(o), _vp(tl),
                        _vp(tr), _vp(tb))
    trees.append((o, tl, tr, tb))
    for o, tl, tr, tb in trees:
        assert sorted(o.tolist()) == list(range(n))
        h = C.c_void_p(emu.emu_build(_vp(tris), n, _vp(o), _vp(tl), _vp(tr), _vp(tb), C.c_float(inflate)))
        assert emu.emu_depth(h) < 60
        hits, cnt = np.zeros((len(rays), 4), np.uint32), np.zeros(4, np.uint64)
        emu.emu_trace(h, _vp(rays), C.c_ulonglong(len(rays)), _vp(hits), None, _vp(cnt))
        assert (hits == want).all()
        res = np.zeros((len(q), 8), np.uint32)
        emu.emu_cpq(h, _vp(q), C.c_ulonglong(len(q)), _vp(res))
        assert (res == want_cp).all()
        emu.emu_free(h)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 17, 1000, 20000])
def test_sah_split_tree_is_consistent_and_deterministic(emu, orc, n):
    """host/sah_split.h (GPURT_BUILD_SAH_SPLIT): order is a permutation, parents / ranges are what the refit and
    collapse kernels expect, duplicates and coincident centroids do not break it, and the wide BVH built from it
    answers like the oracle"""
    tris = soup(n, seed=n + 5, ext=0.05)
    if n >= 17:
        tris[3:9] = tris[3]                 # duplicates: centroids coincide -> middle split
        tris[10, 3:] = np.tile(tris[10, :3], 2)   # a point-sized triangle
    assert emu.emu_sah_split_check(_vp(tris), n) == 0
    if n == 1000:                           # degenerate / huge-offset / non-finite input keeps the tree well-formed
        from scenes import adversarial_scene
        adv = adversarial_scene()
        assert emu.emu_sah_split_check(_vp(adv), len(adv)) == 0
        bad = tris.copy()
        bad[5, 0], bad[6, 4], bad[7, :] = np.nan, np.inf, -np.inf
        assert emu.emu_sah_split_check(_vp(bad), len(bad)) == 0
    if n < 2:
        return
    o, tl, tr, tb = np.zeros(n, np.uint32), np.zeros(n - 1, np.int32), np.zeros(n - 1, np.int32), np.zeros((n - 1, 6), np.float32)
    emu.emu_sah_bvh2(_vp(tris), n, 16, _vp(o), _vp(tl), _vp(tr), _vp(tb))
    b = orc.Bvh(tris)
    inflate = np.float32(max(1e-30, np.abs(b.scene_box()).max()) * 2.0 ** -19)
    h = C.c_void_p(emu.emu_build(_vp(tris), n, _vp(o), _vp(tl), _vp(tr), _vp(tb), C.c_float(inflate)))
    assert emu.emu_depth(h) < 60
    rays = orc.gen_random_rays(5000, 0xC0FFEE, b.scene_box())
    hits, cnt = np.zeros((len(rays), 4), np.uint32), np.zeros(4, np.uint64)
    emu.emu_trace(h, _vp(rays), C.c_ulonglong(len(rays)), _vp(hits), None, _vp(cnt))
    assert (hits == b.closest_hit(rays).view(np.uint32).reshape(-1, 4)).all()
    emu.emu_free(h)


def test_sah_probe_tool_runs(built):
    """tools/sah_probe.py (the CPU tree-quality experiment DESIGN.md §4 quotes) still runs and its trees agree"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sah_probe.py"), "--scenes", "mis_test", "--rays", "20000",
                          "--ploc", "8", "--clusters", "64"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    j = json.loads(out.stdout)["mis_test"]
    assert {"lbvh", "binned_sah", "ploc_r8", "hybrid_64", "lbvh+dp_collapse"} <= set(j)
    assert j["binned_sah"]["bounce"]["nodes_per_ray"] <= j["lbvh"]["bounce"]["nodes_per_ray"] * 1.05

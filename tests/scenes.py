"""Shared scene helpers for the tests (world-space triangles through the oracle's flatten)."""
import os

import numpy as np

from conftest import MEDIA


def load_scene(gpurt, ctx, name):
    s = gpurt.Scene(ctx)
    if name == "cbox":
        s.load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    elif name == "mis_test":
        s.load(os.path.join(MEDIA, "mis_test", "mis_test.gltf"))
    elif name == "cube":
        s.load(os.path.join(MEDIA, "cube.gltf"))
    elif name == "sponza_standin":
        s.make_sponza_standin()
    elif name == "sponza":
        # media/sponza when a complete copy is at hand (GPURT_SPONZA_GLTF, as in bench.py; Sponza.bin is missing from the
        # reference snapshot), the labelled stand-in of the same size otherwise
        path = os.environ.get("GPURT_SPONZA_GLTF")
        if path and os.path.exists(path):
            s.load(path)
        else:
            s.make_sponza_standin()
    else:
        raise KeyError(name)
    return s


def world_tris(orc, scene):
    """oracle-side flattening (N1) of a gpurt.Scene -> (n,9) f32 in global primitive order"""
    parts = [orc.flatten(*scene.object(i), np.array(d.model, np.float32)) for i, d in enumerate(scene.descs())]
    return np.concatenate(parts) if parts else np.zeros((0, 9), np.float32)


def soup(n, seed=1, ext=0.05):
    rng = np.random.default_rng(seed)
    c = rng.random((n, 1, 3), dtype=np.float32)
    return (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * ext).reshape(n, 9).astype(np.float32)


def same_bits(a, b):
    a = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    b = np.ascontiguousarray(b).reshape(-1).view(np.uint8)
    return a.size == b.size and bool((a == b).all())


def adversarial_scene():
    """(n,9) f32 triangles chosen to stress exact-arithmetic corners: coplanar axis-aligned quads with
    shared edges, degenerate (point / collinear / repeated-vertex) triangles, exact duplicates, slivers,
    one scene-spanning triangle, a cluster at 1e-6 scale and a cluster offset to 4096 (coarse fp32 grid)."""
    rng = np.random.default_rng(1234)
    t = []
    for i in range(8):
        for j in range(8):
            x0, x1, y0, y1 = i / 8, (i + 1) / 8, j / 8, (j + 1) / 8
            t += [[x0, y0, 0.25, x1, y0, 0.25, x1, y1, 0.25], [x0, y0, 0.25, x1, y1, 0.25, x0, y1, 0.25]]   # z plane
            t += [[0.5, x0, y0, 0.5, x1, y0, 0.5, x1, y1], [0.5, x0, y0, 0.5, x1, y1, 0.5, x0, y1]]         # x plane
    t += [[0.3, 0.3, 0.6] * 3]                                         # a point
    t += [[0.1, 0.1, 0.7, 0.2, 0.2, 0.7, 0.3, 0.3, 0.7]]               # collinear
    t += [[0.6, 0.6, 0.6, 0.6, 0.6, 0.6, 0.7, 0.6, 0.6]]               # repeated vertex
    t += [[0.2, 0.7, 0.8, 0.4, 0.7, 0.8, 0.3, 0.9, 0.8]] * 3           # exact duplicates
    t += [[0.0, 0.0, 0.9, 1.0, 1e-7, 0.9, 1.0, 0.0, 0.9]]              # sliver
    t += [[-2.0, -2.0, 0.05, 3.0, -2.0, 0.05, 0.5, 3.0, 0.05]]         # scene-spanning
    c = rng.random((40, 1, 3), dtype=np.float32) * 1e-6
    t += (c + (rng.random((40, 3, 3), dtype=np.float32) - 0.5) * 1e-6).reshape(40, 9).tolist()
    c = rng.random((40, 1, 3), dtype=np.float32) + 4096.0
    t += (c + (rng.random((40, 3, 3), dtype=np.float32) - 0.5) * 0.5).reshape(40, 9).tolist()
    t += soup(200, seed=77, ext=0.2).tolist()
    return np.array(t, np.float32)


def adversarial_rays(tris, seed=5):
    """(n,8) f32 rays: axis-aligned, aimed exactly at vertices / edge midpoints / centroids, starting on
    triangle planes, lying inside triangle planes, zero and denormal directions, odd [tmin,tmax]
    intervals, NaN / inf components"""
    rng = np.random.default_rng(seed)
    v = tris.reshape(-1, 3, 3)
    rays = []

    def add(o, d, tmin=1e-5, tmax=1e7):
        rays.append([o[0], o[1], o[2], tmin, d[0], d[1], d[2], tmax])

    axes = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    for a in axes:                                   # axis-aligned through grid lines, cell centres, vertices
        for x in np.linspace(0, 1, 17):
            for y in np.linspace(0, 1, 17):
                o = {0: (-1.0 * a[0] + 0.5 * (1 - abs(a[0])), x, y), 1: (x, -1.0 * a[1] + 0.5 * (1 - abs(a[1])), y),
                     2: (x, y, -1.0 * a[2] + 0.5 * (1 - abs(a[2])))}[int(np.argmax(np.abs(a)))]
                add(o, a)
    pick = rng.integers(0, len(v), 600)
    for k in pick:                                   # exactly at vertices, edge midpoints, centroids
        o = rng.random(3).astype(np.float32) * 2 - 0.5
        for target in (v[k, 0], v[k, 1], (v[k, 0] + v[k, 1]) * np.float32(0.5), v[k].mean(0, dtype=np.float32)):
            d = (target - o).astype(np.float32)
            n = np.float32(np.sqrt((d * d).sum(dtype=np.float32)))
            if n > 0:
                add(o, d / n)
    for k in pick[:200]:                             # origin on the triangle (t = 0 candidates), and in-plane rays
        c = v[k].mean(0, dtype=np.float32)
        add(c, (0.0, 0.0, 1.0), 0.0)
        add(c, (0.0, 0.0, 1.0), -1.0)
        e = (v[k, 1] - v[k, 0]).astype(np.float32)
        n = np.float32(np.sqrt((e * e).sum(dtype=np.float32)))
        if n > 0:
            add(v[k, 0], e / n, 0.0)
    add((0.5, 0.5, -1), (0, 0, 0))                   # zero direction
    add((0.5, 0.5, -1), (1e-40, 1e-42, 1.0))         # denormal components
    add((0.5, 0.5, -1), (0, 0, 1), 2.0, 1.0)         # tmin > tmax
    add((0.5, 0.5, -1), (0, 0, 1), 0.0, 0.0)
    add((0.5, 0.5, -1), (0, 0, 1), 0.0, np.inf)
    add((0.5, 0.5, -1), (0, 0, 1), 1.25, 1.25)       # interval closed on a hit at exactly t = 1.25 (exclusive ends)
    add((0.5, 0.5, -1), (0, 0, 1), 1.25, 1e7)
    add((0.5, 0.5, -1), (0, 0, 1), 1e-5, 1.25)
    add((np.nan, 0.5, -1), (0, 0, 1))
    add((0.5, 0.5, -1), (np.nan, 0, 1))
    add((np.inf, 0.5, -1), (-1, 0, 0))
    add((0.5, 0.5, -1), (0, 0, np.inf))
    add((4096.5, 4096.5, 4090), (0, 0, 1))
    add((5e-7, 5e-7, -1), (0, 0, 1))
    return np.array(rays, np.float32)


def adversarial_points(tris, seed=6):
    """(n,4) f32 queries: on vertices / edges / faces, r2 = 0 / tiny / inf / negative, far away, NaN"""
    rng = np.random.default_rng(seed)
    v = tris.reshape(-1, 3, 3)
    q = []
    for k in rng.integers(0, len(v), 500):
        for p in (v[k, 0], (v[k, 0] + v[k, 1]) * np.float32(0.5), v[k].mean(0, dtype=np.float32)):
            for r2 in (np.inf, 0.0, 1e-12):
                q.append([p[0], p[1], p[2], r2])
    for p in ((0.5, 0.5, 0.5), (1e6, -1e6, 3e5), (4096.5, 4096.5, 4096.5), (0, 0, 0), (np.nan, 0, 0), (np.inf, 0, 0)):
        for r2 in (np.inf, 1.0, -1.0, np.nan):
            q.append([p[0], p[1], p[2], r2])
    return np.array(q, np.float32)

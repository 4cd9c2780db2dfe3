"""The oracle's query semantics against an INDEPENDENT float64 numpy restatement (different formulas, no shared
code): the reference's BVH / traversal live in the NVIDIA driver and the FCPW branch is not in the snapshot
(SURVEY §8c), so this is the second opinion that the oracle computes what the published definitions say:

  * closest hit (Vulkan ray / triangle rules, rt.rgen:257-270): the triangle plane is intersected and the point is
    classified with three edge functions — not Moeller-Trumbore — in float64; hit t within 1e-5 relative
    (BASELINE north_star), primitive ids equal except where two candidates are closer than that tolerance;
  * closest point (FCPW semantics: nearest point on any triangle, its distance and primitive): plane projection +
    three clamped segment projections — not Ericson's region test — in float64; distance within 1e-5 relative."""
import numpy as np

from scenes import soup

REL = 1e-5


def _ray_hits_f64(tris, rays):
    """(n_rays, n_tris) t of the plane / edge-function test, inf where there is no hit in (tmin, tmax)"""
    T = tris.astype(np.float64).reshape(-1, 3, 3)
    o, tmin = rays[:, 0:3].astype(np.float64), rays[:, 3].astype(np.float64)
    d, tmax = rays[:, 4:7].astype(np.float64), rays[:, 7].astype(np.float64)
    a, b, c = T[:, 0], T[:, 1], T[:, 2]
    n = np.cross(b - a, c - a)                                   # (m,3)
    denom = d @ n.T                                              # (r,m)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = ((a * n).sum(1)[None, :] - o @ n.T) / denom
    p = o[:, None, :] + t[:, :, None] * d[:, None, :]            # (r,m,3)

    def edge(u, v):                                              # signed area of (u, v, p) along n
        return (np.cross(v - u, p - u[None]) * n[None]).sum(2)
    area2 = (n * n).sum(1)[None, :]
    w0, w1, w2 = edge(b, c) / area2, edge(c, a) / area2, edge(a, b) / area2
    eps = 1e-9
    inside = (w0 >= -eps) & (w1 >= -eps) & (w2 >= -eps)
    ok = inside & (denom != 0) & (t > tmin[:, None]) & (t < tmax[:, None])
    return np.where(ok, t, np.inf), np.minimum(np.minimum(w0, w1), w2)


def _closest_points_f64(tris, pts):
    """(n_pts, n_tris) distances by plane projection + clamped segment projections"""
    T = tris.astype(np.float64).reshape(-1, 3, 3)
    p = pts[:, 0:3].astype(np.float64)
    a, b, c = T[:, 0], T[:, 1], T[:, 2]
    n = np.cross(b - a, c - a)
    nn = (n * n).sum(1)
    best = np.full((len(p), len(T)), np.inf)
    # interior: foot of the perpendicular, if it falls inside the triangle
    with np.errstate(divide="ignore", invalid="ignore"):
        h = ((p[:, None, :] - a[None]) * n[None]).sum(2) / nn[None]   # signed height / |n|^2
    foot = p[:, None, :] - h[:, :, None] * n[None]

    def edge(u, v):
        return (np.cross(v - u, foot - u[None]) * n[None]).sum(2)
    inside = (edge(a, b) >= 0) & (edge(b, c) >= 0) & (edge(c, a) >= 0) & (nn[None] > 0)
    d_in = np.abs(h) * np.sqrt(nn)[None]
    best = np.where(inside, d_in, best)
    for u, v in ((a, b), (b, c), (c, a)):
        e = v - u
        ee = (e * e).sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.clip(((p[:, None, :] - u[None]) * e[None]).sum(2) / np.where(ee > 0, ee, 1)[None], 0, 1)
        q = u[None] + s[:, :, None] * e[None]
        best = np.minimum(best, np.linalg.norm(p[:, None, :] - q, axis=2))
    return best


def test_closest_hit_matches_an_independent_float64_implementation(orc):
    tris = soup(300, seed=21, ext=0.2)
    b = orc.Bvh(tris)
    rays = orc.gen_random_rays(6000, 0xC0FFEE, b.scene_box())
    got = b.closest_hit(rays)
    t64, margin = _ray_hits_f64(tris, rays)
    ref_t = t64.min(axis=1)
    ref_id = t64.argmin(axis=1)
    hit = got["gid"] != 0xFFFFFFFF
    ref_hit = np.isfinite(ref_t)
    # rays that graze an edge within float32 resolution may be classified either way: look at the best margin
    r = np.arange(len(rays))
    graze = np.abs(margin[r, np.where(ref_hit, ref_id, 0)]) < 1e-5
    if hit.any():
        gm = np.abs(margin[r[hit], got["gid"][hit]]) < 1e-5
        graze[np.flatnonzero(hit)[gm]] = True
    assert (hit == ref_hit)[~graze].all()
    both = hit & ref_hit & ~graze
    assert both.sum() > 800
    rel = np.abs(got["t"][both].astype(np.float64) - ref_t[both]) / ref_t[both]
    assert rel.max() <= REL, rel.max()
    # same primitive, unless the runner-up is within the tolerance of the winner
    diff = both & (got["gid"] != ref_id)
    for i in np.flatnonzero(diff):
        assert abs(t64[i, got["gid"][i]] - ref_t[i]) <= REL * ref_t[i]
    assert diff.sum() <= 3


def test_closest_point_matches_an_independent_float64_implementation(orc):
    tris = soup(300, seed=22, ext=0.2)
    b = orc.Bvh(tris)
    q = orc.gen_random_points(3000, 0xFACADE, b.scene_box())
    got = b.closest_point(q)
    d64 = _closest_points_f64(tris, q)
    ref_d, ref_id = d64.min(axis=1), d64.argmin(axis=1)
    rel = np.abs(got["dist"].astype(np.float64) - ref_d) / np.maximum(ref_d, 1e-12)
    assert rel.max() <= REL, rel.max()
    # the reported point lies on the reported triangle at the reported distance
    p = got["p"].astype(np.float64)
    if True:
        assert np.abs(np.linalg.norm(p - q[:, :3].astype(np.float64), axis=1) - got["dist"]).max() <= 1e-5 * max(1.0, ref_d.max())
        on_tri = _closest_points_f64(tris, np.concatenate([p, np.zeros((len(p), 1))], axis=1).astype(np.float32))
        assert on_tri[np.arange(len(p)), got["gid"]].max() <= 1e-5
    diff = got["gid"] != ref_id
    for i in np.flatnonzero(diff):
        assert abs(d64[i, got["gid"][i]] - ref_d[i]) <= REL * max(ref_d[i], 1e-12)
    # radius-limited queries: found iff something lies within the radius
    r2 = np.float32(0.02)
    qr = orc.gen_random_points(3000, 0xFACADE, b.scene_box(), r2=r2)
    gr = b.closest_point(qr)
    found = gr["gid"] != 0xFFFFFFFF
    near = ref_d ** 2 <= float(r2)
    border = np.abs(ref_d ** 2 - float(r2)) <= 1e-5 * float(r2)
    assert (found == near)[~border].all() and found.any() and (~found).any()


def test_reference_scene_mis_test_against_float64(orc, gpurt):
    """the same two checks on a shipped scene (media/mis_test: spheres, plates, three emitters; 1,544 triangles)"""
    import os
    from conftest import MEDIA
    from scenes import world_tris
    s = gpurt.Scene(None).load(os.path.join(MEDIA, "mis_test", "mis_test.gltf"))
    tris = world_tris(orc, s)
    b = orc.Bvh(tris)
    rays = orc.gen_random_rays(2000, 0xC0FFEE, b.scene_box())
    got = b.closest_hit(rays)
    pts = orc.gen_random_points(2000, 0xFACADE, b.scene_box())
    cp = b.closest_point(pts)
    n_hit = 0
    for i0 in range(0, 2000, 250):
        sl = slice(i0, i0 + 250)
        t64, margin = _ray_hits_f64(tris, rays[sl])
        ref_t, ref_id = t64.min(axis=1), t64.argmin(axis=1)
        g = got[sl]
        hit, ref_hit = g["gid"] != 0xFFFFFFFF, np.isfinite(ref_t)
        r = np.arange(250)
        graze = np.abs(margin[r, np.where(ref_hit, ref_id, 0)]) < 1e-5
        graze |= hit & (np.abs(margin[r, np.where(hit, g["gid"], 0)]) < 1e-5)
        assert (hit == ref_hit)[~graze].all()
        both = hit & ref_hit & ~graze
        n_hit += int(both.sum())
        assert (np.abs(g["t"][both] - ref_t[both]) <= REL * ref_t[both]).all()
        for i in np.flatnonzero(both & (g["gid"] != ref_id)):
            assert abs(t64[i, g["gid"][i]] - ref_t[i]) <= REL * ref_t[i]
        d64 = _closest_points_f64(tris, pts[sl])
        ref_d = d64.min(axis=1)
        c = cp[sl]
        assert (np.abs(c["dist"] - ref_d) <= REL * np.maximum(ref_d, 1e-9)).all()
        assert (np.abs(d64[r, c["gid"]] - ref_d) <= REL * np.maximum(ref_d, 1e-9)).all()
    assert n_hit > 300

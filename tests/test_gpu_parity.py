"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle, bit-exact."""
import os

import numpy as np
import pytest

from scenes import load_scene, same_bits, soup, world_tris

pytestmark = pytest.mark.gpu


def _check_build(gpurt, orc, accel, tris):
    """keys / primitive order / binary topology / node boxes against the oracle's tree of the same kind: the binned-SAH
    split of the default build (N6'), or the Morton LBVH behind GPURT_BUILD_LBVH (N6)"""
    ob = orc.Bvh(tris, sah=not (accel.flags & gpurt.BUILD_LBVH))
    info = accel.info()
    assert info.n_tris == tris.shape[0]
    if tris.shape[0]:
        assert same_bits(np.array(list(info.scene_min) + list(info.scene_max), np.float32), ob.scene_box())
        assert 0 < info.inflation <= 1e-4 * max(1e-30, np.abs(ob.scene_box()).max())  # N7 padding is tiny
    assert (accel.morton_keys() == ob.keys()).all(), "sorted keys differ"
    assert (accel.prim_order() == ob.prim_order()).all(), "canonical primitive order differs"
    l, r, b = accel.bvh2()
    ol, orr, obx = ob.bvh2()
    assert (l == ol).all() and (r == orr).all(), "binary topology differs"
    assert same_bits(b, obx), "refit boxes differ"
    return ob


def _check_queries(gpurt, orc, accel, ob, tris, n_rays, brute_n, box=None):
    if box is None:
        box = ob.scene_box() if tris.shape[0] else np.array([0, 0, 0, 1, 1, 1], np.float32)
    rays = orc.gen_random_rays(n_rays, 0xC0FFEE, box)
    hits = accel.trace_closest(rays)
    ref = ob.closest_hit(rays)
    assert same_bits(hits, ref), f"closest hit: {(hits.view(np.uint32).reshape(-1,4) != ref.view(np.uint32).reshape(-1,4)).any(axis=1).sum()} rays differ"
    if tris.shape[0] > 1:
        assert same_bits(accel.trace_closest(rays, bvh2=True), ref), "binary-LBVH trace differs"
    occ = accel.trace_any(rays)
    assert (occ == ob.any_hit(rays)).all()
    assert ((hits["prim"] != gpurt.NO_HIT) == (occ != 0)).all()
    # brute force on a subset pins the oracle BVH itself
    sub = rays[:brute_n]
    assert same_bits(orc.closest_hit_brute(tris, sub), ref[:brute_n])
    # shadow-style rays: short segments (rt.rgen:272-291)
    seg = rays.copy()
    seg[:, 7] = np.linspace(0.01, 3.0, n_rays, dtype=np.float32)
    assert (accel.trace_any(seg) == ob.any_hit(seg)).all()
    for r2 in (np.inf, 0.05):
        q = orc.gen_random_points(n_rays, 0xFACADE, box, r2=r2)
        cp = accel.closest_points(q)
        cref = ob.closest_point(q)
        for f in ("p", "dist", "u", "v"):
            assert same_bits(cp[f], cref[f]), f"closest point field {f} (r2={r2})"
        assert (cp["prim"] == cref["gid"]).all()
        cb = orc.closest_point_brute(tris, q[:brute_n])
        assert same_bits(cb, cref[:brute_n])
    return hits


@pytest.mark.parametrize("build", ["default", "lbvh"])
@pytest.mark.parametrize("name", ["cube", "mis_test", "cbox"])
def test_reference_scenes(gpurt, orc, ctx, name, build):
    scene = load_scene(gpurt, ctx, name)
    tris = world_tris(orc, scene)
    accel = gpurt.Accel(scene, gpurt.BUILD_KEEP_BVH2 | (gpurt.BUILD_LBVH if build == "lbvh" else 0))
    ob = _check_build(gpurt, orc, accel, tris)
    n = 1 << 20 if name == "cbox" else 1 << 17   # 1,048,576 rays / queries
    hits = _check_queries(gpurt, orc, accel, ob, tris, n, 1 << 15)
    # hit ids map back to (gl_InstanceCustomIndexEXT, gl_PrimitiveID)
    offs = scene.tri_offsets()
    h = hits["prim"][hits["prim"] != gpurt.NO_HIT]
    obj = np.searchsorted(offs, h, side="right") - 1
    assert (obj >= 0).all() and (obj < len(offs) - 1).all() and (h - offs[obj] < offs[obj + 1] - offs[obj]).all()
    accel.close(), scene.close()


@pytest.mark.parametrize("build", ["default", "lbvh"])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 9, 33, 1000, 100000])
def test_soups_and_edge_sizes(gpurt, orc, ctx, n, build):
    tris = soup(n, seed=n + 7)
    if n >= 33:
        tris[10:20] = tris[10]          # duplicate triangles -> equal Morton keys, equal t ties
    scene = gpurt.Scene(ctx)
    if n:
        scene.add_triangles(tris)
    accel = gpurt.Accel(scene, gpurt.BUILD_KEEP_BVH2 | (gpurt.BUILD_LBVH if build == "lbvh" else 0))
    if n == 0:
        rays = orc.gen_random_rays(1000, 1, np.array([0, 0, 0, 1, 1, 1], np.float32))
        assert (accel.trace_closest(rays)["prim"] == gpurt.NO_HIT).all()
        assert (accel.trace_any(rays) == 0).all()
        assert (accel.closest_points(orc.gen_random_points(1000, 2, np.array([0, 0, 0, 1, 1, 1], np.float32)))["prim"] == gpurt.NO_HIT).all()
    else:
        ob = _check_build(gpurt, orc, accel, tris)
        _check_queries(gpurt, orc, accel, ob, tris, 1 << 15, 1 << 12)
    accel.close(), scene.close()


def test_coplanar_grid_ties(gpurt, orc, ctx):
    """axis-aligned quads sharing edges and exactly duplicated layers: equal-t ties -> lowest id"""
    g = []
    for layer in range(2):
        for i in range(24):
            for j in range(24):
                x0, x1, y0, y1 = i / 24, (i + 1) / 24, j / 24, (j + 1) / 24
                g.append([x0, y0, 0.5, x1, y0, 0.5, x1, y1, 0.5])
                g.append([x0, y0, 0.5, x1, y1, 0.5, x0, y1, 0.5])
    tris = np.array(g, np.float32)
    scene = gpurt.Scene(ctx)
    scene.add_triangles(tris)
    accel = gpurt.Accel(scene, gpurt.BUILD_KEEP_BVH2)
    ob = _check_build(gpurt, orc, accel, tris)
    hits = _check_queries(gpurt, orc, accel, ob, tris, 1 << 16, 1 << 12, box=np.array([0, 0, 0, 1, 1, 1], np.float32))
    hit = hits["prim"] != gpurt.NO_HIT
    assert hit.any() and (hits["prim"][hit] < len(g) // 2).all(), "ties must resolve to the lower (first-layer) id"
    accel.close(), scene.close()


def test_device_buffers_match_host_buffers(gpurt, orc, ctx):
    import torch
    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    box = np.array(list(accel.info().scene_min) + list(accel.info().scene_max), np.float32)
    rays = orc.gen_random_rays(200000, 5, box)
    host = accel.trace_closest(rays)
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = accel.trace_closest(d_rays)
    torch.cuda.synchronize()
    assert same_bits(d_hits.cpu().numpy(), host)
    # pinned host arrays: the kernels read / write them in place over PCIe (zero-copy path of the GPURT_MEM_HOST calls)
    p_rays = torch.from_numpy(rays).pin_memory()
    p_hits = torch.empty((len(rays), 4), dtype=torch.float32).pin_memory()
    accel.trace_closest(p_rays.numpy(), p_hits.numpy())
    assert same_bits(p_hits.numpy(), host)
    p_occ = torch.empty(len(rays), dtype=torch.uint8).pin_memory()
    accel.trace_any(p_rays.numpy(), p_occ.numpy())
    assert (p_occ.numpy() == accel.trace_any(rays)).all()
    q = orc.gen_random_points(100000, 9, box)
    p_q = torch.from_numpy(q).pin_memory()
    p_cp = torch.empty((len(q), 8), dtype=torch.float32).pin_memory()
    accel.closest_points(p_q.numpy(), p_cp.numpy())
    assert same_bits(p_cp.numpy(), accel.closest_points(q))
    st = accel.trace_stats(d_rays, d_hits)
    assert st.rays == 200000 and st.hits == int((host["prim"] != gpurt.NO_HIT).sum())
    assert st.nodes_visited > st.rays
    accel.close(), scene.close()


def test_sponza_standin_properties(gpurt, orc, ctx):
    """full-size stand-in (262,267 tris): oracle BVH agreement + size-independent properties"""
    scene = load_scene(gpurt, ctx, "sponza_standin")
    c = scene.counts()
    assert c["tris"] == 262267 and c["objs"] == 103 and c["lights"] == 0
    tris = world_tris(orc, scene)
    accel = gpurt.Accel(scene, gpurt.BUILD_KEEP_BVH2)
    ob = _check_build(gpurt, orc, accel, tris)
    info = accel.info()
    assert np.allclose(list(info.scene_min), [-1921, -126, -1183]) and np.allclose(list(info.scene_max), [1800, 1429, 1105])
    rays = orc.gen_random_rays(1 << 19, 11, ob.scene_box(), frac=-0.05)
    hits = accel.trace_closest(rays)
    assert same_bits(hits, ob.closest_hit(rays))
    # property: the reported hit point re-derived from barycentrics lies on the ray at t
    h = hits["prim"] != gpurt.NO_HIT
    assert h.mean() > 0.8  # origins inside the atrium; only the light-well is open to the sky
    t9 = tris[hits["prim"][h]].reshape(-1, 3, 3).astype(np.float64)
    u, v = hits["u"][h].astype(np.float64)[:, None], hits["v"][h].astype(np.float64)[:, None]
    p_tri = t9[:, 0] * (1 - u - v) + t9[:, 1] * u + t9[:, 2] * v
    p_ray = rays[h, 0:3].astype(np.float64) + hits["t"][h].astype(np.float64)[:, None] * rays[h, 4:7].astype(np.float64)
    assert np.abs(p_tri - p_ray).max() < 0.5  # scene extent ~3700 units, fp32
    # property: shortening the ray to just before / after the hit flips visibility
    seg = rays[h].copy()
    seg[:, 7] = hits["t"][h] * 0.999
    assert (accel.trace_any(seg) == 0).all()
    seg[:, 7] = hits["t"][h] * 1.001 + 1e-3
    assert (accel.trace_any(seg) == 1).all()
    # property: a closest point is never farther than any triangle vertex; distance is idempotent
    q = orc.gen_random_points(1 << 18, 13, ob.scene_box(), frac=0.1)
    cp = accel.closest_points(q)
    assert same_bits(cp["dist"], ob.closest_point(q)["dist"])
    q2 = q.copy()
    q2[:, 0:3] = cp["p"]
    assert (accel.closest_points(q2)["dist"] <= 1e-3 * 3700).all()
    accel.close(), scene.close()


def test_config4_full_size_soup(gpurt, built):
    """BASELINE config 4 at full mesh size: 10,000,000-triangle soup, 20 M sharded queries on this GPU;
    primitive order and a 100 k-query subsample are compared with the CPU oracle bit for bit."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "config4_cpq.py"), "--tris", "10000000",
                          "--queries", "20000000", "--check", "100000"], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["tris"] == 10_000_000 and res["queries"] == 20_000_000
    assert res["check"]["bit_exact_vs_oracle"] and res["check"]["prim_order_equals_oracle"]
    assert res["mqueries_s"] > 50


def test_shared_result_buffer_single_process(gpurt, orc, ctx):
    """gpurt_shared_alloc: a query writes through `buf.at(offset)`; the owner's view holds the same bytes"""
    import torch
    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    box = np.array(list(accel.info().scene_min) + list(accel.info().scene_max), np.float32)
    rays = orc.gen_random_rays(50000, 21, box)
    want = accel.trace_closest(rays)
    buf = ctx.shared_alloc(2 * 50000 * 16)
    assert len(buf.handle) == 64
    d_rays = torch.from_numpy(rays).cuda()
    accel.trace_closest(d_rays, buf.at(50000 * 16))   # second half of the buffer
    torch.cuda.synchronize()
    got = buf.tensor()[50000 * 16:].cpu().numpy().view(gpurt.HIT_DT)
    assert same_bits(got, want)
    buf.close(), accel.close(), scene.close()


def test_large_incoherent_point_batch_on_a_small_scene_is_ordered_and_unchanged(gpurt, orc, ctx):
    """closest-point batches of >= 2^20 points are processed in Morton order when a sampled probe finds them incoherent —
    on every scene above 4 MB since round 2 (order.cu; it used to take a BVH > 64 MB); the results are those of the plain path (the same points in calls
    below the batch threshold) and of the oracle, bit for bit, for device and host arrays; a coherent batch is left alone"""
    import torch
    scene = gpurt.Scene(ctx).make_sponza_standin()       # 16 MB of nodes + triangles: above the 4 MB gate, far below L2
    accel = gpurt.Accel(scene)
    ob = orc.Bvh(world_tris(orc, scene), sah=True)
    n = (1 << 20) + 777
    q = orc.gen_random_points(n, 77, ob.scene_box())
    dq = torch.from_numpy(q).cuda()
    whole = accel.closest_points(dq).cpu().numpy()
    parts = np.concatenate([accel.closest_points(dq[a:a + 300_000].contiguous()).cpu().numpy() for a in range(0, n, 300_000)])
    assert (whole.view(np.uint32) == parts.view(np.uint32)).all()
    host = accel.closest_points(q)                       # numpy in / out: the staged host pipeline answers in chunks
    assert (np.ascontiguousarray(host).view(np.uint32).reshape(-1) == whole.view(np.uint32).reshape(-1)).all()
    sub = np.arange(0, n, 97)
    ref = ob.closest_point(q[sub])
    got = whole.view(gpurt.CPQ_DT).reshape(-1)[sub]
    assert same_bits(got["dist"], ref["dist"]) and (got["prim"] == ref["gid"]).all()
    # a coherent batch (points along a scan of the floor): same answers as in small calls, too
    box = ob.scene_box()
    g = np.stack(np.meshgrid(np.linspace(box[0], box[3], 1100), np.linspace(box[2], box[5], 1000), indexing="ij"), -1).reshape(-1, 2).astype(np.float32)
    qc = np.zeros((g.shape[0], 4), np.float32)
    qc[:, 0], qc[:, 2], qc[:, 1], qc[:, 3] = g[:, 0], g[:, 1], 0.5 * (box[1] + box[4]), np.inf
    dqc = torch.from_numpy(qc).cuda()
    a = accel.closest_points(dqc).cpu().numpy()
    b = np.concatenate([accel.closest_points(dqc[k:k + 250_000].contiguous()).cpu().numpy() for k in range(0, qc.shape[0], 250_000)])
    assert (a.view(np.uint32) == b.view(np.uint32)).all()
    accel.close(), scene.close()


def test_p2p_result_placement_two_gpus(gpurt, built):
    """config 4 on two ranks with the kernels storing into rank 0's buffer over NVLink; rank 0 checks
    the half written by rank 1 against the oracle (needs 2 GPUs)"""
    import json
    import os
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731",
                          os.path.join(ROOT, "tools", "config4_cpq.py"), "--tris", "1600000", "--queries", "4400000",
                          "--chunk", "1100000", "--check", "50000", "--p2p"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    # 1.6 M triangles (BVH > 64 MB) and chunks of 1.1 M queries: the batches are Morton-ordered, staged locally and
    # written to rank 0 coalesced (order.cu)
    assert res["n_gpus"] == 2 and res["queries"] == 4_400_000 and res["check"]["bit_exact_vs_oracle"]


def test_pose_edit_update_equals_fresh_build(gpurt, orc, ctx):
    """gpurt_scene_set_transform + gpurt_accel_update (GPURT::build_accel after an edit, src/gpurt.cpp:220-241):
    the in-place rebuild equals a fresh build of the edited scene and the oracle, bit for bit."""
    def model(k):
        a = 0.3 + 0.1 * k
        m = np.eye(4, dtype=np.float32)
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
        m[:3, 3] = [0.5 * k, 0.25, -0.3 * k]
        return m.T.reshape(16).copy()  # column-major

    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    before = accel.prim_order().copy()
    n_nodes_before = accel.info().n_wide_nodes
    cam = gpurt.camera(0, 256, 256)
    params = gpurt.pipe_params(max_frames=1, samples_per_frame=2, max_depth=3, integrator=2, seed=5)
    params0 = gpurt.pipe_params(max_frames=1, samples_per_frame=2, max_depth=3, integrator=0, seed=6)
    pipe.render_frame(params, cam, 256, 256)   # builds the pipe's light BVH / light-run boxes for the OLD pose of the light
    pipe.reset_frame()
    pipe.render_frame(params0, cam, 256, 256)  # ... and its world-space light vertices (light_sample)
    light = scene.lights()[0].index

    def light_model(k):
        m = np.array(list(scene.descs()[light].model), np.float32).reshape(4, 4).T.copy()
        m[:3, 3] += np.float32(0.05 * k) * np.array([1, -1, 0.5], np.float32)
        return m.T.reshape(16).copy()
    light_models = [light_model(k) for k in (1, 2)]
    for k in (1, 2):                       # two successive edits reuse the same buffers
        scene.set_transform(3, model(k))
        scene.set_transform(7, model(k + 2))
        scene.set_transform(light, light_models[k - 1])
        accel.update()
    fresh_scene = load_scene(gpurt, ctx, "cbox")
    fresh_scene.set_transform(3, model(2))
    fresh_scene.set_transform(7, model(4))
    fresh_scene.set_transform(light, light_models[1])
    fresh = gpurt.Accel(fresh_scene)
    assert (accel.prim_order() == fresh.prim_order()).all() and (accel.morton_keys() == fresh.morton_keys()).all()
    assert not (accel.prim_order() == before).all() or accel.info().n_wide_nodes != n_nodes_before
    tris = world_tris(orc, scene)
    ob = orc.Bvh(tris, sah=True)
    assert (accel.prim_order() == ob.prim_order()).all()
    rays = orc.gen_random_rays(100000, 31, ob.scene_box())
    hits = accel.trace_closest(rays)
    assert same_bits(hits, fresh.trace_closest(rays)) and same_bits(hits, ob.closest_hit(rays))
    q = orc.gen_random_points(50000, 32, ob.scene_box())
    cp, cref = accel.closest_points(q), ob.closest_point(q)
    assert same_bits(cp["dist"], cref["dist"]) and (cp["prim"] == cref["gid"]).all()
    # the pipe created (and used) before the edit keeps working on the updated accel and matches a fresh pipe
    pipe.reset_frame()
    pipe.render_frame(params, cam, 256, 256)
    fresh_pipe = gpurt.RTPipe(fresh_scene, fresh)
    fresh_pipe.render_frame(params, cam, 256, 256)
    assert same_bits(pipe.read_image(), fresh_pipe.read_image())
    pipe.reset_frame(), fresh_pipe.reset_frame()
    pipe.render_frame(params0, cam, 256, 256)
    fresh_pipe.render_frame(params0, cam, 256, 256)
    assert same_bits(pipe.read_image(), fresh_pipe.read_image())
    for o in (pipe, fresh_pipe, accel, fresh, scene, fresh_scene):
        o.close()


def test_adversarial_inputs(gpurt, orc, ctx):
    """degenerate triangles, coplanar axis-aligned quads, slivers, 1e-6 and 4096-offset clusters; rays that
    are axis-aligned, aimed exactly at vertices / edges, lie inside triangle planes, have zero / denormal /
    NaN / inf components or odd [tmin,tmax]; queries on vertices / edges / faces with r2 = 0, NaN, negative"""
    from scenes import adversarial_points, adversarial_rays, adversarial_scene
    tris = adversarial_scene()
    scene = gpurt.Scene(ctx)
    scene.add_triangles(tris)
    accel = gpurt.Accel(scene)
    ob = _check_build(gpurt, orc, accel, tris)
    rays = adversarial_rays(tris)
    ref = ob.closest_hit(rays)
    assert same_bits(ref, orc.closest_hit_brute(tris, rays))
    hits = accel.trace_closest(rays)
    bad = np.nonzero((hits.view(np.uint32).reshape(-1, 4) != ref.view(np.uint32).reshape(-1, 4)).any(1))[0]
    assert bad.size == 0, f"rays {bad[:10]}: {rays[bad[:3]]} gpu {hits[bad[:3]]} oracle {ref[bad[:3]]}"
    assert same_bits(accel.trace_closest(rays, bvh2=True), ref)
    assert (accel.trace_any(rays) == ob.any_hit(rays)).all()
    q = adversarial_points(tris)
    cref = ob.closest_point(q)
    assert same_bits(cref, orc.closest_point_brute(tris, q))
    cp = accel.closest_points(q)
    for f in ("p", "dist", "u", "v"):
        assert same_bits(cp[f], cref[f]), f"closest point field {f}"
    assert (cp["prim"] == cref["gid"]).all()
    accel.close(), scene.close()


def test_large_scene_batches_in_spatial_order(gpurt, orc, ctx):
    """order.cu: on a scene whose BVH exceeds 64 MB, device batches of >= 2^20 incoherent rays / points are processed
    in Morton order through a sorted index; results land at their storage positions and equal the oracle's"""
    import torch
    tris = soup(1_600_000, seed=5, ext=0.01)
    scene = gpurt.Scene(ctx)
    scene.add_triangles(tris)
    accel = gpurt.Accel(scene)
    info = accel.info()
    assert info.node_bytes + info.tri_bytes > (64 << 20)
    ob = orc.Bvh(tris)
    n = (1 << 20) + 777
    rays = orc.gen_random_rays(n, 71, ob.scene_box())
    d_rays = torch.from_numpy(rays).cuda()
    hits = accel.trace_closest(d_rays)
    occ = accel.trace_any(d_rays)
    torch.cuda.synchronize()
    assert same_bits(hits.cpu().numpy(), ob.closest_hit(rays))
    assert (occ.cpu().numpy() == ob.any_hit(rays)).all()
    q = orc.gen_random_points(n, 72, ob.scene_box(), frac=0.2)
    cp = accel.closest_points(torch.from_numpy(q).cuda())
    torch.cuda.synchronize()
    cref = ob.closest_point(q)
    got = cp.cpu().numpy().view(np.uint32)
    ref = cref.view(np.uint32).reshape(-1, 8)
    assert (got[:, [0, 1, 2, 3, 4, 6, 7]] == ref[:, [0, 1, 2, 3, 4, 6, 7]]).all()
    # results on another GPU: slices of the processing order are staged and stored to their storage positions by a second
    # stream while the next slice is traversed (order.cu; GPURT_PLACE_FORCE runs that path with local results)
    for slices in (None, "5"):   # default: stage + one coalesced pass; GPURT_PLACE_SLICES: slice-by-slice scatter on a second stream
        os.environ["GPURT_PLACE_FORCE"] = "1"
        if slices:
            os.environ["GPURT_PLACE_SLICES"] = slices
        try:
            hits_p = accel.trace_closest(d_rays)
            cp_p = accel.closest_points(torch.from_numpy(q).cuda())
            occ_p = accel.trace_any(d_rays)
            torch.cuda.synchronize()
        finally:
            del os.environ["GPURT_PLACE_FORCE"]
            os.environ.pop("GPURT_PLACE_SLICES", None)
        assert same_bits(hits_p.cpu().numpy(), hits.cpu().numpy()) and torch.equal(cp_p.view(torch.int32), cp.view(torch.int32))
        assert torch.equal(occ_p, occ)
    # a coherent batch (sorted input) takes the unsorted path and gives the same answers
    order = np.lexsort((rays[:, 2], rays[:, 1], rays[:, 0]))
    hits2 = accel.trace_closest(torch.from_numpy(rays[order].copy()).cuda())
    torch.cuda.synchronize()
    assert same_bits(hits2.cpu().numpy(), hits.cpu().numpy()[order])
    accel.close(), scene.close()


def test_gather_inbox_single_process(gpurt, orc, ctx):
    """gpurt_gather_*: two contexts on one GPU play owner and sender.  The sender's ordered batch reaches the owner's array
    through the inbox (slices copied in processing order + storage indices, flags, owner-side scatter on a side stream); a
    small batch goes by direct stores and the owner's receivers are told to skip.  Both rounds == the plain call."""
    import torch
    tris = soup(1_600_000, seed=5, ext=0.01)
    ctx2 = gpurt.Context(0)
    scenes, accels = [], []
    for c in (ctx, ctx2):
        sc = gpurt.Scene(c)
        sc.add_triangles(tris)
        scenes.append(sc), accels.append(gpurt.Accel(sc))
    assert accels[0].info().node_bytes + accels[0].info().tri_bytes > (64 << 20)
    ob = orc.Bvh(tris)
    for n0, n1 in (((1 << 20) + 5, (1 << 20) + 4321), (3000, 70000)):   # ordered batches through the inbox; small ones direct
        n = n0 + n1
        q = orc.gen_random_points(n, 72 + n0, ob.scene_box(), frac=0.2)
        d_q = torch.from_numpy(q).cuda()
        want_cp = torch.cat([accels[0].closest_points(d_q[:n0]), accels[1].closest_points(d_q[n0:])]).clone()
        rays = orc.gen_random_rays(n, 71 + n0, ob.scene_box())
        d_rays = torch.from_numpy(rays).cuda()
        want_hit = torch.cat([accels[0].trace_closest(d_rays[:n0]), accels[1].trace_closest(d_rays[n0:])]).clone()
        torch.cuda.synchronize()
        for rb, call, src, want in ((32, "closest_points", d_q, want_cp), (16, "trace_closest", d_rays, want_hit)):
            g0 = gpurt.Gather.create(ctx, n, rb, [0, n0, n])
            g1 = gpurt.Gather.open(ctx2, n, rb, [0, n0, n], 1, base=g0.base())
            for _ in range(3):                                  # rounds reuse the inbox: flags, acknowledgements
                g0.tensor().zero_()
                torch.cuda.synchronize()
                g0.begin()
                getattr(accels[1], call)(src[n0:], g1.mine())
                getattr(accels[0], call)(src[:n0], g0.mine())
                assert g0.end(sync=True) == 0, "a flag wait timed out"
                torch.cuda.synchronize()
                got = g0.tensor().view(torch.int32).view(n, rb // 4)
                ref = want.view(torch.int32).view(n, rb // 4)
                assert torch.equal(got, ref), f"{call}: gathered array differs ({n0}+{n1})"
            g1.close(), g0.close()
    for a in accels:
        a.close()
    for sc in scenes:
        sc.close()
    ctx2.close()


def test_sah_optimal_collapse_flag(gpurt, orc, ctx):
    """GPURT_BUILD_SAH_COLLAPSE: same primitive order and query results as the default build, fewer wide nodes"""
    scene = load_scene(gpurt, ctx, "sponza_standin")
    tris = world_tris(orc, scene)
    base = gpurt.Accel(scene)
    sah = gpurt.Accel(scene, gpurt.BUILD_SAH_COLLAPSE)
    ob = _check_build(gpurt, orc, sah, tris)
    assert sah.info().n_wide_nodes < 0.9 * base.info().n_wide_nodes
    hits = _check_queries(gpurt, orc, sah, ob, tris, 1 << 18, 1 << 10)
    rays = orc.gen_random_rays(1 << 18, 0xC0FFEE, ob.scene_box())
    assert same_bits(hits, base.trace_closest(rays))
    scene.set_transform(5, np.eye(4, dtype=np.float32).reshape(16) * np.float32(1.0))
    sah.update()                                       # the in-place rebuild keeps the flag
    assert sah.info().n_wide_nodes < 0.9 * base.info().n_wide_nodes
    sah.close(), base.close(), scene.close()


def test_sah_split_build_flag(gpurt, orc, ctx):
    """GPURT_BUILD_SAH_SPLIT: host-side binned-SAH order + topology, device refit / collapse / traversal: same query
    results as the default build and the oracle, also combined with the SAH-optimal collapse and after a pose edit"""
    scene = load_scene(gpurt, ctx, "cbox")
    tris = world_tris(orc, scene)
    ob = orc.Bvh(tris)
    rays = orc.gen_random_rays(1 << 18, 41, ob.scene_box())
    q = orc.gen_random_points(1 << 16, 42, ob.scene_box())
    ref_h, ref_a, ref_c = ob.closest_hit(rays), ob.any_hit(rays), ob.closest_point(q)
    base = gpurt.Accel(scene)
    for flags in (gpurt.BUILD_SAH_SPLIT, gpurt.BUILD_SAH_SPLIT | gpurt.BUILD_SAH_COLLAPSE):
        accel = gpurt.Accel(scene, flags)
        assert sorted(accel.prim_order().tolist()) == list(range(len(tris)))
        assert same_bits(accel.trace_closest(rays), ref_h) and (accel.trace_any(rays) == ref_a).all()
        cp = accel.closest_points(q)
        assert same_bits(cp["dist"], ref_c["dist"]) and (cp["prim"] == ref_c["gid"]).all()
        print(f"flags {flags}: {accel.info().n_wide_nodes} wide nodes (default {base.info().n_wide_nodes}), "
              f"build {accel.info().build_ms:.2f} ms")
        accel.close()
    accel = gpurt.Accel(scene, gpurt.BUILD_SAH_SPLIT)
    m = np.array(list(scene.descs()[3].model), np.float32).reshape(4, 4).T.copy()
    m[:3, 3] += np.float32(0.2)
    scene.set_transform(3, m.T.reshape(16).copy())
    accel.update()
    ob2 = orc.Bvh(world_tris(orc, scene))
    assert same_bits(accel.trace_closest(rays), ob2.closest_hit(rays))
    cam = gpurt.camera(0, 128, 128)
    prm = gpurt.pipe_params(max_frames=1, samples_per_frame=1, max_depth=3, integrator=2, seed=3)
    base.update()
    imgs = []
    for a in (accel, base):
        pipe = gpurt.RTPipe(scene, a)
        pipe.render_frame(prm, cam, 128, 128)
        imgs.append(pipe.read_image().copy())
        pipe.close()
    assert same_bits(imgs[0], imgs[1])
    accel.close(), base.close(), scene.close()


def test_refit_keeps_order_and_topology_and_answers_like_a_fresh_build(gpurt, orc, ctx):
    """gpurt_accel_refit / gpurt_accel_update_auto (the reference's rebuild_tlas-only case, src/gpurt.cpp:228-237): after a
    pose edit the primitive order and binary topology stay, boxes and wide nodes follow the new poses, and every query
    answers like the oracle on the moved geometry; a far move makes update_auto rebuild; geometry edits are refused"""
    scene = load_scene(gpurt, ctx, "cbox")
    accel = gpurt.Accel(scene, gpurt.BUILD_KEEP_BVH2)
    order0 = accel.prim_order().copy()
    l0, r0, _ = accel.bvh2()
    full_ms = accel.info().build_ms

    def moved(obj, d):
        m = np.array(list(scene.descs()[obj].model), np.float32)
        m[12:15] += np.asarray(d, np.float32)
        return m
    scene.set_transform(3, moved(3, (0.2, 0.0, -0.1)))
    scene.set_transform(0, moved(0, (-0.1, 0.15, 0.0)))
    accel.refit()
    info = accel.info()
    assert info.refits == 1 and info.tree_cost > 0 and info.tree_cost_at_build > 0
    assert (accel.prim_order() == order0).all()
    l1, r1, b1 = accel.bvh2()
    assert (l1 == l0).all() and (r1 == r0).all()
    tris = world_tris(orc, scene)
    ob = orc.Bvh(tris, sah=True)                     # a fresh tree over the moved geometry: different order, same answers
    assert not (ob.prim_order() == order0).all()
    _check_queries(gpurt, orc, accel, ob, tris, 1 << 17, 1 << 10)
    pipe = gpurt.RTPipe(scene, accel)
    cam, prm = gpurt.camera(0, 128, 72), gpurt.pipe_params(integrator=2, brdf=1, samples_per_frame=2, max_depth=3, seed=4)
    pipe.render_frame(prm, cam, 128, 72)
    fresh = gpurt.Accel(scene)
    fpipe = gpurt.RTPipe(scene, fresh)
    fpipe.render_frame(prm, cam, 128, 72)
    assert same_bits(pipe.read_image(), fpipe.read_image())
    print(f"cbox: full build {full_ms:.3f} ms, refit {info.build_ms:.3f} ms, cost x{info.tree_cost / info.tree_cost_at_build:.3f}")
    # update_auto: a small move refits, a far one (the tree's cost grows) rebuilds
    scene.set_transform(3, moved(3, (0.05, 0.0, 0.0)))
    accel.update_auto()
    assert accel.info().refits == 2 and (accel.prim_order() == order0).all()
    scene.set_transform(0, moved(0, (40.0, 25.0, -30.0)))
    accel.update_auto()
    assert accel.info().refits == 0 and accel.info().tree_cost == accel.info().tree_cost_at_build
    tris = world_tris(orc, scene)
    _check_build(gpurt, orc, accel, tris)
    # geometry edits need a full update
    scene.add_triangles(_soup(10, 1, 0.1))
    with pytest.raises(gpurt.GpurtError):
        accel.refit()
    accel.update_auto()
    assert accel.info().n_tris == len(tris) + 10
    for o in (fpipe, pipe, fresh, accel, scene):
        o.close()


def _soup(n, seed, ext):
    rng = np.random.default_rng(seed)
    c = rng.random((n, 1, 3), dtype=np.float32)
    return (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * np.float32(ext)).reshape(n, 9).astype(np.float32)


def test_sah_split_device_build_equals_the_host_definition(gpurt, orc, ctx, monkeypatch):
    """sah_build.cu against host/sah_split.h (the definition): primitive order and binary topology bit for bit — sizes around
    the 1024-item regime boundary, duplicates (middle splits), clustered input (deep, unbalanced splits), the
    adversarial scene, the shipped scenes — and the queries on the device-built tree against the oracle"""
    from scenes import adversarial_scene
    cases = {f"soup{n}": _soup(n, n + 1, 0.05) for n in (2, 3, 5, 33, 1000, 1024, 1025, 2049, 5000, 70000)}
    dup = _soup(4000, 9, 0.02)
    dup[100:1500] = dup[100]                                    # 1400 coincident triangles: middle splits across regimes
    cases["duplicates"] = dup
    cl = _soup(30000, 3, 0.001)
    cl[:, :] = (cl.reshape(-1, 3, 3) ** 3).reshape(-1, 9)        # strongly clustered towards the origin
    cases["clustered"] = cl
    cases["adversarial"] = np.ascontiguousarray(adversarial_scene(), np.float32)
    for name in ("cbox", "mis_test", "sponza_standin"):
        cases[name] = None
    for name, tris in cases.items():
        if tris is None:
            scene = load_scene(gpurt, ctx, name)
        else:
            scene = gpurt.Scene(ctx)
            scene.add_triangles(tris)
        monkeypatch.setenv("GPURT_SAH_HOST", "1")
        host = gpurt.Accel(scene, gpurt.BUILD_SAH_SPLIT | gpurt.BUILD_KEEP_BVH2)
        monkeypatch.setenv("GPURT_SAH_HOST", "0")
        dev = gpurt.Accel(scene, gpurt.BUILD_SAH_SPLIT | gpurt.BUILD_KEEP_BVH2)
        assert (dev.prim_order() == host.prim_order()).all(), f"{name}: primitive order differs from host/sah_split.h"
        hl, hr, hb = host.bvh2()
        dl, dr, db = dev.bvh2()
        assert (dl == hl).all() and (dr == hr).all(), f"{name}: binary topology differs from host/sah_split.h"
        assert same_bits(db, hb), f"{name}: node boxes differ"
        assert dev.info().n_wide_nodes == host.info().n_wide_nodes and dev.info().wide_depth == host.info().wide_depth
        if name in ("cbox", "soup5000", "duplicates"):
            t = world_tris(orc, scene)
            ob = orc.Bvh(t)
            rays = orc.gen_random_rays(1 << 16, 5, ob.scene_box())
            assert same_bits(dev.trace_closest(rays), ob.closest_hit(rays))
            q = orc.gen_random_points(1 << 14, 6, ob.scene_box())
            cp, ref = dev.closest_points(q), ob.closest_point(q)
            assert same_bits(cp["dist"], ref["dist"]) and (cp["prim"] == ref["gid"]).all()
        print(f"{name}: {dev.info().n_tris} triangles, device SAH build {dev.info().build_ms:.3f} ms, host-side {host.info().build_ms:.3f} ms")
        dev.close(), host.close(), scene.close()


def test_no_device_memory_growth_over_create_destroy_cycles(gpurt, ctx):
    """scene / accel / pipe / update cycles return their device memory (cudaMemGetInfo stays flat)"""
    import torch
    w, h = 128, 72
    cam = gpurt.camera(0, w, h)
    prm = gpurt.pipe_params(max_frames=1, samples_per_frame=1, max_depth=2, integrator=2)

    def cycle():
        scene = load_scene(gpurt, ctx, "cbox")
        accel = gpurt.Accel(scene)
        pipe = gpurt.RTPipe(scene, accel)
        pipe.render_frame(prm, cam, w, h)
        scene.set_transform(2, np.eye(4, dtype=np.float32).reshape(16))
        accel.update()
        pipe.reset_frame()
        pipe.render_frame(prm, cam, w, h)
        pipe.read_image()
        pipe.close(), accel.close(), scene.close()

    for _ in range(5):
        cycle()                      # pools and arenas reach their steady size
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(40):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (8 << 20), f"device memory shrank by {(free0 - free1) >> 20} MiB over 40 cycles"

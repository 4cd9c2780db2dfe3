#!/usr/bin/env python
"""Host-buffer (end-to-end) path probe: plain pinned H2D / D2H bandwidth of this box and trace_closest from
pinned host buffers for several pipeline chunk sizes (GPURT_HOST_CHUNK)."""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-rt_b200"))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1:
    import bench
    import gpurt
    ctx = gpurt.Context(0)
    scene, _ = bench.build_scene(gpurt, ctx)
    accel = gpurt.Accel(scene)
    pipe = gpurt.RTPipe(scene, accel)
    prm = gpurt.pipe_params(integrator=1, brdf=1, max_depth=2, samples_per_frame=1, max_frames=1, use_rr=0, env_scale=1.0)
    ctx.use_torch_stream()
    pipe.render_frame(prm, gpurt.camera(1, bench.W, bench.H, bench.CAM_POS, bench.CAM_AT, bench.VFOV), bench.W, bench.H)
    rays = torch.cat([pipe.bounce_rays(0), pipe.bounce_rays(1)]).cpu().pin_memory()
    hits = torch.empty((rays.shape[0], 4), dtype=torch.float32).pin_memory()
    hr, hh = rays.numpy(), hits.numpy().view(gpurt.HIT_DT).reshape(-1)
    if os.environ.get("E2E_WC") == "1":   # rays in write-combined pinned memory
        import ctypes as C
        rt = C.CDLL("libcudart.so.12")
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(hr.nbytes), C.c_uint(4)) == 0
        wc = np.frombuffer((C.c_uint8 * hr.nbytes).from_address(p.value), np.float32).reshape(hr.shape)
        wc[:] = hr
        hr = wc
    for _ in range(3):
        accel.trace_closest(hr, hh)
    t0 = time.time()
    for _ in range(10):
        accel.trace_closest(hr, hh)
    dt = (time.time() - t0) / 10
    print(f"chunk {os.environ.get('GPURT_HOST_CHUNK', 'default'):>8s} zero_copy {os.environ.get('GPURT_ZERO_COPY', '1')} wc {os.environ.get('E2E_WC', '0')}: "
          f"{dt * 1e3:6.3f} ms  {rays.shape[0] / dt / 1e6:7.1f} Mrays/s")
else:
    n = 112 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        print(f"{name} pinned 112 MB: {n * 10 / (time.time() - t0) / 1e9:.1f} GB/s")
    for c in ("32768", "65536", "131072", "262144", "524288", "1048576"):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, GPURT_HOST_CHUNK=c))

import ctypes as C, time, sys, os
import numpy as np, torch
rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
torch.cuda.init(); torch.zeros(1, device="cuda")
def host_alloc(n, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    return p
n = 112 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
hout = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, flags in (("default", 0), ("write-combined", 4)):
    p = host_alloc(n, flags)
    buf = (C.c_uint8 * n).from_address(p.value)
    np.frombuffer(buf, np.uint8)[:] = 1
    def h2d(stream=None):
        rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(stream.cuda_stream if stream else 0))
    for _ in range(3): h2d()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(10): h2d()
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"H2D {name}: {n * 10 / dt / 1e9:.1f} GB/s")
    # bidirectional: H2D of n on s1, D2H of n/2 on s2 concurrently
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(10):
        h2d(s1)
        with torch.cuda.stream(s2): hout.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"  with concurrent D2H of half the bytes: H2D {n * 10 / dt / 1e9:.1f} GB/s (wall for both)")

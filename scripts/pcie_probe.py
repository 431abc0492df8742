"""PCIe / host-copy probe on the GPU box: how fast can 1.2 GB of samples reach a caller's buffer?
(a) D2H into pinned memory, (b) cudaHostRegister cost of a pageable buffer, (c) pinned->pageable memcpy with T threads."""
import ctypes, sys, time, threading
import numpy as np, torch
N = 151552 * 100 * 10
d = torch.randn(N, dtype=torch.float64, device="cuda")
pin = torch.empty(N, dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
for _ in range(3):
    t = time.perf_counter(); pin.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("D2H pinned 1.2GB: %.1f ms  %.1f GB/s" % (dt * 1e3, N * 8 / dt / 1e9), flush=True)
page = np.empty(N)
page[:] = 0
rt = torch.cuda.cudart()
for k in range(2):
    t = time.perf_counter(); rc = rt.cudaHostRegister(page.ctypes.data, N * 8, 0); dt = time.perf_counter() - t
    print("cudaHostRegister 1.2GB: rc %s %.1f ms" % (rc, dt * 1e3), flush=True)
    tp = torch.from_numpy(page)
    t = time.perf_counter(); tp.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt2 = time.perf_counter() - t
    print("  D2H into registered: %.1f ms %.1f GB/s" % (dt2 * 1e3, N * 8 / dt2 / 1e9), flush=True)
    t = time.perf_counter(); rt.cudaHostUnregister(page.ctypes.data); dt3 = time.perf_counter() - t
    print("  unregister %.1f ms" % (dt3 * 1e3), flush=True)
t = time.perf_counter(); tp = torch.from_numpy(page); tp.copy_(d); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("D2H pageable (driver staged) %.1f ms %.1f GB/s" % (dt * 1e3, N * 8 / dt / 1e9), flush=True)
libc = ctypes.CDLL("libc.so.6")
src = pin.numpy()
for T in (1, 2, 4, 8, 16, 32):
    part = (N // T + 511) & ~511
    def work(i):
        o = i * part
        l = min(part, N - o)
        if l > 0:
            libc.memcpy(ctypes.c_void_p(page.ctypes.data + o * 8), ctypes.c_void_p(src.ctypes.data + o * 8), ctypes.c_size_t(l * 8))
    best = 1e9
    for _ in range(3):
        th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
        t = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]; best = min(best, time.perf_counter() - t)
    print("memcpy pinned->pageable T=%d: %.1f ms %.1f GB/s" % (T, best * 1e3, N * 8 / best / 1e9), flush=True)

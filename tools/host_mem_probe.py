"""Dev probe: first-touch and copy rates of host memory on the GPU box (decides how invert_from_model stages its results)."""
import ctypes, time, threading, sys
import numpy as np
import torch

def t(f):
    t0 = time.perf_counter(); r = f(); return time.perf_counter() - t0, r

n = 1 << 30
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
src = np.ones(n, dtype=np.uint8)
dt, a = t(lambda: np.empty(4 * n, dtype=np.uint8))
dt2, _ = t(lambda: a.__setitem__(slice(None), 1))
print(f"np.empty 4 GiB {dt*1e3:.1f} ms; first touch (fill) {4/dt2:.2f} GiB/s")
dt2, _ = t(lambda: a.__setitem__(slice(None), 2))
print(f"second touch {4/dt2:.2f} GiB/s")
b = np.empty(4 * n, dtype=np.uint8)
libc = ctypes.CDLL("libc.so.6", use_errno=True)
addr = b.ctypes.data & ~0xFFF
rc = libc.madvise(ctypes.c_void_p(addr), ctypes.c_size_t(4 * n), 14)
dt2, _ = t(lambda: b.__setitem__(slice(None), 1))
print(f"madvise(HUGEPAGE) rc={rc}; first touch {4/dt2:.2f} GiB/s")
c = np.empty(4 * n, dtype=np.uint8)
def part(k, m):
    lo = k * (4 * n // m); c[lo: lo + 4 * n // m] = 1
for m in (2, 4):
    c = np.empty(4 * n, dtype=np.uint8)
    ths = [threading.Thread(target=part, args=(k, m)) for k in range(m)]
    t0 = time.perf_counter(); [x.start() for x in ths]; [x.join() for x in ths]; dt2 = time.perf_counter() - t0
    print(f"first touch with {m} threads {4/dt2:.2f} GiB/s")
dt2, _ = t(lambda: np.copyto(a[:n], src))
print(f"memcpy 1 GiB warm {1/dt2:.2f} GiB/s")
dt, p = t(lambda: torch.empty(n, dtype=torch.uint8, pin_memory=True))
print(f"pinned alloc 1 GiB {dt*1e3:.0f} ms")
d = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
dt, _ = t(lambda: (p.copy_(d, non_blocking=True), torch.cuda.synchronize()))
print(f"D2H pinned {1/dt:.1f} GiB/s")
dt, _ = t(lambda: np.copyto(a[n:2*n], p.numpy()))
print(f"pinned -> pageable memcpy {1/dt:.2f} GiB/s")
import os
print("cpus", os.cpu_count())

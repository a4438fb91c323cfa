"""Dev measurement: invert_from_model end to end with pageable vs pinned host arrays."""
import sys, time, warnings
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from xsarsea_b200 import windspeed

lines = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
inc, s_co, s_cr, anc = bench.synth_scene_device(lines, 25000, 0)
pinned = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in (inc, s_co, s_cr, anc)]
pageable = [p.numpy().copy() for p in pinned]
del inc, s_co, s_cr, anc
torch.cuda.empty_cache()
model = ("gmf_cmod5n", "gmf_s1_v2")
def run(arrs):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        co, du = windspeed.invert_from_model(arrs[0], arrs[1], arrs[2], ancillary_wind=arrs[3], dsig_cr=0.1, model=model)
        x = float(np.nanmean(np.abs(du[0])))
        return time.perf_counter() - t0
t_first = run([p.numpy() for p in pinned])
print('first call (plan + LUTs + page-locking the staging buffers) s', round(t_first, 2))
n = lines * 25000
for name, arrs in (("pinned", [p.numpy() for p in pinned]), ("pageable", pageable)):
    ts = [run(arrs) for _ in range(3)]
    print(name, "Mpx/s", [round(n / t / 1e6, 1) for t in ts])

"""Kernel-level timing of the cross-pol pass (k_cross): cross-pol-only and the dual-pol tail (dev tool)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from xsarsea_b200 import _device as D
from xsarsea_b200 import _native as nat

H = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
gi, gwc = np.linspace(16, 66, 501), np.linspace(3, 80, 771)
cr = D.lut_to_db(D.lut_build(nat.GMF_IDS["gmf_s1_v2"], gi, gwc, None))
plan = D.InversionPlan(cr=(cr, gi, gwc))
inc, s_co, s_cr, anc = bench.synth_scene_device(H, W, 3)
o = torch.empty(inc.shape, dtype=torch.float64, device="cuda")
dsig = torch.full_like(inc, 0.1)
for name, d in (("scalar dsig", 0.1), ("raster dsig", dsig)):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.invert(inc, None, s_cr, d, None, cr_abs=True, out_cr=o)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(json.dumps(dict(case="cross-only " + name, H=H, W=W, ms=min(ts), Gpx_s=H * W / min(ts) / 1e6)), flush=True)
print("mean wspd", torch.nanmean(o).item())

"""Kernel-level timing of xs_invert on a synthetic IW-like strip (dev tool; bench.py is the contract)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from xsarsea_b200 import _device as D
from xsarsea_b200 import _native as nat

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
gi = np.linspace(16, 66, 501)
gw = np.linspace(0.2, 50, 499)
gp = np.linspace(0, 180, 181)
gwc = np.linspace(3, 80, 771)
t0 = time.time()
co = D.lut_to_db(D.lut_build(nat.GMF_IDS["gmf_cmod5n"], gi, gw, gp))
cr = D.lut_to_db(D.lut_build(nat.GMF_IDS["gmf_s1_v2"], gi, gwc, None))
torch.cuda.synchronize()
t1 = time.time()
plan = D.InversionPlan(co=(co, gi, gw, gp), cr=(cr, gi, gwc))
torch.cuda.synchronize()
t2 = time.time()
print(f"lut build {t1-t0:.3f}s plan {t2-t1:.3f}s", flush=True)
g = torch.Generator(device="cuda").manual_seed(0)
inc = (30 + 16 * torch.arange(W, device="cuda", dtype=torch.float64) / (W - 1)).expand(H, W).contiguous()
wspd = 2 + 23 * torch.rand(H, W, generator=g, device="cuda", dtype=torch.float64)
phi = 360 * torch.rand(H, W, generator=g, device="cuda", dtype=torch.float64)
s_co = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], inc, wspd, phi) * torch.exp(0.05 * torch.randn(H, W, generator=g, device="cuda", dtype=torch.float64))
s_cr = D.gmf_eval(nat.GMF_IDS["gmf_s1_v2"], inc, wspd, None) * torch.exp(0.05 * torch.randn(H, W, generator=g, device="cuda", dtype=torch.float64))
anc = torch.polar(wspd + 2 * torch.randn(H, W, generator=g, device="cuda", dtype=torch.float64),
                  torch.deg2rad(phi + 20 * torch.randn(H, W, generator=g, device="cuda", dtype=torch.float64)))
for dual in (False, True):
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        oc, ox, _, _ = plan.invert(inc, s_co, s_cr if dual else None, 0.1, anc, merge_dual=dual, mode=mode, timed=(mode == 0))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = plan.last_stats()
        if mode == 0:
            st['scan_ms'], st['refine_ms'] = plan.last_scan_ms()
        print(json.dumps(dict(dual=dual, H=H, W=W, ms=ms, Mpx_s=H * W / ms / 1e3, **st,
                              frac_fp32_peak=H * W * (727178 if dual else 722552) / (ms * 1e-3) / 74.4e12)), flush=True)
print("debug counters", plan.debug_counters())
print("mean |co|", torch.nanmean(oc.abs()).item(), "launches", nat.launch_count())

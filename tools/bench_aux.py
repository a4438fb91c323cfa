"""Secondary measurements of SURVEY.md section 8 D4 (not the headline): LUT generation, sigma0_detrend, cross-pol-only
and co-pol-only inversion.  Prints one JSON object; CUDA-event timing, 3 warm-ups, best-of-5 and mean."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xsarsea_b200 import _device as D
from xsarsea_b200 import _native as nat

HBM_GBS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, warm=3, reps=5, inner=1):
    """best / mean over `reps` timings of `inner` back-to-back calls (inner > 1 for sub-millisecond operations, so
    that launch latency is not what is measured)."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return float(np.min(ts)), float(np.mean(ts))


out = {}
# ---- LUT generation (config 2): direct high-res evaluation and the default low-res + interpolation path ----
gi, gw, gp = np.linspace(16, 66, 501), np.linspace(0.2, 50, 499), np.linspace(0, 180, 181)
best, mean = timeit(lambda: D.lut_build(nat.GMF_IDS["gmf_cmod5n"], gi, gw, gp), warm=5, reps=7, inner=5)
out["lut_build_cmod5n_high_501x499x181"] = dict(ms=best, ms_mean=mean, Gevals_per_s=gi.size * gw.size * gp.size / best / 1e6,
                                                 reference_cpu_s=12.6)
li, lw, lp = np.linspace(16, 66, 51), np.linspace(0.2, 50, 250), np.linspace(0, 180, 73)


def default_path():
    lut = D.lut_build(nat.GMF_IDS["gmf_cmod5n"], li, lw, lp)
    lut = D.lut_interp_axis(lut, 0, li, gi)
    lut = D.lut_interp_axis(lut, 1, lw, gw)
    lut = D.lut_interp_axis(lut, 2, lp, gp)
    return D.lut_to_db(lut)


best, mean = timeit(default_path, warm=5, reps=7, inner=5)
out["to_lut_default_path_cmod5n_dB"] = dict(ms=best, ms_mean=mean, reference_cpu_s=6.2)
co = default_path()
gwc = np.linspace(3, 80, 771)
cr = D.lut_to_db(D.lut_build(nat.GMF_IDS["gmf_s1_v2"], gi, gwc, None))
best, mean = timeit(lambda: D.InversionPlan(co=(co, gi, gw, gp), cr=(cr, gi, gwc)).close())
out["plan_create_scan_image"] = dict(ms=best, ms_mean=mean)

# ---- sigma0_detrend (HBM-bound, 16 B/px f64) on an EW-sized raster (config 5) ----
H, W = 10000, 10400
s0 = torch.rand(H, W, dtype=torch.float64, device="cuda") * 0.2 + 0.01
prof = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], torch.linspace(19, 47, W, dtype=torch.float64, device="cuda"),
                  torch.full((W,), 10.0, dtype=torch.float64, device="cuda"), torch.full((W,), 45.0, dtype=torch.float64, device="cuda"))
best, mean = timeit(lambda: D.detrend(s0, prof), warm=10, reps=7, inner=20)
out["detrend_f64_10000x10400"] = dict(ms=best, ms_mean=mean, GBps=16 * H * W / best / 1e6, frac_of_measured_hbm=16 * H * W / best / 1e6 / HBM_GBS,
                                      Gpx_per_s=H * W / best / 1e6)
s32 = s0.float()
best, mean = timeit(lambda: D.detrend(s32, prof), warm=10, reps=7, inner=20)
out["detrend_f32_10000x10400"] = dict(ms=best, ms_mean=mean, GBps=8 * H * W / best / 1e6, frac_of_measured_hbm=8 * H * W / best / 1e6 / HBM_GBS)
del s0, s32

# ---- dsig_cr pre-processors (row F1; HBM-bound FP64 element-wise / per-line fit) on the same EW-sized raster ----
inc2 = torch.linspace(19, 47, W, dtype=torch.float64, device="cuda").expand(H, W).contiguous()
scr = 10 ** (torch.rand(H, W, dtype=torch.float64, device="cuda") * 2.5 - 3.7)
nesz = 10 ** (-3.2 + 0.05 * torch.randn(H, W, dtype=torch.float64, device="cuda"))
for name, did, bpp in (("gmf_s1_v2", 0, 32), ("gmf_rs2_v2", 1, 24), ("nc_lut_cmodms1ahw", 2, 24)):
    best, mean = timeit(lambda: D.dsig(did, inc2 if did == 0 else None, scr, nesz), warm=5, reps=7, inner=10)
    out[f"get_dsig_{name}_10000x10400"] = dict(ms=best, ms_mean=mean, algorithmic_B_per_px=bpp, GBps=bpp * H * W / best / 1e6,
                                               frac_of_measured_hbm=bpp * H * W / best / 1e6 / HBM_GBS)
best, mean = timeit(lambda: D.dsig_wspd(1, scr * 1e3, nesz * 1e4), warm=5, reps=7, inner=10)
out["get_dsig_wspd_s1_ew_rec_v3_10000x10400"] = dict(ms=best, ms_mean=mean, algorithmic_B_per_px=24, GBps=24 * H * W / best / 1e6,
                                                      frac_of_measured_hbm=24 * H * W / best / 1e6 / HBM_GBS)
best, mean = timeit(lambda: D.nesz_flatten(nesz, inc2), warm=5, reps=7, inner=5)
# column means read noise + incidence (16 B/px), the fit reads the noise again and writes the result (16 B/px)
out["nesz_flattening_10000x10400"] = dict(ms=best, ms_mean=mean, algorithmic_B_per_px=32, GBps=32 * H * W / best / 1e6,
                                          frac_of_measured_hbm=32 * H * W / best / 1e6 / HBM_GBS)
del inc2, scr, nesz

# ---- local_gradients (row F4): image read once (8 B/px), three half-size outputs (32 B per 4 px) ----
img = 0.1 + 0.02 * torch.rand(H, W, dtype=torch.float64, device="cuda")
best, mean = timeit(lambda: D.local_gradients(img), warm=5, reps=7, inner=5)
out["local_gradients_10000x10400"] = dict(ms=best, ms_mean=mean, algorithmic_B_per_px=16, GBps=16 * H * W / best / 1e6,
                                          frac_of_measured_hbm=16 * H * W / best / 1e6 / HBM_GBS, Gpx_per_s=H * W / best / 1e6)
del img

# ---- inversion: cross-pol only (config 4, 10000 x 10000) and co-pol only (config 1, 1000 x 1000) ----
plan_x = D.InversionPlan(cr=(cr, gi, gwc))
inc, s_co, s_cr, anc = bench.synth_scene_device(10000, 10000, 3)
o = torch.empty(inc.shape, dtype=torch.float64, device="cuda")
best, mean = timeit(lambda: plan_x.invert(inc, None, s_cr, 0.1, None, cr_abs=True, out_cr=o))
n = inc.numel()
out["cross_pol_only_10000x10000"] = dict(ms=best, ms_mean=mean, Mpx_per_s=n / best / 1e3, hbm_GBps=24 * n / best / 1e6,
                                         fp32_equiv_TFLOPs=4626 * n / best / 1e9, reference_cpu_Mpx_per_s=1.6)
del inc, s_co, s_cr, anc, o
plan_c = D.InversionPlan(co=(co, gi, gw, gp))
H = W = 1000
g = torch.Generator(device="cuda").manual_seed(0)
f64 = dict(device="cuda", dtype=torch.float64)
inc = (17.5 + 32 * torch.arange(W, **f64) / (W - 1)).expand(H, W).contiguous()
w = 2 + 23 * torch.rand(H, W, generator=g, **f64)
p = 360 * torch.rand(H, W, generator=g, **f64)
s_co = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], inc, w, p) * torch.exp(0.05 * torch.randn(H, W, generator=g, **f64))
anc = torch.polar((w + 2 * torch.randn(H, W, generator=g, **f64)).abs(), torch.deg2rad(p + 20 * torch.randn(H, W, generator=g, **f64)))
oc = torch.empty_like(anc)
ox = torch.empty_like(anc)
best, mean = timeit(lambda: plan_c.invert(inc, s_co, None, 0.1, anc, out_co=oc, out_cr=ox))
out["co_pol_only_1000x1000_config1"] = dict(ms=best, ms_mean=mean, Mpx_per_s=H * W / best / 1e3, frac_fp32_peak=722552 * H * W / best / 1e9 / 74.45,
                                            reference_cpu_kpx_per_s=6.3, stats=plan_c.last_stats())
print(json.dumps(out, indent=1))

#!/usr/bin/env python
"""bench.py -- inverted pixels/s of the dual-pol (gmf_cmod5n + nc_lut_cmodms1ahw) wind inversion.

A "step" is one full inversion of a synthetic Sentinel-1 IW GRD-sized dual-pol scene (16700 lines x 25000 samples,
incidence 30..46 deg) on every rank: `value` times the C-ABI call `xs_invert` with the rasters resident in HBM;
`e2e` times the public API `xsarsea_b200.windspeed.invert_from_model` with pinned HOST numpy arrays (host->device
and device->host copies inside the timed region).  With N > 1 ranks every GPU inverts its own scene (the batch-of-
scenes sharding of BASELINE.json configs[4]; no data-path collective), so scaling is "weak".

`--impl reference` times the reference's CPU path instead (the in-repo numba port of windspeed.py:183-323 with the
reference's own guvectorize arguments -- xarray/dask are not installable here) on a bounded sample
of the same workload.

`nc_lut_cmodms1ahw` is a synthetic stand-in: the real LUT file is not part of the reference repository; a NetCDF-3
file in the reference's schema (331 x 771, inc 17..50, wspd 3..80, dB) is generated from 10*log10(gmf_s1_v2).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LINES, SAMPLES = 16700, 25000
INC_NEAR, INC_FAR = 30.0, 46.0
FLOP_PER_PX = 8 * 499 * 181 + 6 * 771  # 8 flop per co-pol candidate + 6 per cross-pol candidate
METRIC = "inverted pixels/sec (dual-pol cmod5n+ms1ahw)"


def workload_text(args):
    """The workload both arms of the bench report (config.workload)."""
    return (f"dual-pol invert: gmf_cmod5n (default LUT 501x499x181) + nc_lut_cmodms1ahw (synthetic stand-in 331x771), "
            f"S1 IW {args.lines}x{args.samples} px per GPU, inc {INC_NEAR}-{INC_FAR} deg, 1% NaN, dsig_cr=0.1, ancillary wind")


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_mhz = float(p.get("sm_max_mhz", 1965.0))
    return dict(fp32_tflops=148 * 128 * 2 * sm_mhz * 1e6 / 1e12, sm_max_mhz=sm_mhz,
                source="148 SM x 128 lanes x 2 x sm_max_mhz of MEASURED_PEAKS.json" if p else
                "148 SM x 128 lanes x 2 x 1965 MHz (fallback: MEASURED_PEAKS.json absent)")


def profiled_traffic_per_px():
    """DRAM bytes per pixel of k_scan_co from the committed `ncu --set full` capture (profiles/): that capture was taken
    on `bench.py --lines 400` (10 Mpx per launch); traffic is proportional to the pixel count (rasters in, results out,
    4 B/px of pixel list), so it is reported scaled to this run's launch size."""
    path = os.path.join(ROOT, "profiles", "r1_k_scan_co_ncu_raw_selected.csv")
    try:
        vals = {}
        for line in open(path):
            f = line.strip().split(",")
            if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[1]]
                vals[f[0]] = float(f[2]) * scale
        return (vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]) / (400 * 25000)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                if float(r[2]) > 300:  # under load
                    sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples_under_load=len(sm))


def write_ms1ahw_standin(dirname):
    """nc_lut_cmodms1ahw.nc in the reference's schema (models.py:232-262, 368-379), NetCDF-3 via scipy."""
    from scipy.io import netcdf_file

    from xsarsea_b200 import _device as D
    from xsarsea_b200 import _native as nat

    inc = np.linspace(17.0, 50.0, 331)
    wspd = np.linspace(3.0, 80.0, 771)
    # 10*log10(gmf_s1_v2 + 1e-15), evaluated with the package's own device GMF (the oracle stays out of this arm)
    lut_db = D.lut_to_db(D.lut_build(nat.GMF_IDS["gmf_s1_v2"], inc, wspd, None)).cpu().numpy()
    path = os.path.join(dirname, "nc_lut_cmodms1ahw.nc")
    with netcdf_file(path, "w") as nc:
        nc.units, nc.pol, nc.model, nc.resolution = "dB", "VH", "cmodms1ahw", "high"
        nc.inc_range, nc.wspd_range = np.array([17.0, 50.0]), np.array([3.0, 80.0])
        nc.inc_step, nc.wspd_step = 0.1, 0.1
        nc.createDimension("incidence", inc.size)
        nc.createDimension("wspd", wspd.size)
        nc.createVariable("incidence", "d", ("incidence",))[:] = inc
        nc.createVariable("wspd", "d", ("wspd",))[:] = wspd
        nc.createVariable("sigma0_model", "d", ("incidence", "wspd"))[:] = lut_db
    return path


def synth_scene_device(lines, samples, seed):
    """SURVEY.md D2 recipe generated on the device (torch CUDA generator, seed stated in the JSON line)."""
    import torch

    from xsarsea_b200 import _device as D
    from xsarsea_b200 import _native as nat

    g = torch.Generator(device="cuda").manual_seed(seed)
    f64 = dict(device="cuda", dtype=torch.float64)
    inc = (INC_NEAR + (INC_FAR - INC_NEAR) * torch.arange(samples, **f64) / (samples - 1)).expand(lines, samples).contiguous()
    wspd = 2 + 23 * torch.rand(lines, samples, generator=g, **f64)
    phi = 360 * torch.rand(lines, samples, generator=g, **f64)
    s_co = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], inc, wspd, phi)
    s_co *= torch.exp(0.05 * torch.randn(lines, samples, generator=g, **f64))
    s_cr = D.gmf_eval(nat.GMF_IDS["gmf_s1_v2"], inc, wspd, None)
    s_cr *= torch.exp(0.05 * torch.randn(lines, samples, generator=g, **f64))
    wa = wspd + 2 * torch.randn(lines, samples, generator=g, **f64)
    pa = torch.deg2rad(phi + 20 * torch.randn(lines, samples, generator=g, **f64))
    anc = torch.polar(wa.abs(), pa + (wa < 0) * np.pi)
    land = torch.rand(lines, samples, generator=g, device="cuda") < 0.01   # 1 % NaN (land mask)
    s_co[land] = float("nan")
    s_cr[land] = float("nan")
    return inc, s_co, s_cr, anc


def make_cpu_port(threads=None):
    """The reference's CPU program shape (oracle/numba_port.py: the numba gufunc of windspeed.py:183-323 with the
    reference's decorator arguments) for the bench workload.  Returns run(sample) -> (px/s, threads, seconds)."""
    import numba

    import oracle
    from oracle import lut as olut
    from oracle import numba_port

    if threads:
        numba.set_num_threads(threads)
    co_lut, (gi, gw, gp) = olut.to_lut("gmf_cmod5n", units="dB")
    gic, gwc = np.linspace(17.0, 50.0, 331), np.linspace(3.0, 80.0, 771)
    cr_lut = 10 * np.log10(oracle.lut_build("gmf_s1_v2", gic, gwc) + 1e-15)
    f = numba_port.make_inverter(co_lut, gi, gw, gp, cr_lut, gic, gwc)
    tiny = (np.full((2, 8), 35.0), np.full((2, 8), -15.0), np.full((2, 8), -25.0), np.full((2, 8), 0.1),
            np.full((2, 8), 5 + 5j))
    f(*tiny)  # JIT compilation, excluded from every timing

    def run(sample):
        inc, s_co, s_cr, anc = sample
        with np.errstate(all="ignore"):
            co_db, cr_db = 10 * np.log10(s_co + 1e-15), 10 * np.log10(s_cr + 1e-15)   # windspeed.py:126-128
        dsig = np.full(inc.shape, 0.1)
        t0 = time.perf_counter()
        f(inc, co_db, cr_db, dsig, anc)
        dt = time.perf_counter() - t0
        return inc.size / dt, numba.get_num_threads(), dt

    return run


def sized_cpu_lines(port, make_sample, target_s, lo=16, hi=4096):
    """Number of 1000-sample lines that keeps one CPU pass near `target_s` seconds (probe with `lo` lines first)."""
    rate, _, _ = port(make_sample(lo))
    return int(min(hi, max(lo, round(target_s * rate / 1000.0 / 16) * 16)))


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    import oracle  # inputs of the reference arm are made on the host: none of our kernels on this path

    def make_sample(lines):
        rng = np.random.default_rng(args.seed)
        inc = np.broadcast_to(np.linspace(INC_NEAR, INC_FAR, 1000), (lines, 1000)).copy()
        w, p = rng.uniform(2, 25, inc.shape), rng.uniform(0, 360, inc.shape)
        s_co = oracle.gmf_eval("gmf_cmod5n", inc, w, p) * np.exp(rng.normal(0, 0.05, inc.shape))
        s_cr = oracle.gmf_eval("gmf_s1_v2", inc, w) * np.exp(rng.normal(0, 0.05, inc.shape))
        anc = (w + rng.normal(0, 2, inc.shape)) * np.exp(1j * np.deg2rad(p + rng.normal(0, 20, inc.shape)))
        land = rng.uniform(size=inc.shape) < 0.01
        s_co[land] = np.nan
        s_cr[land] = np.nan
        return inc, s_co, s_cr, anc

    port = make_cpu_port()
    lines = args.cpu_lines or sized_cpu_lines(port, make_sample, args.cpu_seconds)
    sample = make_sample(lines)
    rates, secs = [], []
    threads = None
    for it in range(args.warmup + args.steps):
        r, threads, dt = port(sample)
        if it >= args.warmup:
            rates.append(r)
            secs.append(dt)
    v = float(np.mean(rates))
    sample_txt = f"{sample[0].shape[0]}x{sample[0].shape[1]} px crop across the swath of the same synthetic scene per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "px/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args), "sample": sample_txt},
        "cpu_baseline": {"value": v, "unit": "px/s", "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": v, "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lines", type=int, default=LINES)
    ap.add_argument("--samples", type=int, default=SAMPLES)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-lines", type=int, default=0, help="lines of the 1000-sample crop timed on the CPU (0: sized for --cpu-seconds)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of one CPU pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout carries exactly one JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner, for one)
    # goes to stderr instead
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    from xsarsea_b200 import _native as nat
    from xsarsea_b200 import windspeed
    from xsarsea_b200.windspeed import windspeed as ws_impl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- models through the public API: gmf_cmod5n (device GMF) + the synthetic nc_lut_cmodms1ahw file ----
    tmp = tempfile.mkdtemp(prefix=f"xs_bench_{rank}_")
    write_ms1ahw_standin(tmp)
    windspeed.register_nc_luts(tmp)
    model = ("gmf_cmod5n", "nc_lut_cmodms1ahw")
    t0 = time.perf_counter()
    plan = ws_impl._get_plan(windspeed.get_model(model[0]), windspeed.get_model(model[1]), 0.1, {})
    torch.cuda.synchronize()
    lut_s = time.perf_counter() - t0

    inc, s_co, s_cr, anc = synth_scene_device(args.lines, args.samples, args.seed + rank)
    n_px = inc.numel()
    out_co = torch.empty_like(anc)
    out_cr = torch.empty_like(anc)
    torch.cuda.synchronize()

    def step():
        plan.invert(inc, s_co, s_cr, 0.1, anc, merge_dual=True, out_co=out_co, out_cr=out_cr, timed=True)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = nat.launch_count()
    scan_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
        scan_ms.append(plan.last_scan_ms()[0])   # waits for this step's scan + refine kernels only
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = nat.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    stats = plan.last_stats()
    value = world * n_px * args.steps / (ms * 1e-3)
    scan_avg_ms = float(np.mean(scan_ms))
    pk = peaks()
    achieved = FLOP_PER_PX * n_px / (scan_avg_ms * 1e-3) / 1e12
    bpp = profiled_traffic_per_px()

    # ---- end to end through the public API with pinned host arrays ----
    e2e = None
    if not args.no_e2e:
        host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (inc, s_co, s_cr, anc)]
        for h, t in zip(host, (inc, s_co, s_cr, anc)):
            h.copy_(t)
        torch.cuda.synchronize()
        h_inc, h_co, h_cr, h_anc = (h.numpy() for h in host)
        h2d = sum(h.numel() * h.element_size() for h in host)
        d2h = 2 * n_px * 16
        del inc, s_co, s_cr, anc, out_co, out_cr
        torch.cuda.empty_cache()
        import warnings

        def e2e_step():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                co, dual = windspeed.invert_from_model(h_inc, h_co, h_cr, ancillary_wind=h_anc, dsig_cr=0.1, model=model)
            return float(np.nanmean(np.abs(dual[0])))  # read of the result on the host

        e2e_steps = max(1, min(args.steps, 2))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        chk = [e2e_step() for _ in range(e2e_steps)]
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * n_px * e2e_steps / dt, "unit": "px/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "api": "xsarsea_b200.windspeed.invert_from_model(numpy f64/c128, pinned host)",
               "check_mean_abs_dual_line0": chk[-1]}
        cpu_src = (h_inc, h_co, h_cr, h_anc)
    else:
        cpu_src = None

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        step_s = max(1, args.samples // 1000)

        def make_sample(lines):
            lines = min(lines, args.lines)
            sl = (slice(0, lines), slice(0, step_s * 1000, step_s))
            if cpu_src is not None:
                return tuple(np.ascontiguousarray(a[sl]) for a in cpu_src)
            return tuple(t[sl].contiguous().cpu().numpy() for t in (inc, s_co, s_cr, anc))

        port = make_cpu_port()
        lines = args.cpu_lines or sized_cpu_lines(port, make_sample, args.cpu_seconds)
        sample = make_sample(lines)
        r, threads, dt = port(sample)
        cpu = {"value": r, "unit": "px/s", "cores": threads, "kind": "port", "seconds": dt,
               "sample": f"{sample[0].shape[0]}x{sample[0].shape[1]} px crop across the swath of the same scene, numba "
                         f"gufunc with the reference's decorator arguments (oracle/numba_port.py), JIT excluded"}

    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "px/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 scan + f64 refinement (f64/c128 rasters)", "data": f"synthetic (SURVEY D2 recipe, torch CUDA generator, seed {args.seed}+rank)",
            "config": {"workload": workload_text(args),
                       "l2": "inputs (40 B/px x %.1f Mpx = %.1f GB) exceed L2 (126 MB)" % (n_px / 1e6, 40 * n_px / 1e9),
                       "sharding": "one scene per GPU, no data-path collective", "lut_build_s": lut_s},
            "roofline": {"bound": "fp32 cuda-core (FMA pipe)", "kernel": "k_scan_co", "achieved": achieved,
                         "peak": pk["fp32_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["fp32_tflops"],
                         "traffic": None if bpp is None else bpp * n_px,
                         "traffic_note": "DRAM read+write bytes per launch, scaled by pixel count from the ncu --set full "
                                         "capture of a 10 Mpx launch (profiles/r1_k_scan_co_ncu_raw_selected.csv: %s B/px; "
                                         "algorithmic 40 B/px in + 16 B/px out + 4 B/px list)" % (None if bpp is None else round(bpp, 1)),
                         "peak_source": pk["source"], "scan_ms_per_launch": scan_avg_ms,
                         "flop_per_px": FLOP_PER_PX, "px_per_launch": n_px,
                         "share_of_step": scan_avg_ms / (ms / args.steps)},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "stats": stats,
        }))
    sys.stdout.flush()
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

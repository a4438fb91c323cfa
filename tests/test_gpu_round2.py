"""Round-2 GPU tests: oracle parity on the benchmarked scenes (BASELINE.json configs[0] and configs[2]), re-entrancy of the
plans, the fused speed / direction epilogue (row F2), the hostile scene, and the device-resident
row-sharded gather over NCCL (row E1; needs two GPUs, skipped otherwise)."""
import os
import socket
import sys
import threading
import warnings

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def env():
    import torch

    assert torch.cuda.is_available()
    from xsarsea_b200 import _device, _native, windspeed
    from xsarsea_b200.windspeed import windspeed as impl

    return torch, _device, _native, windspeed, impl


def reset_steps(m):
    for k, v in dict(inc_step=0.1, wspd_step=0.1, phi_step=1.0, inc_step_lr=1.0, wspd_step_lr=0.2, phi_step_lr=2.5).items():
        setattr(m, k, v)
    return m


def default_plan(ws, impl):
    plan = impl._get_plan(reset_steps(ws.get_model("gmf_cmod5n")), reset_steps(ws.get_model("gmf_s1_v2")), 0.1, {})
    assert plan.co_lut.shape == (501, 499, 181) and plan.cr_lut.shape == (501, 771)
    return plan


def oracle_sample(plan, idx, inc, s_co, s_cr, anc, dsig=0.1):
    """oracle.invert (windspeed.py:183-282) on the pixels `idx`, with the LUTs the device inverted with (downloaded) and the
    dB prologue of windspeed.py:126-128 done by numpy."""
    h = lambda t: t.reshape(-1)[idx].cpu().numpy()
    kw = dict(co_lut=plan.co_lut.cpu().numpy(), inc_grid=plan.co_grids[0], wspd_grid=plan.co_grids[1], phi_grid=plan.co_grids[2])
    if s_cr is not None:
        kw.update(cr_lut=plan.cr_lut.cpu().numpy(), inc_cr_grid=plan.cr_grids[0], wspd_cr_grid=plan.cr_grids[1])
    with np.errstate(all="ignore"):
        co_db = 10 * np.log10(h(s_co) + 1e-15)
        cr_db = 10 * np.log10(h(s_cr) + 1e-15) if s_cr is not None else np.full(idx.numel(), np.nan)
        return oracle.invert(h(inc), co_db, cr_db, dsig, h(anc), **kw)


def assert_same_winds(got, want, what):
    """NaN pattern identical; values identical up to 1e-9 except ulp-level near-ties of the fused log10 prologue
    (at most 1e-4 of the pixels)."""
    assert np.array_equal(np.isnan(got), np.isnan(want)), what
    ok = ~np.isnan(want)
    bad = np.abs(got[ok] - want[ok]) > 1e-9
    assert bad.mean() <= 1e-4, f"{what}: {bad.sum()} of {bad.size} pixels differ"


def test_config3_full_iw_scene_sampled_against_the_oracle(env):
    """BASELINE.json configs[2], the benchmarked workload (bench.py's scene recipe, dual-pol, default LUTs): 24 000 pixels
    sampled over a 4 000-line strip of the scene equal the oracle, index for index."""
    torch, D, nat, ws, impl = env
    import bench

    plan = default_plan(ws, impl)
    inc, s_co, s_cr, anc = bench.synth_scene_device(4000, 25000, 0)
    co, du, ic, ix = plan.invert(inc, s_co, s_cr, 0.1, anc, merge_dual=False, want_idx=True)
    g = torch.Generator(device="cuda").manual_seed(1)
    idx = torch.randint(0, inc.numel(), (24000,), generator=g, device="cuda")
    o_co, o_du, o_ic, o_ix = oracle_sample(plan, idx, inc, s_co, s_cr, anc)
    h = lambda t: t.reshape(-1)[idx].cpu().numpy()
    assert_same_winds(h(co), o_co, "wind_co")
    assert_same_winds(h(du), o_du, "wind_dual")
    assert (h(ic) != o_ic).mean() <= 1e-4 and (h(ix) != o_ix).mean() <= 1e-4
    # merged output (windspeed.py:426-428) of the same call path the bench times
    _, merged, _, _ = plan.invert(inc, s_co, s_cr, 0.1, anc, merge_dual=True)
    with np.errstate(invalid="ignore"):
        want = np.where((np.abs(o_co) < 5) | (np.abs(o_du) < 5), o_co, o_du)
        # 5.0 is a node of both wspd grids: numpy's abs(w * exp(1j * phi)) of a wind on that node is 5 -+ 1 ulp depending
        # on libm, so there the merge may legitimately pick the other branch
        tie = (np.abs(np.abs(o_co) - 5) < 1e-9) | (np.abs(np.abs(o_du) - 5) < 1e-9)
    got = h(merged)
    assert 0 < tie.sum() < 0.02 * tie.size
    assert_same_winds(got[~tie], want[~tie], "merged dual")
    assert np.all(np.isclose(got[tie], o_co[tie], atol=1e-9) | np.isclose(got[tie], o_du[tie], atol=1e-9))


def test_config1_scene_sampled_against_the_oracle(env):
    """BASELINE.json configs[0] (1000 x 1000 co-pol, default cmod5n LUT): 20 000 sampled pixels equal the oracle."""
    torch, D, nat, ws, impl = env
    plan = default_plan(ws, impl)
    g = torch.Generator(device="cuda").manual_seed(0)
    H = W = 1000
    f64 = dict(device="cuda", dtype=torch.float64)
    inc = (17.5 + 32 * torch.arange(W, **f64) / (W - 1)).expand(H, W).contiguous()
    w = 2 + 23 * torch.rand(H, W, generator=g, **f64)
    p = 360 * torch.rand(H, W, generator=g, **f64)
    s_co = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], inc, w, p) * torch.exp(0.05 * torch.randn(H, W, generator=g, **f64))
    anc = torch.polar((w + 2 * torch.randn(H, W, generator=g, **f64)).abs(), torch.deg2rad(p + 20 * torch.randn(H, W, generator=g, **f64)))
    co, _, ic, _ = plan.invert(inc, s_co, None, 0.1, anc, want_idx=True)
    idx = torch.randint(0, H * W, (20000,), generator=g, device="cuda")
    o_co, _, o_ic, _ = oracle_sample(plan, idx, inc, s_co, None, anc)
    assert_same_winds(co.reshape(-1)[idx].cpu().numpy(), o_co, "wind_co")
    assert (ic.reshape(-1)[idx].cpu().numpy() != o_ic).mean() <= 1e-4


def test_hostile_scene_fast_equals_fp64(env):
    """bench.py's hostile scene (random incidence per pixel, +-15 dB sea/land runs, ancillary wind 10 m/s and 90 deg off,
    20 % NaN): the FP32 scan + refinement gives exactly the indices of the exhaustive FP64 kernel."""
    torch, D, nat, ws, impl = env
    import bench

    plan = default_plan(ws, impl)
    inc, s_co, s_cr, anc = bench.synth_scene_device(60, 25000, 3, scene="hostile")
    a, ax, ia, ixa = plan.invert(inc, s_co, s_cr, 0.1, anc, merge_dual=True, want_idx=True)
    st = plan.last_stats()
    b, bx, ib, ixb = plan.invert(inc, s_co, s_cr, 0.1, anc, merge_dual=True, want_idx=True, mode=nat.MODE_FP64)
    assert torch.equal(ia, ib) and torch.equal(ixa, ixb)
    bits = lambda z: torch.view_as_real(z).contiguous().view(torch.int64)
    assert torch.equal(bits(a), bits(b)) and torch.equal(bits(ax), bits(bx))
    n_co = int((ia >= 0).sum())
    assert st["scan_pixels"] + st["exhaustive_pixels"] == n_co and 0.7 * inc.numel() < n_co < 0.9 * inc.numel()


def test_concurrent_calls_on_shared_plans(env):
    """SURVEY.md section 8 B2 / ADVICE r1: the operator is called concurrently by dask's threaded scheduler.  Four host
    threads, each on its own CUDA stream, alternate between two model pairs through the numpy-level operator while the
    plan cache holds a single plan (every switch evicts the plan another thread may be using): results must be
    bit-identical to the serial ones."""
    torch, D, nat, ws, impl = env
    kw = dict(inc_step_lr=2.0, wspd_step_lr=1.0, phi_step_lr=10.0, inc_step=0.5, wspd_step=0.25, phi_step=2.5)
    rng = np.random.default_rng(5)
    n = 40000
    inc = rng.uniform(18, 48, n)
    w, p = rng.uniform(2, 25, n), rng.uniform(0, 360, n)
    with np.errstate(all="ignore"):
        co_db = 10 * np.log10(oracle.gmf_eval("gmf_cmod5n", inc, w, p) * np.exp(rng.normal(0, 0.05, n)) + 1e-15)
        cr_db = 10 * np.log10(oracle.gmf_eval("gmf_s1_v2", inc, w) * np.exp(rng.normal(0, 0.05, n)) + 1e-15)
    anc = (w + rng.normal(0, 2, n)) * np.exp(1j * np.deg2rad(p + rng.normal(0, 20, n)))
    dsig = np.full(n, 0.1)
    pairs = [(ws.get_model("gmf_cmod5n"), ws.get_model("gmf_s1_v2")), (ws.get_model("gmf_cmodifr2"), ws.get_model("gmf_rs2_v2"))]
    serial = [impl._invert_from_model_numpy(m, 0.1, dict(kw), inc, co_db, cr_db, dsig, anc) for m in pairs]
    old_max, impl._PLAN_CACHE_MAX = impl._PLAN_CACHE_MAX, 1
    errors, results = [], {}

    def work(tid):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for it in range(6):
                    k = (tid + it) % 2
                    results[(tid, it)] = (k, impl._invert_from_model_numpy(pairs[k], 0.1, dict(kw), inc, co_db, cr_db, dsig, anc))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    try:
        threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        impl._PLAN_CACHE_MAX = old_max
    assert not errors, errors
    assert len(results) == 24
    for (tid, it), (k, (oc, ox)) in results.items():
        assert np.array_equal(oc, serial[k][0], equal_nan=True) and np.array_equal(ox, serial[k][1], equal_nan=True), (tid, it)


def test_speed_direction_epilogue(env):
    """Row F2: the fused epilogue equals the callers' post-processing of the complex result
    (docs/examples/windspeed_retrieval_L1.ipynb cell 33: np.abs, (90 - np.angle(deg) + ground_heading) % 360)."""
    torch, D, nat, ws, impl = env
    kw = dict(inc_step_lr=2.0, wspd_step_lr=1.0, phi_step_lr=10.0, inc_step=0.5, wspd_step=0.25, phi_step=2.5)
    rng = np.random.default_rng(8)
    shape = (50, 120)
    inc = rng.uniform(18, 48, shape)
    w, p = rng.uniform(2, 25, shape), rng.uniform(0, 360, shape)
    s_co = oracle.gmf_eval("gmf_cmod5n", inc, w, p) * np.exp(rng.normal(0, 0.05, shape))
    s_cr = oracle.gmf_eval("gmf_s1_v2", inc, w) * np.exp(rng.normal(0, 0.05, shape))
    anc = (w + rng.normal(0, 2, shape)) * np.exp(1j * np.deg2rad(p + rng.normal(0, 20, shape)))
    s_co[rng.uniform(size=shape) < 0.02] = np.nan
    inc[0, :5] = np.nan
    gh = rng.uniform(0, 360, shape)
    model = ("gmf_cmod5n", "gmf_s1_v2")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        co, dual = ws.invert_from_model(inc, s_co, s_cr, ancillary_wind=anc, model=model, **kw)
        (sp_co, d_co), (sp_du, d_du) = ws.invert_to_speed_dir(inc, s_co, s_cr, ancillary_wind=anc, model=model, ground_heading=gh, **kw)
        (sa_co, da_co), _ = ws.invert_to_speed_dir(inc, s_co, s_cr, ancillary_wind=anc, model=model, **kw)       # antenna convention
        (s32, d32), _ = ws.invert_to_speed_dir(inc, s_co, s_cr, ancillary_wind=anc, model=model, ground_heading=190.0,
                                               dtype=np.float32, **kw)
        mono = ws.invert_to_speed_dir(inc, s_co, ancillary_wind=anc, model="gmf_cmod5n", ground_heading=gh, **kw)
        xonly = ws.invert_to_speed_dir(inc, s_cr, model="gmf_s1_v2", **kw)
        x_ref = ws.invert_from_model(inc, s_cr, model="gmf_s1_v2", **kw)
        t = lambda a: torch.from_numpy(a).cuda()
        (ts, td), _ = ws.invert_to_speed_dir(t(inc), t(s_co), t(s_cr), ancillary_wind=t(anc), model=model, ground_heading=t(gh), **kw)
    for z, sp, dr in ((co, sp_co, d_co), (dual, sp_du, d_du)):
        assert sp.shape == shape and sp.dtype == np.float64
        np.testing.assert_allclose(sp, np.abs(z), rtol=1e-13, atol=0, equal_nan=True)
        want = (90 - np.angle(z, deg=True) + gh) % 360
        d = np.abs(dr - want)
        d = np.minimum(d, 360 - d)          # 359.999.. vs 0 at the wrap
        assert np.array_equal(np.isnan(dr), np.isnan(want)) and np.nanmax(d) < 1e-9
        assert np.nanmin(dr) >= 0 and np.nanmax(dr) <= 360
    np.testing.assert_allclose(sa_co, np.abs(co), rtol=1e-13, equal_nan=True)
    np.testing.assert_allclose(da_co, np.angle(co, deg=True), rtol=0, atol=1e-9, equal_nan=True)
    assert s32.dtype == np.float32 and d32.dtype == np.float32
    np.testing.assert_allclose(s32, np.abs(co).astype(np.float32), rtol=1e-6, equal_nan=True)
    np.testing.assert_allclose(mono[0], sp_co, rtol=0, atol=0, equal_nan=True)
    np.testing.assert_allclose(mono[1], d_co, rtol=0, atol=0, equal_nan=True)
    np.testing.assert_array_equal(xonly, x_ref)
    assert ts.is_cuda and np.array_equal(ts.cpu().numpy(), sp_co, equal_nan=True) and np.array_equal(td.cpu().numpy(), d_co, equal_nan=True)


def test_streamed_blocks_pageable_and_pinned_inputs(env, monkeypatch):
    """The host path stages pageable inputs and all outputs through block-sized pinned buffers: results must not depend on
    the block size, on the inputs being page-locked, or on a previous call's staging buffers being reused."""
    torch, D, nat, ws, impl = env
    kw = dict(inc_step_lr=2.0, wspd_step_lr=1.0, phi_step_lr=10.0, inc_step=0.5, wspd_step=0.25, phi_step=2.5)
    rng = np.random.default_rng(2)
    shape = (37, 211)
    inc = rng.uniform(18, 48, shape)
    w, p = rng.uniform(2, 25, shape), rng.uniform(0, 360, shape)
    s_co = oracle.gmf_eval("gmf_cmod5n", inc, w, p)
    s_cr = oracle.gmf_eval("gmf_s1_v2", inc, w)
    anc = w * np.exp(1j * np.deg2rad(p + rng.normal(0, 20, shape)))
    model = ("gmf_cmod5n", "gmf_s1_v2")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ws.invert_from_model(inc, s_co, s_cr, ancillary_wind=anc, model=model, **kw)
        monkeypatch.setattr(impl, "BLOCK_PIXELS", 1000)       # 8 blocks, the last one ragged
        for _ in range(2):
            got = ws.invert_from_model(inc, s_co, s_cr, ancillary_wind=anc, model=model, **kw)
            assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, ref))
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
        got = ws.invert_from_model(pin(inc), pin(s_co), pin(s_cr), ancillary_wind=pin(anc), model=model, **kw)
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, ref))
        sd = ws.invert_to_speed_dir(inc, s_co, s_cr, ancillary_wind=anc, model=model, ground_heading=10.0, **kw)
        # a raster below one block is cut into four blocks (so that its copies overlap its kernels): same results
        monkeypatch.setattr(impl, "BLOCK_PIXELS", 1 << 26)
        monkeypatch.setattr(impl, "MIN_BLOCK_PIXELS", 1500)
        assert len(impl._block_edges(inc.size)) == 5
        got = ws.invert_from_model(inc, s_co, s_cr, ancillary_wind=anc, model=model, **kw)
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, ref))
    np.testing.assert_allclose(sd[1][0], np.abs(ref[1]), rtol=1e-13, equal_nan=True)


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import bench
    from xsarsea_b200 import parallel, windspeed
    from xsarsea_b200.windspeed import windspeed as impl

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        kw = dict(inc_step_lr=2.0, wspd_step_lr=1.0, phi_step_lr=10.0, inc_step=0.5, wspd_step=0.25, phi_step=2.5)
        plan = impl._get_plan(windspeed.get_model("gmf_cmod5n"), windspeed.get_model("gmf_s1_v2"), 0.1, kw)
        lines = 101
        full = bench.synth_scene_device(lines, 700, 7, 20.0, 45.0)          # same seed on every rank: the same scene
        lo, hi = parallel.row_shard(lines, world, rank)
        res = parallel.invert_rows_resident(plan, [t[lo:hi] for t in full], lines, lo, hi, dst=0, merge_dual=True)
        # ... and with the rows inverted in three sub-blocks whose transfers overlap the next inversion (side stream)
        res3 = parallel.invert_rows_resident(plan, [t[lo:hi] for t in full], lines, lo, hi, dst=0, merge_dual=True, pieces=3)
        torch.cuda.synchronize()
        ok = True
        if rank == 0:
            single = plan.invert(*full[:3], 0.1, full[3], merge_dual=True)
            bits = lambda z: torch.view_as_real(z).contiguous().view(torch.int64)
            ok = torch.equal(bits(res[0]), bits(single[0])) and torch.equal(bits(res[1]), bits(single[1]))
            ok = ok and torch.equal(bits(res3[0]), bits(single[0])) and torch.equal(bits(res3[1]), bits(single[1]))
        else:
            ok = res == (None, None)
        # the host-array API on top of it
        h = [t.cpu().numpy() for t in full]
        import warnings as w

        with w.catch_warnings():
            w.simplefilter("ignore")
            out = parallel.invert_sharded(h[0], h[1], h[2], ancillary_wind=h[3], model=("gmf_cmod5n", "gmf_s1_v2"), gather=0, **kw)
            if rank == 0:
                one = windspeed.invert_from_model(h[0], h[1], h[2], ancillary_wind=h[3], model=("gmf_cmod5n", "gmf_s1_v2"), **kw)
                ok = ok and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(out, one))
            else:
                ok = ok and out is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_row_sharded_gather_nccl_world2(env):
    """Row E1: one scene row-partitioned over 2 GPUs, results gathered on the device over NCCL into rank 0: bit-identical
    to the single-GPU inversion."""
    torch = env[0]
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_cross_pol_only_step_function_kernel(env, dtype):
    """Config 4's fast path (k_cross_only: the argmin as a step function of LINEAR sigma0, no log10 per pixel) against the
    general cross-pol pass (`cr_full_scan`: cooperative scan of all 771 candidates with the reference's operations) on
    sigma0 on and next to the midpoints between LUT nodes, on the nodes, outside the LUT's range, zero / negative / NaN /
    infinite sigma0, NaN incidence, and raster as well as scalar dsig; plus an oracle sample.  The two must agree bit for
    bit (both take the dB value from the same device log10 when they need one)."""
    torch, D, nat, ws, impl = env
    gi, gwc = np.linspace(16, 66, 501), np.linspace(3, 80, 771)
    cr = D.lut_to_db(D.lut_build(nat.GMF_IDS["gmf_s1_v2"], gi, gwc, None))
    plan = D.InversionPlan(cr=(cr, gi, gwc))
    crh = cr.cpu().numpy()
    rng = np.random.default_rng(77)
    n = 400_000
    b = rng.integers(0, gi.size, n)
    inc = gi[b] + rng.uniform(-0.04, 0.04, n)
    k = rng.integers(0, gwc.size - 1, n)
    kind = rng.integers(0, 8, n)
    s_db = rng.uniform(-50, 0, n)
    mid = 0.5 * crh[b, k] + 0.5 * crh[b, k + 1]
    s_db = np.where(kind == 0, crh[b, k], s_db)
    s_db = np.where(kind == 1, mid, s_db)
    x = 10.0 ** (s_db / 10.0) - 1e-15
    rel = rng.choice([0.0, 1e-16, -1e-16, 2e-16, 1e-14, -1e-14, 1e-12, -1e-12, 1e-10, -1e-10], n)
    x = np.where(kind <= 2, x * (1.0 + rel), x)
    x[3::1001] = 0.0
    x[4::1001] = -1e-15
    x[5::1001] = -1e-3
    x[6::1001] = np.nan
    x[7::1001] = np.inf
    x[8::1001] = 1e-300
    inc[9::1001] = np.nan
    dsig = 10.0 ** rng.uniform(-6, 3, n)
    dsig[10::1001] = 0.0
    dsig[11::1001] = np.nan
    dsig[12::1001] = -0.2
    if dtype == "f32":
        inc, x, dsig = (a.astype(np.float32) for a in (inc, x, dsig))
    d_inc, d_x, d_dsig = D.to_device(inc), D.to_device(x), D.to_device(dsig)
    for dsig_arg in (d_dsig, 0.1):
        for kw in (dict(), dict(cr_abs=True), dict(speed_dir=True)):
            _, fast, _, ifast = plan.invert(d_inc, None, d_x, dsig_arg, None, want_idx=True, **kw)
            st = plan.last_stats()
            _, slow, _, islow = plan.invert(d_inc, None, d_x, dsig_arg, None, want_idx=True, cr_full_scan=True, **kw)
            assert torch.equal(ifast, islow), (dtype, kw, torch.nonzero(ifast != islow)[:5])
            fr = torch.view_as_real(fast) if fast.is_complex() else fast
            sr = torch.view_as_real(slow) if slow.is_complex() else slow
            assert torch.equal(fr.view(torch.int64) if fr.dtype == torch.float64 else fr.view(torch.int32),
                               sr.view(torch.int64) if sr.dtype == torch.float64 else sr.view(torch.int32)), (dtype, kw)
            # the step-function kernel settled the bulk: what it leaves are the deliberately degenerate pixels and the guard hits
            assert 0 < st["cross_listed_pixels"] < 0.2 * n, st   # one pixel in eight sits on a midpoint here
    # an oracle sample (numpy's log10 may differ from the device's by an ulp: exact midpoints are excluded)
    sel = np.flatnonzero((kind > 1) | ((kind == 0) & (rel != 0)))[:30000]
    with np.errstate(all="ignore"):
        s_ref = 10 * np.log10(x[sel].astype(np.float64) + 1e-15)
        _, o_du, _, o_ix = oracle.invert(inc[sel].astype(np.float64), np.nan, s_ref, dsig[sel].astype(np.float64), np.nan + 0j,
                                         cr_lut=crh, inc_cr_grid=gi, wspd_cr_grid=gwc)
    _, _, _, ix = plan.invert(d_inc, None, d_x, d_dsig, None, want_idx=True)
    got = ix.cpu().numpy()[sel]
    assert (got != o_ix).mean() < 1e-4, np.flatnonzero(got != o_ix)[:10]
    # dB input takes the same kernel with the dB midpoint table
    s_in = np.where(np.isfinite(x) & (x + 1e-15 > 0), 10 * np.log10(np.abs(x.astype(np.float64)) + 1e-15), -30.0)
    d_s = D.to_device(s_in.astype(inc.dtype))
    _, _, _, i1 = plan.invert(d_inc, None, d_s, d_dsig, None, want_idx=True, sigma0_db=True)
    assert plan.last_stats()["cross_listed_pixels"] < 0.2 * n
    _, _, _, i2 = plan.invert(d_inc, None, d_s, d_dsig, None, want_idx=True, sigma0_db=True, cr_full_scan=True)
    assert torch.equal(i1, i2)


@pytest.mark.parametrize("scene", ["friendly", "hostile"])
def test_chunk_pruning_equals_brute_force(env, scene):
    """Exact chunk pruning (k_tile_plan): the scan that skips the chunks whose lower bound exceeds a seed's cost gives the
    indices of the brute-force scan (XS_FLAG_NO_PRUNE) and of the exhaustive FP64 kernel, bit for bit -- on the benchmark
    recipe and on the hostile scene -- and it really skips most of the slab."""
    torch, D, nat, ws, impl = env
    import bench

    plan = default_plan(ws, impl)
    inc, s_co, s_cr, anc = bench.synth_scene_device(160, 25000, 5, scene=scene)
    a, _, ia, _ = plan.invert(inc, s_co, None, 0.1, anc, want_idx=True)
    st = plan.last_stats()
    b, _, ib, _ = plan.invert(inc, s_co, None, 0.1, anc, want_idx=True, no_prune=True)
    st0 = plan.last_stats()
    c, _, ic, _ = plan.invert(inc, s_co, None, 0.1, anc, want_idx=True, mode=nat.MODE_FP64)
    assert torch.equal(ia, ib) and torch.equal(ia, ic)
    bits = lambda z: torch.view_as_real(z).contiguous().view(torch.int64)
    assert torch.equal(bits(a), bits(b)) and torch.equal(bits(a), bits(c))
    n_chunks = (499 + 7) // 8 * 8 // 16 + (1 if ((499 + 7) // 8 * 8) % 16 else 0)
    assert st0["chunks_streamed"] == st0["tiles"] * n_chunks and st0["warp_chunk_phi"] == 4 * 181 * st0["chunks_streamed"]
    assert st["tiles"] == st0["tiles"] and st["scan_pixels"] == st0["scan_pixels"]
    assert 4 * st["tiles"] <= st["chunks_streamed"] < 0.6 * st0["chunks_streamed"], (st, st0)
    assert st["warp_chunk_phi"] < 0.5 * st0["warp_chunk_phi"]


def test_chunk_pruning_small_and_ragged_rasters(env):
    """Sparse bins (tiles whose pixels differ widely in sigma0), far-off and zero sigma0 / ancillary winds, coarse LUTs with
    few chunks and a long wspd grid (several chunks per mask bit): pruned == brute force == FP64."""
    torch, D, nat, ws, impl = env
    g = torch.Generator(device="cuda").manual_seed(11)
    f64 = dict(device="cuda", dtype=torch.float64)
    # the last grid has 69 chunks: a bit of the 32-bit chunk masks then covers 4 chunks
    for (n_w, n_p, w_hi) in ((499, 181, 50.0), (70, 37, 35.0), (130, 73, 80.0), (1100, 91, 60.0)):
        gi, gw, gp = np.linspace(17, 50, 34), np.linspace(0.2, w_hi, n_w), np.linspace(0, 180, n_p)
        co = D.lut_to_db(D.lut_build(nat.GMF_IDS["gmf_cmod5n"], gi, gw, gp))
        plan = D.InversionPlan(co=(co, gi, gw, gp), cr=None)
        n = 30011
        inc = 17 + 33 * torch.rand(n, generator=g, **f64)
        w = 0.5 + 40 * torch.rand(n, generator=g, **f64)
        p = 360 * torch.rand(n, generator=g, **f64)
        s = D.gmf_eval(nat.GMF_IDS["gmf_cmod5n"], inc, w, p) * torch.exp(0.5 * torch.randn(n, generator=g, **f64))
        s[::97] *= 1e4  # far outside the LUT
        s[5::101] = 0.0
        anc = torch.polar(60 * torch.rand(n, generator=g, **f64), 6.3 * torch.rand(n, generator=g, **f64))
        anc[7::89] = 0
        _, _, ia, _ = plan.invert(inc, s, None, 0.1, anc, want_idx=True)
        st = plan.last_stats()
        _, _, ib, _ = plan.invert(inc, s, None, 0.1, anc, want_idx=True, no_prune=True)
        _, _, ic, _ = plan.invert(inc, s, None, 0.1, anc, want_idx=True, mode=nat.MODE_FP64)
        assert torch.equal(ia, ib) and torch.equal(ia, ic), (n_w, n_p)
        assert st["chunks_streamed"] >= 4 * st["tiles"]
        plan.close()

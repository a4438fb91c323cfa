"""CPU-only checks: the C-ABI library loads and exports every symbol of include/xsarsea_b200.h, the host-side mirror
of the reference interface (registry, aliases, labelled-array shim, utils), the loud failure without a GPU, and the
row sharding + gather on a world_size-2 gloo group."""
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from xsarsea_b200 import _native

    _native.build()
    lib = _native.load()
    header = open(os.path.join(ROOT, "include", "xsarsea_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(xs_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.xs_abi_version() == 2
    assert lib.xs_launch_count() == 0    # loading the library launches nothing


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from xsarsea_b200 import _native, windspeed

    n = 16
    with pytest.raises(_native.NativeError, match="CUDA device"):
        windspeed.invert_from_model(np.full(n, 35.0), np.full(n, 0.05), ancillary_wind=np.full(n, 5 + 5j),
                                    model="gmf_cmod5n")
    with pytest.raises(_native.NativeError):
        windspeed.get_model("gmf_cmod5n").to_lut()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "xsarsea_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "xs_oracle" not in src, f


def test_registry_and_aliases():
    from xsarsea_b200 import windspeed as ws
    from xsarsea_b200.windspeed.models import LutModel, Model

    df = ws.available_models()
    assert list(df.columns) == ["alias", "pol", "model"]
    for name in ("gmf_cmod5", "gmf_cmod5n", "gmf_cmod5n_pr_zhangA", "gmf_cmod5n_pr_mouche1", "gmf_cmodifr2", "gmf_rs2_v2",
                 "gmf_s1_v2", "gmf_rcm_noaa", "gmf_s1_v3_ew_rec", "gmf_rs2_v3", "gmf_rcm_v3", "gmf_rcm_v4", "gmf_rs2_v4"):
        assert name in df.index
    assert ws.get_model("cmod5n") is ws.get_model("gmf_cmod5n")
    m = ws.get_model("gmf_cmod5n")
    assert ws.get_model(m) is m and m.iscopol and m.pol == "VV" and m.phi_range == [0.0, 180.0]
    assert m.wspd_range == [0.2, 50.0] and m.inc_range == [16.0, 66.0] and m.units == "linear"
    x = ws.get_model("gmf_rcm_v4")
    assert x.iscrosspol and x.phi_range is None and x.wspd_range == [3.0, 80.0]
    assert set(ws.available_models(pol="HH").index) == {"gmf_cmod5n_pr_zhangA", "gmf_cmod5n_pr_mouche1"}
    with pytest.raises(KeyError, match="not found"):
        ws.get_model("nope")

    # alias ownership by priority (models.py:477-482): a priority-1 model with the same short name takes the alias
    # from the GmfModel (priority 3), exactly how Cmod7Model (1) would win over a gmf_cmod7 GmfModel
    class Fake(LutModel):
        _name_prefix = "fake_"
        _priority = 1

    try:
        f = Fake("fake_cmod5n", pol="VV")
        df = ws.available_models()
        assert df.loc["fake_cmod5n", "alias"] == "cmod5n" and df.loc["gmf_cmod5n", "alias"] is None
        assert ws.get_model("cmod5n") is f and ws.get_model("gmf_cmod5n") is m
    finally:
        Model._available_models.pop("fake_cmod5n", None)
    assert ws.get_model("cmod5n") is m


def test_gmf_register_decorator_host_side():
    from xsarsea_b200 import windspeed as ws
    from xsarsea_b200.windspeed.models import Model

    with pytest.raises(ValueError, match="must start with"):
        ws.GmfModel.register(pol="VV")(lambda i, w, p: 1.0)

    try:
        @ws.GmfModel.register(name="gmf_host360", pol="VV", units="linear", defer=False)
        def _f(inc, wspd, phi):
            return 1e-3 * wspd * (2 + np.cos(np.deg2rad(phi)) + 0.3 * np.sin(np.deg2rad(phi)))

        @ws.GmfModel.register(name="gmf_host180", pol="HH", units="linear", defer=False)
        def _g(inc, wspd, phi):
            return 1e-3 * wspd * (2 + np.cos(np.deg2rad(phi)))

        @ws.GmfModel.register(name="gmf_hostx", pol="VH", units="linear", defer=True)
        def _h(inc, wspd, phi=None):
            return 1e-4 * wspd

        # gmfs.py:146-158 probes phi in [0, 90, 180, 270] and takes the *min* of |f(phi) - f(-phi)|: phi = 0 always gives
        # 0, so the reference classifies every gmf that uses phi as [0, 180] -- reproduced as is
        assert ws.get_model("gmf_host360").phi_range == [0.0, 180.0]
        assert ws.get_model("gmf_host180").phi_range == [0.0, 180.0]
        assert "gmf_hostx" not in Model._available_models               # deferred
        ws.GmfModel.activate_gmfs_impl(gmfs_names=["gmf_hostx"])
        assert ws.get_model("gmf_hostx").phi_range is None and ws.get_model("gmf_hostx").wspd_range == [3.0, 80.0]
    finally:
        for n in ("gmf_host360", "gmf_host180", "gmf_hostx"):
            Model._available_models.pop(n, None)
        ws.GmfModel._deferred_registrations[:] = [d for d in ws.GmfModel._deferred_registrations if d[1] != "gmf_hostx"]


def test_dataarray_lite():
    from xsarsea_b200._xr import DataArrayLite, is_labelled, like

    a = DataArrayLite(np.arange(24.0).reshape(2, 3, 4), ("incidence", "wspd", "phi"),
                      dict(incidence=[1, 2], wspd=[1, 2, 3], phi=[0, 1, 2, 3]), dict(units="dB"), "lut")
    assert is_labelled(a) and a.shape == (2, 3, 4) and np.asarray(a).sum() == 276
    assert np.array_equal(np.asarray(a.wspd), [1, 2, 3])
    t = a.transpose("wspd", "phi", "incidence")
    assert t.dims == ("wspd", "phi", "incidence") and t.shape == (3, 4, 2) and t.values[1, 2, 1] == a.values[1, 1, 2]
    s = a.isel(incidence=0)
    assert s.dims == ("wspd", "phi") and "incidence" not in s.coords
    assert a.isel(phi=[0]).squeeze("phi").dims == ("incidence", "wspd")
    b = like(a, np.zeros(a.shape), name="z")
    assert b.dims == a.dims and b.attrs == {} and b.name == "z"
    assert not is_labelled(np.zeros(3))


def test_windspeed_utils_host_side():
    """Argument errors of the dsig helpers are raised on the host, as in the reference (utils.py:88-91, :121-122);
    the arithmetic itself is device-only (tests/test_gpu_dsig.py) and refuses to run without a GPU."""
    from xsarsea_b200 import windspeed as ws
    from xsarsea_b200._native import NativeError

    x = np.ones(4)
    with pytest.raises(ValueError, match="dsig names different"):
        ws.get_dsig("other", x, x, x)
    with pytest.raises(UnboundLocalError):
        ws.get_dsig_wspd("dsig_wspd_other", x, x)
    with pytest.raises(IndexError, match="Only 2D"):
        ws.nesz_flattening(x, x)
    import torch

    if not torch.cuda.is_available():
        for call in (lambda: ws.get_dsig("gmf_rs2_v2", x, x, x), lambda: ws.get_dsig_wspd("dsig_wspd_rcm_v3", x, x),
                     lambda: ws.nesz_flattening(np.ones((2, 4)), np.ones((2, 4)))):
            with pytest.raises(NativeError):
                call()


def test_row_shard_partition():
    from xsarsea_b200.parallel import row_shard

    for n, w in ((16700, 8), (10, 3), (5, 8), (0, 2)):
        blocks = [row_shard(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(b[1] == c[0] for b, c in zip(blocks, blocks[1:]))
        sizes = [b[1] - b[0] for b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from xsarsea_b200.parallel import invert_rows_resident, invert_sharded, row_shard

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        inc = rng.uniform(20, 40, (11, 7))
        s0, s1 = rng.uniform(0.01, 0.1, (11, 7)), rng.uniform(0.001, 0.01, (11, 7))
        anc = rng.normal(size=(11, 7)) + 1j * rng.normal(size=(11, 7))
        seen = {}

        def stub(i, a, b=None, ancillary_wind=None, dsig_cr=0.1, model=None):
            assert isinstance(i, torch.Tensor)      # the sharded path hands the backend's tensors to the inversion
            seen["rows"] = i.shape[0]
            co = (i + a) * ancillary_wind
            return (co, co * dsig_cr + b) if b is not None else co

        full = invert_sharded(inc, s0, s1, ancillary_wind=anc, dsig_cr=0.5, model="m", gather="all", _invert=stub)
        lo, hi = row_shard(11, world, rank)
        ok = seen["rows"] == hi - lo
        want_co = (inc + s0) * anc
        ok &= np.array_equal(full[0], want_co) and np.array_equal(full[1], want_co * 0.5 + s1)
        only0 = invert_sharded(inc, s0, ancillary_wind=anc, model="m", _invert=stub)          # default: rank 0 gets the result
        ok &= (only0 is None) if rank != 0 else np.array_equal(only0, want_co)
        only1 = invert_sharded(inc, s0, s1, ancillary_wind=anc, dsig_cr=0.5, model="m", gather=1, _invert=stub)
        ok &= (only1 is None) if rank != 1 else (np.array_equal(only1[0], want_co) and np.array_equal(only1[1], want_co * 0.5 + s1))
        mine = invert_sharded(inc, s0, ancillary_wind=anc, model="m", gather=None, _invert=stub)
        ok &= np.array_equal(mine, want_co[lo:hi])
        # fewer lines than ranks: a rank without rows still takes part in the gather
        one = invert_sharded(inc[:1], s0[:1], ancillary_wind=anc[:1], model="m", gather="all", _invert=stub)
        ok &= np.array_equal(one, want_co[:1])
        # the device-resident form bench.py times (here: CPU tensors, a stub in place of plan.invert)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a[lo:hi]))

        def stub_rows(i, a, b, d, c, oc, ox):
            oc.copy_((i + a) * c)
            ox.copy_((i + a) * c * d + b)

        co_f, cr_f = invert_rows_resident(None, (t(inc), t(s0), t(s1), t(anc)), 11, lo, hi, dst=0, dsig_cr=0.5, _invert=stub_rows)
        if rank == 0:
            ok &= np.array_equal(co_f.numpy(), want_co) and np.array_equal(cr_f.numpy(), want_co * 0.5 + s1)
        else:
            ok &= co_f is None and cr_f is None
        # the same with the rows of every rank inverted and sent in three sub-blocks (the pipelined form), gathered on rank 1
        co_f, cr_f = invert_rows_resident(None, (t(inc), t(s0), t(s1), t(anc)), 11, lo, hi, dst=1, dsig_cr=0.5, pieces=3,
                                          _invert=stub_rows)
        if rank == 1:
            ok &= np.array_equal(co_f.numpy(), want_co) and np.array_equal(cr_f.numpy(), want_co * 0.5 + s1)
        else:
            ok &= co_f is None and cr_f is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_row_sharded_gather_gloo_world2():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_direction_helpers():
    import xsarsea_b200 as x

    a = np.array([-190.0, -180.0, 0.0, 179.0, 180.0, 370.0])
    np.testing.assert_allclose(x.dir_to_180(a), [170.0, -180.0, 0.0, 179.0, -180.0, 10.0])
    np.testing.assert_allclose(x.dir_to_360(a), [170.0, 180.0, 0.0, 179.0, 180.0, 10.0])
    np.testing.assert_allclose(x.dir_oceano_to_meteo(x.dir_meteo_to_oceano(a)), x.dir_to_360(a))
    assert x.dir_sample_to_meteo(30.0, 350.0) == 410.0                     # detrend.py: 90 - sample_dir + heading
    assert abs(x.dir_meteo_to_sample(100.0, 10.0) - 0.0) < 1e-15           # pi/2 - deg2rad(meteo - heading)
    back = x.dir_sample_to_meteo(np.rad2deg(x.dir_meteo_to_sample(a, 12.0)), 12.0)
    np.testing.assert_allclose(back, a)


def test_gradients_host_side():
    """Argument errors of local_gradients are raised on the host; without a GPU the call refuses to run (no CPU
    fallback); the dataset stand-in gives attribute and item access like xarray.Dataset."""
    import torch

    from xsarsea_b200 import _xr, gradients
    from xsarsea_b200._native import NativeError

    with pytest.raises(ValueError, match="2D image"):
        gradients.local_gradients(np.zeros((2, 3, 4)))
    with pytest.raises(ValueError, match="2D image"):
        gradients.local_gradients(torch.zeros(5))
    if not torch.cuda.is_available():
        with pytest.raises(NativeError):
            gradients.local_gradients(np.zeros((8, 8)))
    ds = _xr.DatasetLite(G2=np.ones(2), c=np.zeros(2))
    assert ds.G2 is ds["G2"] and set(ds.data_vars) == {"G2", "c"}
    with pytest.raises(AttributeError):
        ds.G3


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path; runs without a GPU) prints exactly one JSON line with the
    keys the driver reads, on the same metric / unit / workload as the CUDA arm."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-lines", "16"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    import bench

    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "px/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "gmf_cmod5n" in d["config"]["workload"] and "sample" in d["config"]


def test_nc_lut_files_host_side(tmp_path):
    """Row A6 on the host: LUT files in the reference's schema (models.py:232-262, 368-379) are parsed into models --
    name = file stem with the `nc_lut_` prefix (:446), alias = the `model` attribute, low-resolution files move their
    steps into `*_lr` (:392-395), files without the prefix are ignored, the raw table comes back as written."""
    from scipy.io import netcdf_file

    from xsarsea_b200 import windspeed as ws
    from xsarsea_b200.windspeed.models import Model, NcLutModel

    def write(path, resolution, inc, wspd, phi=None, units="dB", pol="VH", model="testlut"):
        with netcdf_file(str(path), "w") as nc:
            nc.units, nc.pol, nc.model, nc.resolution = units, pol, model, resolution
            nc.inc_range, nc.wspd_range = np.array([inc[0], inc[-1]]), np.array([wspd[0], wspd[-1]])
            nc.inc_step, nc.wspd_step = float(inc[1] - inc[0]), float(wspd[1] - wspd[0])
            nc.createDimension("incidence", inc.size)
            nc.createDimension("wspd", wspd.size)
            nc.createVariable("incidence", "d", ("incidence",))[:] = inc
            nc.createVariable("wspd", "d", ("wspd",))[:] = wspd
            dims = ("incidence", "wspd")
            shape = (inc.size, wspd.size)
            if phi is not None:
                nc.phi_range, nc.phi_step = np.array([phi[0], phi[-1]]), float(phi[1] - phi[0])
                nc.createDimension("phi", phi.size)
                nc.createVariable("phi", "d", ("phi",))[:] = phi
                dims, shape = dims + ("phi",), shape + (phi.size,)
            table = -30.0 + np.arange(np.prod(shape), dtype=np.float64).reshape(shape) * 1e-3
            nc.createVariable("sigma0_model", "d", dims)[:] = table
        return table

    inc, wspd, phi = np.linspace(17, 50, 34), np.linspace(3, 80, 78), np.linspace(0, 180, 37)
    t_hi = write(tmp_path / "nc_lut_testlut.nc", "high", inc, wspd)
    write(tmp_path / "nc_lut_testlow.nc", "low", inc[::2], wspd[::2], phi, pol="VV", model="testlow")
    write(tmp_path / "other_testlut.nc", "high", inc, wspd)          # no nc_lut_ prefix: not registered
    try:
        ws.register_nc_luts(str(tmp_path))
        df = ws.available_models()
        assert "nc_lut_testlut" in df.index and "nc_lut_testlow" in df.index and "other_testlut" not in df.index
        hi = ws.get_model("nc_lut_testlut")
        assert isinstance(hi, NcLutModel) and ws.get_model("testlut") is hi and hi.iscrosspol and hi.pol == "VH"
        assert hi.units == "dB" and hi.resolution == "high" and hi.inc_range == [17.0, 50.0] and hi.wspd_range == [3.0, 80.0]
        assert hi.phi_range is None and abs(hi.inc_step - 1.0) < 1e-12 and abs(hi.wspd_step - 1.0) < 1e-12
        lo = ws.get_model("testlow")
        assert lo.iscopol and lo.resolution == "low" and lo.phi_range == [0.0, 180.0]
        assert abs(lo.inc_step_lr - 2.0) < 1e-12 and abs(lo.wspd_step_lr - 2.0) < 1e-12 and abs(lo.phi_step_lr - 5.0) < 1e-12
        assert lo.inc_step == 0.1 and lo.wspd_step == 0.1 and lo.phi_step == 1           # defaults of models.py:46-48
        vals, g_inc, g_wspd, g_phi, units, res = hi._raw_lut_host()
        assert np.array_equal(vals, t_hi) and np.array_equal(g_inc, inc) and np.array_equal(g_wspd, wspd) and g_phi is None
        assert units == "dB" and res == "high"
        assert ws.register_nc_luts(str(tmp_path), gmf_names=["nothing"]) is None   # filter by name: nothing added
    finally:
        for n in ("nc_lut_testlut", "nc_lut_testlow"):
            Model._available_models.pop(n, None)


def test_cmod7_table_reader_host_side(tmp_path):
    """Row A7 on the host: the KNMI table layout (cmod7.py:19-75) -- little-endian float32, one leading and one trailing
    record marker dropped, (wspd, phi, inc) in Fortran order, linear units, low resolution -- and the registration
    (name gmf_cmod7, priority 1 so that it owns the `cmod7` alias)."""
    from xsarsea_b200 import windspeed as ws
    from xsarsea_b200.windspeed.models import Model

    n_w, n_p, n_i = 250, 73, 51
    rng = np.random.default_rng(1)
    table = rng.uniform(1e-4, 1.0, (n_i, n_w, n_p)).astype(np.float32)              # [inc][wspd][phi]
    fortran = np.transpose(table, (1, 2, 0))                                         # (wspd, phi, inc)
    rec = np.concatenate([[123.0], fortran.reshape(-1, order="F"), [456.0]]).astype("<f4")
    d = tmp_path / "cmod7"
    d.mkdir()
    rec.tofile(str(d / "gmf_cmod7_vv.dat_little_endian"))
    try:
        ws.register_cmod7(str(d))
        m = ws.get_model("gmf_cmod7")
        assert ws.get_model("cmod7") is m and m.pol == "VV" and m._priority == 1 and m.iscopol
        vals, inc, wspd, phi, units, res = m._raw_lut_host()
        assert vals.shape == (n_i, n_w, n_p) and vals.dtype == np.float64 and units == "linear" and res == "low"
        assert np.array_equal(vals, table.astype(np.float64))
        assert inc[0] == 16 and inc[-1] == 66 and inc.size == n_i and phi[0] == 0 and phi[-1] == 180 and phi.size == n_p
        assert abs(wspd[0] - 0.2) < 1e-12 and abs(wspd[-1] - 50.0) < 1e-9 and wspd.size == n_w
    finally:
        Model._available_models.pop("gmf_cmod7", None)
    import pytest as _pytest

    try:
        ws.register_cmod7(str(tmp_path / "missing"))
        with _pytest.raises(FileNotFoundError):
            ws.get_model("gmf_cmod7")._raw_lut_host()
    finally:
        Model._available_models.pop("gmf_cmod7", None)


def test_piece_edges_shrink_and_cover():
    """Sub-blocks of a rank's rows (parallel.piece_edges): contiguous, covering, non-increasing in size, every one non-empty
    when there are enough rows, and the last one small (its transfer is the exposed one)."""
    from xsarsea_b200 import parallel as P

    for lo, hi, p in ((0, 2087, 3), (100, 8450, 4), (5, 9, 4), (0, 3, 4), (0, 0, 2), (10, 11, 3), (0, 16700, 1)):
        e = P.piece_edges(lo, hi, p)
        assert len(e) == p + 1 and e[0] == lo and e[-1] == hi
        sizes = [b - a for a, b in zip(e, e[1:])]
        assert all(s >= 0 for s in sizes) and sum(sizes) == hi - lo
        if hi - lo >= p:
            assert all(s >= 1 for s in sizes)
        if hi - lo >= 100 * p:
            assert all(a >= b for a, b in zip(sizes, sizes[1:])) and (p == 1 or sizes[-1] <= 0.25 * (hi - lo) + 1)
    assert P.n_pieces(16700, 25000, 8) == 3 and P.n_pieces(16700, 25000, 2) == 4 and P.n_pieces(100, 100, 2) == 1


def test_block_edges_of_the_host_path():
    """Compute blocks of `_run_device`: contiguous cover; one block for small rasters, four for a raster below one block,
    short first / last blocks beyond."""
    from xsarsea_b200.windspeed import windspeed as impl

    for n in (1, 1000, impl.MIN_BLOCK_PIXELS, 2 * impl.MIN_BLOCK_PIXELS + 5, 50_000_000, impl.BLOCK_PIXELS, impl.BLOCK_PIXELS + 1, 417_500_000):
        e = impl._block_edges(n)
        assert e[0] == 0 and e[-1] == n and all(b > a for a, b in zip(e, e[1:]))
        assert max(b - a for a, b in zip(e, e[1:])) <= impl.BLOCK_PIXELS + impl.STAGE_PIXELS
    assert len(impl._block_edges(1000)) == 2 and len(impl._block_edges(50_000_000)) == 5
    big = impl._block_edges(417_500_000)
    assert big[1] - big[0] == impl.STAGE_PIXELS and big[-1] - big[-2] == impl.STAGE_PIXELS

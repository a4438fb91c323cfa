"""Strong-scaling measurement (N GPUs, ONE scene): the rows of a full Sentinel-1 IW dual-pol scene are partitioned over
the ranks (xsarsea_b200.parallel.row_shard), every rank inverts its block with no communication, and the two complex128
result rasters are assembled on every rank with one NCCL all-gather each.  Device-resident
on both sides; time = max over ranks of (invert + gather), CUDA events.
Run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [lines samples]"""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xsarsea_b200 import windspeed
from xsarsea_b200.parallel import row_shard
from xsarsea_b200.windspeed import windspeed as ws_impl

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
H = int(sys.argv[1]) if len(sys.argv) > 1 else bench.LINES
W = int(sys.argv[2]) if len(sys.argv) > 2 else bench.SAMPLES
sys.stdout.flush()
fd = os.dup(1)
os.dup2(2, 1)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tmp = tempfile.mkdtemp(prefix=f"xs_sh_{rank}_")
bench.write_ms1ahw_standin(tmp)
windspeed.register_nc_luts(tmp)
plan = ws_impl._get_plan(windspeed.get_model("gmf_cmod5n"), windspeed.get_model("nc_lut_cmodms1ahw"), 0.1, {})
lo, hi = row_shard(H, world, rank)
rows_max = max(row_shard(H, world, r)[1] - row_shard(H, world, r)[0] for r in range(world))
inc, s_co, s_cr, anc = bench.synth_scene_device(hi - lo, W, 100 + rank)     # this rank's row block
blk_co = torch.zeros(rows_max, W, dtype=torch.complex128, device="cuda")      # padded to the largest block
blk_du = torch.zeros_like(blk_co)
full_co = torch.empty(world * rows_max, W, dtype=torch.complex128, device="cuda")
full_du = torch.empty_like(full_co)


def step():
    plan.invert(inc, s_co, s_cr, 0.1, anc, merge_dual=True, out_co=blk_co[: hi - lo], out_cr=blk_du[: hi - lo])
    dist.all_gather_into_tensor(torch.view_as_real(full_co), torch.view_as_real(blk_co))
    dist.all_gather_into_tensor(torch.view_as_real(full_du), torch.view_as_real(blk_du))


for _ in range(2):
    step()
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
plan.invert(inc, s_co, s_cr, 0.1, anc, merge_dual=True, out_co=blk_co[: hi - lo], out_cr=blk_du[: hi - lo])
e[1].record()
dist.all_gather_into_tensor(torch.view_as_real(full_co), torch.view_as_real(blk_co))
dist.all_gather_into_tensor(torch.view_as_real(full_du), torch.view_as_real(blk_du))
e[2].record()
torch.cuda.synchronize()
t = torch.tensor([e[0].elapsed_time(e[2]), e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
# every rank now holds every block: rank r's rows sit at [r*rows_max, r*rows_max + its block size)
bits = lambda z: torch.view_as_real(z).contiguous().view(torch.int64)   # bit comparison: NaN pixels must match too
ok = torch.equal(bits(full_co[rank * rows_max: rank * rows_max + hi - lo]), bits(blk_co[: hi - lo]))
nxt = (rank + 1) % world   # and a neighbour's block really arrived: its NaN pattern is not all-zero padding
ok &= bool(full_du[nxt * rows_max: nxt * rows_max + 8].isnan().any() or full_du[nxt * rows_max: nxt * rows_max + 8].abs().sum() > 0)
sys.stdout.flush()
os.dup2(fd, 1)
if rank == 0:
    tot, inv, gat = (float(x) for x in t)
    print(json.dumps({"what": "one scene row-sharded over N GPUs, results all-gathered over NCCL", "n_gpus": world, "lines": H,
                      "samples": W, "ms_total": tot, "ms_invert": inv, "ms_all_gather": gat, "Mpx_per_s": H * W / tot / 1e3,
                      "gather_GBps_per_rank_out": 2 * 16 * rows_max * W * (world - 1) / gat / 1e6, "scaling": "strong",
                      "self_check": bool(ok)}))
sys.stdout.flush()
os.dup2(2, 1)
dist.destroy_process_group()

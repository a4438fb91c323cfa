"""torchrun check (N GPUs): `parallel.invert_sharded` over NCCL equals the single-GPU inversion of the same scene.
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_nccl_sharded.py"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xsarsea_b200 import windspeed
from xsarsea_b200.parallel import invert_sharded

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(0)
H, W = 203, 500
inc = np.broadcast_to(np.linspace(25, 45, W), (H, W)).copy()
s_co = rng.uniform(0.005, 0.2, (H, W))
s_cr = rng.uniform(0.0005, 0.01, (H, W))
anc = rng.uniform(2, 20, (H, W)) * np.exp(1j * rng.uniform(-np.pi, np.pi, (H, W)))
kw = dict(ancillary_wind=anc, dsig_cr=0.1, model=("gmf_cmod5n", "gmf_s1_v2"), inc_step=0.5, wspd_step=0.5, phi_step=2.5,
          inc_step_lr=2.0, wspd_step_lr=1.0, phi_step_lr=10.0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    co, dual = invert_sharded(inc, s_co, s_cr, **kw)                 # all_gather over NCCL
    only0 = invert_sharded(inc, s_co, s_cr, gather=0, **kw)          # gather to rank 0
    ref_co, ref_dual = windspeed.invert_from_model(inc, s_co, s_cr, **kw)
ok = np.array_equal(co, ref_co, equal_nan=True) and np.array_equal(dual, ref_dual, equal_nan=True)
ok &= (only0 is None) if rank != 0 else (np.array_equal(only0[0], ref_co, equal_nan=True))
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NCCL_SHARDED_OK" if t.item() == 1 else "NCCL_SHARDED_MISMATCH", "world", dist.get_world_size())
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)

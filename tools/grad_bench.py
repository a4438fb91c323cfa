"""Timing of xs_local_gradients on an EW-sized raster (dev tool)."""
import sys

import torch

sys.path.insert(0, ".")
from xsarsea_b200 import _device as D

H = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 10400
img = 0.1 + 0.02 * torch.rand(H, W, dtype=torch.float64, device="cuda")
for _ in range(3):
    D.local_gradients(img)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    D.local_gradients(img)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"local_gradients {H}x{W}: {ms:.3f} ms, {H*W/ms/1e6:.1f} Gpx/s, {16*H*W/ms/1e6:.0f} GB/s algorithmic")

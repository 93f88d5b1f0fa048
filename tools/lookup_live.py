"""Live (warm, back-to-back) timing of the lookup operator alone for the kernel builds (option lookup_mode)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnpose_b200 import ops
dev = torch.device("cuda:0"); B, h, w = 32, 30, 40
f1 = torch.randn(B, 256, h, w, device=dev); f2 = torch.randn(B, 256, h, w, device=dev)
pyr = ops.corr_pyramid(f1, f2)
yy, xx = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
coords = (torch.stack([xx, yy], -1).float()[None].repeat(B, 1, 1, 1) + 3 * torch.randn(B, h, w, 2, device=dev)).reshape(-1, 2).contiguous()
for mode in (1, 2, 4, 5, 6, 1, 2):
    ops.set_option("lookup_mode", mode)
    for _ in range(3): ops.corr_lookup(pyr, coords, B, h, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ops.corr_lookup(pyr, coords, B, h, w)
    e1.record(); torch.cuda.synchronize()
    print(f"lookup_mode {mode}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call (fp32 output, live)")

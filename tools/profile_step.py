"""Minimal driver for ncu: a couple of passes of the fused inner loop at the bench shape on cheap random inputs.
usage: python tools/profile_step.py [--fp32] [--passes N] [--batch B]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnpose_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fp32", action="store_true")
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--time", action="store_true")
ap.add_argument("--iters", type=int, default=4)
a = ap.parse_args()
H, W, B = 240, 320, a.batch
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
sd = torch.load(os.path.join(os.path.dirname(__file__), "..", "rnnpose_b200", "weights", "gru_update.pth"), map_location="cpu")
packed = ops.pack_weights({k[len("update_block."):]: v.float() for k, v in sd.items()}, dev)
f1 = torch.randn(B, 256, H // 8, W // 8, device=dev); f2 = torch.randn(B, 256, H // 8, W // 8, device=dev)
ctx = 0.1 * torch.randn(B, 256, H, W, device=dev)
g1 = torch.nn.functional.normalize(torch.randn(B, 32, H, W, device=dev), dim=1)
g2 = torch.nn.functional.normalize(torch.randn(B, 32, H, W, device=dev), dim=1)
yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
depth = (((yy - H / 2) ** 2 / (0.35 * H) ** 2 + (xx - W / 2) ** 2 / (0.3 * W) ** 2) < 1).float()[None].repeat(B, 1, 1) * 0.9
K = torch.tensor([[600.0, 0, W / 2], [0, 600.0, H / 2], [0, 0, 1]], device=dev)[None].repeat(B, 1, 1).contiguous()
ws = ops.RefineWorkspace(B, H, W, dev)
flags = ops.FLAG_EXACT_FP32 if a.fp32 else ops.FLAG_TENSOR_CORES
for i in range(a.passes):
    G = torch.eye(4, device=dev)[None].repeat(B, 1, 1).contiguous()
    ops.refine_iters(packed, f1, f2, ctx, g1, g2, depth.contiguous(), K, G, 1.0, a.iters, 3, workspace=ws, flags=flags)
torch.cuda.synchronize()
if a.time:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        G = torch.eye(4, device=dev)[None].repeat(B, 1, 1).contiguous()
        ops.refine_iters(packed, f1, f2, ctx, g1, g2, depth.contiguous(), K, G, 1.0, a.iters, 3, workspace=ws, flags=flags)
    e1.record(); torch.cuda.synchronize()
    print(f"ms per pass: {e0.elapsed_time(e1) / 10:.3f}  poses/s: {B * 10 / (e0.elapsed_time(e1) * 1e-3):.0f}")
print("ok", torch.isfinite(G).all().item())

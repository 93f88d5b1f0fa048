"""Times the host-buffer entry (b200pose_refine_iters_host) at the bench shape: pinned inputs, H2D + D2H inside the timed region.
both with the plain copy of the first descriptor map and with the depth-masked fetch.  usage: python tools/e2e_time.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnpose_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
sd = torch.load(os.path.join(os.path.dirname(__file__), "..", "rnnpose_b200", "weights", "gru_update.pth"), map_location="cpu")
packed = ops.pack_weights({k[len("update_block."):]: v.float() for k, v in sd.items()}, dev)
H, W, B = 240, 320, 32
g = torch.Generator().manual_seed(0)
pin = lambda t: t.contiguous().pin_memory()
f1 = pin(torch.randn(B, 256, H // 8, W // 8, generator=g)); f2 = pin(torch.randn(B, 256, H // 8, W // 8, generator=g))
ctx = pin(torch.empty(B, 256, H, W).uniform_(-0.2, 0.2))
g1 = pin(torch.nn.functional.normalize(torch.empty(B, 32, H, W).uniform_(-1, 1), dim=1))
g2 = pin(torch.nn.functional.normalize(torch.empty(B, 32, H, W).uniform_(-1, 1), dim=1))
yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
depth = pin((((yy - H / 2) ** 2 / (0.35 * H) ** 2 + (xx - W / 2) ** 2 / (0.3 * W) ** 2) < 1).float()[None].repeat(B, 1, 1) * 0.9)
K = pin(torch.tensor([[600.0, 0, W / 2], [0, 600.0, H / 2], [0, 0, 1]])[None].repeat(B, 1, 1))
G0 = torch.eye(4)[None].repeat(B, 1, 1)
Gh = pin(G0.clone())
Gd = G0.clone().to(dev)
ops.refine_iters(packed, f1.to(dev), f2.to(dev), ctx.to(dev), g1.to(dev), g2.to(dev), depth.to(dev), K.to(dev), Gd, 1.0, 4, 3)
torch.cuda.synchronize()
scratch = None
for mode in ("0", "1"):
    ops.set_option("sparse_g1", int(mode))
    res = []
    for i in range(6):
        Gh.copy_(G0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, scratch = ops.refine_iters_host(packed, f1, f2, ctx, g1, g2, depth, K, Gh, 1.0, 4, 3, scratch=scratch)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1))
    print(f"B200POSE_SPARSE_G1={mode}: ms per batch of {B}: median {sorted(res)[len(res) // 2]:.2f} min {min(res):.2f}; "
          f"max |host - device entry| = {(Gh.to(dev) - Gd).abs().max().item():.3e}")

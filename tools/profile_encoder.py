"""Minimal driver for ncu: the image encoder (b200pose_image_encoder) on `batch` crop pairs at the bench shape.
usage: python tools/profile_encoder.py [--batch 32] [--passes 2] [--time]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnpose_b200 import ops  # noqa: E402
from rnnpose_b200.assets import load_encoder_weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
H, W, B = 240, 320, a.batch
dev = torch.device("cuda:0")
packed = ops.encoder_pack_weights(load_encoder_weights(), dev)
g = torch.Generator(device="cpu").manual_seed(0)
x1 = (torch.rand(B, 3, H, W, generator=g) * 255).to(dev); x2 = (torch.rand(B, 3, H, W, generator=g) * 255).to(dev)
ws = ops._ws(ops._lib.lib().b200pose_encoder_workspace_bytes(B, H, W), dev)
for _ in range(a.passes):
    f1, f2 = ops.image_encoder(packed, x1, x2, workspace=ws)
torch.cuda.synchronize()
if a.time:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.image_encoder(packed, x1, x2, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    print(f"ms per batch of {B} pairs: {e0.elapsed_time(e1) / 10:.3f}")
print("ok", torch.isfinite(f1).all().item(), float(f1.abs().mean()))

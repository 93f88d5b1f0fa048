"""Per-layer timing of the tensor-core convolution variants (B200POSE_CONV_MODE) at the bench shape, warm and back to back:
each timed call is b200pose_conv_layer (memset + operand split + the convolution); the helpers are identical across modes, so
differences between modes are differences of the convolution kernel including its launch overhead.
usage: python tools/conv_layer_bench.py [--batch 32] [--modes 0,1,2,3,4] [--layers 0,1,3,4,5,6,7,8,9,10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnpose_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--modes", default="0,1,2,3,4")
ap.add_argument("--layers", default="0,1,3,4,5,6,7,8,9,10")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
sd = torch.load(os.path.join(os.path.dirname(__file__), "..", "rnnpose_b200", "weights", "gru_update.pth"), map_location="cpu")
packed = ops.pack_weights({k[len("update_block."):]: v.float() for k, v in sd.items()}, dev)
B, h, w = a.batch, 30, 40
P = B * h * w
names = {0: "C1", 1: "C2", 3: "F2", 4: "ENC", 5: "ZR1", 6: "Q1", 7: "ZR2", 8: "Q2", 9: "HEADS", 10: "MASK2"}
modes = [int(m) for m in a.modes.split(",")]
print("layer   " + "".join(f"mode{m:>2d}   " for m in modes) + "(us per call incl. helpers)")
for layer in [int(x) for x in a.layers.split(",")]:
    cin0, cin1, cout, kh, kw = ops.conv_layer_info(layer)
    in0 = torch.randn(P, cin0, device=dev)
    in1 = torch.randn(P, cin1, device=dev) if cin1 else None
    out = torch.empty(P, (cout + 3) // 4 * 4, device=dev)
    nb = ops._lib.lib().b200pose_conv_layer_workspace_bytes(B, h, w)
    ws = ops._ws(nb, dev)
    line = f"{names.get(layer, layer):6s}"
    for m in modes:
        ops.set_option("conv_mode", m)
        for _ in range(3):
            ops.conv_layer(packed, layer, in0, in1, B, h, w, flags=1, workspace=ws, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            ops.conv_layer(packed, layer, in0, in1, B, h, w, flags=1, workspace=ws, out=out)
        e1.record(); torch.cuda.synchronize()
        line += f" {e0.elapsed_time(e1) / a.reps * 1e3:8.1f}"
    print(line)

"""Experiment: one batch of 32 through b200pose_refine_iters vs two half batches on two streams (the HBM-bound kernels of one half
can overlap the tensor-bound chained launch of the other).  usage: python tools/split_streams.py [--batch 32] [--parts 2]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnpose_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--parts", type=int, default=2)
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--skew", type=int, default=1, help="delay stream k by k/parts of an iteration (a dummy sleep kernel) so the halves interleave")
a = ap.parse_args()
H, W, B = 240, 320, a.batch
dev = torch.device("cuda:0")
sd = torch.load(os.path.join(os.path.dirname(__file__), "..", "rnnpose_b200", "weights", "gru_update.pth"), map_location="cpu")
packed = ops.pack_weights({k[len("update_block."):]: v.float() for k, v in sd.items()}, dev)
f1 = torch.randn(B, 256, H // 8, W // 8, device=dev); f2 = torch.randn(B, 256, H // 8, W // 8, device=dev)
ctx = 0.1 * torch.randn(B, 256, H, W, device=dev)
g1 = torch.nn.functional.normalize(torch.randn(B, 32, H, W, device=dev), dim=1)
g2 = torch.nn.functional.normalize(torch.randn(B, 32, H, W, device=dev), dim=1)
yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
depth = ((((yy - H / 2) ** 2 / (0.35 * H) ** 2 + (xx - W / 2) ** 2 / (0.3 * W) ** 2) < 1).float()[None].repeat(B, 1, 1) * 0.9).contiguous()
K = torch.tensor([[600.0, 0, W / 2], [0, 600.0, H / 2], [0, 0, 1]], device=dev)[None].repeat(B, 1, 1).contiguous()


def run(parts, reps):
    n = B // parts
    wss = [ops.RefineWorkspace(n, H, W, dev) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    sl = [slice(k * n, (k + 1) * n) for k in range(parts)]
    views = [dict(f1=f1[s].contiguous(), f2=f2[s].contiguous(), ctx=ctx[s].contiguous(), g1=g1[s].contiguous(), g2=g2[s].contiguous(),
                  depth=depth[s].contiguous(), K=K[s].contiguous()) for s in sl]
    Gs = [None] * parts

    def once():
        cur = torch.cuda.current_stream()
        for k in range(parts):
            streams[k].wait_stream(cur)
            with torch.cuda.stream(streams[k]):
                if a.skew and k:
                    torch.cuda._sleep(int(k * 350_000 * 1.6 / parts))       # ~ k/parts of one iteration (clock cycles)
                G = torch.eye(4, device=dev)[None].repeat(n, 1, 1).contiguous()
                v = views[k]
                ops.refine_iters(packed, v["f1"], v["f2"], v["ctx"], v["g1"], v["g2"], v["depth"], v["K"], G, 1.0, a.iters, 3, workspace=wss[k])
                Gs[k] = G
        for k in range(parts):
            cur.wait_stream(streams[k])
    once(); once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, torch.cat(Gs)


t1, G1 = run(1, 10)
print(f"one stream, batch {B}: {t1:.3f} ms -> {B / t1 * 1e3:.0f} poses/s")
for parts in sorted({2, a.parts}):
    tp, Gp = run(parts, 10)
    print(f"{parts} streams x batch {B // parts} (skew {a.skew}): {tp:.3f} ms -> {B / tp * 1e3:.0f} poses/s; max |dG| vs one stream {float((Gp - G1).abs().max()):.2e}")

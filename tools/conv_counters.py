"""Where the conv kernels' warps wait inside the real fused loop (B200POSE_V2_DEBUG=16 clock counters, accumulated per layer
over one refine call at the bench shape).  usage: B200POSE_CONV_MODE=m python tools/conv_counters.py   (m with bit 4 set: the chained launch, per-layer counters)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnpose_b200 import ops  # noqa: E402

ops.set_option("conv_debug", 16)
CHAIN = (ops.get_option("conv_mode") & 16) != 0

dev = torch.device("cuda:0")
sd = torch.load(os.path.join(os.path.dirname(__file__), "..", "rnnpose_b200", "weights", "gru_update.pth"), map_location="cpu")
packed = ops.pack_weights({k[len("update_block."):]: v.float() for k, v in sd.items()}, dev)
H, W, B = 240, 320, 32
f1 = torch.randn(B, 256, H // 8, W // 8, device=dev); f2 = torch.randn(B, 256, H // 8, W // 8, device=dev)
ctx = 0.1 * torch.randn(B, 256, H, W, device=dev)
g1 = torch.nn.functional.normalize(torch.randn(B, 32, H, W, device=dev), dim=1)
g2 = torch.nn.functional.normalize(torch.randn(B, 32, H, W, device=dev), dim=1)
yy, xx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
depth = (((yy - H / 2) ** 2 / (0.35 * H) ** 2 + (xx - W / 2) ** 2 / (0.3 * W) ** 2) < 1).float()[None].repeat(B, 1, 1) * 0.9
K = torch.tensor([[600.0, 0, W / 2], [0, 600.0, H / 2], [0, 0, 1]], device=dev)[None].repeat(B, 1, 1).contiguous()
ws = ops.RefineWorkspace(B, H, W, dev)
L = ops._lib.lib()
L.b200pose_debug_conv_counters.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * (12 * 160 * 8))()


def call():
    G = torch.eye(4, device=dev)[None].repeat(B, 1, 1).contiguous()
    ops.refine_iters(packed, f1, f2, ctx, g1, g2, depth.contiguous(), K, G, 1.0, 4, 3, workspace=ws, flags=ops.FLAG_TENSOR_CORES)


call(); call()
L.b200pose_debug_conv_counters(None, 1)
call()
L.b200pose_debug_conv_counters(buf, 0)
v = torch.tensor(list(buf), dtype=torch.float64).view(12, 160, 8)
names = ["(other)", "C1", "C2", "F1", "F2", "ENC", "ZR1", "Q1", "ZR2", "Q2", "HEADS", "MASK2"]
if CHAIN:
    # conv_chain_kernel: one launch per recurrent iteration; counters per layer, summed over the units of a cluster
    print(f"conv_mode {ops.get_option('conv_mode')} (chained launch): clocks per launch and cluster (leader CTA), mean over clusters")
    print("layer   units | MMA loop  per unit | wait full  wait tmem | prod wait empty  dep wait | epi wait full  epi busy  busy/unit")
    tot = torch.zeros(8, dtype=torch.float64)
    for i, n in enumerate(names):
        x = v[i]
        act = x[x[:, 6] > 0]
        if act.numel() == 0:
            continue
        m = act.mean(0) / 4.0
        pe = x[x[:, 5] > 0].mean(0) / 4.0
        pp = x[(x[:, 3] > 0) | (x[:, 7] > 0)]
        pw = (pp.mean(0) / 4.0) if pp.numel() else torch.zeros(8, dtype=torch.float64)
        tot += torch.stack([m[0], m[1], m[2], pw[3], pe[4], pe[5], m[6], pw[7]])
        print(f"{n:7s} {m[6]:5.2f} | {m[0]:8.0f} {m[0] / max(m[6], 1e-9):9.0f} | {m[1]:9.0f} {m[2]:10.0f} | {pw[3]:15.0f} {pw[7]:9.0f} | "
              f"{pe[4]:13.0f} {pe[5]:9.0f} {pe[5] / max(m[6], 1e-9):10.0f}")
    print(f"sum     {tot[6]:5.1f} | {tot[0]:8.0f}           | {tot[1]:9.0f} {tot[2]:10.0f} | {tot[3]:15.0f} {tot[7]:9.0f} | {tot[4]:13.0f} {tot[5]:9.0f}")
    print(f"  (1 us = 1965 clocks; MMA-loop sum = {tot[0] / 1965:.0f} us per launch)")
    per_cta = v[:, :, 0].sum(0) / 4.0                       # MMA-loop clocks per leader CTA, summed over layers (and pre-sums)
    act = per_cta[per_cta > 0]
    units = (v[:, :, 6].sum(0) / 4.0)[per_cta > 0]
    print(f"  per-cluster MMA-loop total: min {act.min() / 1965:.0f}  mean {act.mean() / 1965:.0f}  max {act.max() / 1965:.0f} us; "
          f"units per cluster: min {units.min():.1f} mean {units.mean():.1f} max {units.max():.1f}")
print(f"mode {ops.get_option('conv_mode')}: clocks per launch, mean over the CTAs that issue MMAs (4 launches per layer + GRU pre-sums)")
print("layer   units stages | MMA-loop | wait full  wait tmem | prod wait empty | epi wait full  epi busy | loop clk/stage  epi busy/unit")
for i, n in enumerate(names):
    x = v[i]
    act = x[x[:, 0] > 0]
    if act.numel() == 0:
        continue
    m = act.mean(0) / 4.0
    pe = x[x[:, 5] > 0].mean(0) / 4.0
    print(f"{n:7s} {m[6]:5.1f} {m[7]:6.1f} | {m[0]:8.0f} | {m[1]:9.0f} {m[2]:10.0f} | {pe[3]:15.0f} | {pe[4]:13.0f} {pe[5]:9.0f} | {m[0] / max(m[7], 1):14.0f} {pe[5] / max(m[6], 1e-9):14.0f}")

# timeline of CTA 0 over the last call: entry -> dependency resolved (griddepcontrol.wait returned) -> exit
log = (C.c_ulonglong * (512 * 4))()
L.b200pose_debug_conv_log.argtypes = [C.c_void_p]
n = L.b200pose_debug_conv_log(log)
ev = sorted([(log[i * 4], log[i * 4 + 1], log[i * 4 + 2], int(log[i * 4 + 3])) for i in range(n)])
print(f"\nCTA 0 timeline of the conv launches ({n} launches), us relative to the first; one recurrent iteration shown")
t0 = ev[0][0]
prev_exit = None
shown = 0
for e in ev:
    if shown >= 30:
        break
    if e[0] - t0 < 1.6e6:      # skip the per-call part (pre-sum GEMMs) and the first iteration
        prev_exit = e[2]
        continue
    gap = (e[0] - prev_exit) / 1e3 if prev_exit else 0.0
    print(f"  {names[e[3]]:6s} entry {(e[0] - t0) / 1e3:9.1f}  blocked on previous kernel {(e[1] - e[0]) / 1e3:6.1f}  run {(e[2] - e[1]) / 1e3:6.1f}  "
          f"(entry - previous conv exit {gap:7.1f})")
    prev_exit = e[2]
    shown += 1

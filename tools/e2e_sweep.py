#!/usr/bin/env python
"""Times the host-buffer entry (b200pose_refine_iters_host2) for the ways of moving the context map: rows read in place from
pinned memory (threads = -1) vs texels gathered by T host threads; also the host gather alone.  One GPU.
Usage: python tools/e2e_sweep.py [--batch 32] [--threads -1,1,2,4,8,12,16] [--steps 6]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rnnpose_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--H", type=int, default=240)
    ap.add_argument("--W", type=int, default=320)
    ap.add_argument("--threads", default="-1,1,2,4,8,12,16")
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    B, H, W = a.batch, a.H, a.W
    inputs = bench.make_inputs(0, B, 8, H, W)
    keys = ("fmap1", "fmap2", "context", "geofea1", "geofea2", "depth", "K", "G0")
    host = {k: inputs[k].pin_memory() for k in keys}
    packed = ops.pack_weights(bench.load_weights(), dev)
    Gh = host["G0"].clone().pin_memory()
    staging = ops.host_staging(B, H, W)
    print(f"cpus available: {len(os.sched_getaffinity(0))}; batch {B} at {H}x{W}", flush=True)
    out = torch.empty(B, 256, (H // 8) * (W // 8), 4).pin_memory()
    for T in [int(t) for t in a.threads.split(",") if int(t) > 0]:
        ops.context_gather_texels(host["context"], threads=T, out=out)
        t0 = time.time()
        for _ in range(3):
            ops.context_gather_texels(host["context"], threads=T, out=out)
        dt = (time.time() - t0) / 3
        print(f"gather alone, {T:2d} threads: {dt * 1e3:7.2f} ms per batch ({B * 256 * H * W / dt / 1e9:.1f} GB/s of touched rows)", flush=True)
    scratch = None
    ref = None
    for T in [int(t) for t in a.threads.split(",")]:
        def step():
            nonlocal scratch
            Gh.copy_(host["G0"])
            _, scratch = ops.refine_iters_host(packed, host["fmap1"], host["fmap2"], host["context"], host["geofea1"], host["geofea2"],
                                               host["depth"], host["K"], Gh, 1.0, 4, 3, scratch=scratch,
                                               staging=staging if T >= 0 else None, threads=max(T, 0))
        step(); step()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
        dt = (time.time() - t0) / a.steps
        if ref is None:
            ref = Gh.clone()
        print(f"host entry, threads {T:3d}: {dt * 1e3:7.2f} ms per batch = {B / dt:7.0f} poses/s; identical to first variant: {torch.equal(ref, Gh)}", flush=True)


if __name__ == "__main__":
    main()

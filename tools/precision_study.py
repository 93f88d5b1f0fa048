#!/usr/bin/env python
"""How many fp16 terms does each update-block convolution need?  CPU study with the oracle (test infrastructure).

The tensor-core path computes a*w as a_hi*w_hi + a_hi*w_lo + a_lo*w_hi on fp16 halves (three MMAs, ~22 bits).  This script
emulates, inside oracle/refine_oracle.py::update_block, the cheaper products
    A : a_hi * (w_hi + w_lo)     activations rounded to fp16, weights 22-bit          (two MMAs, half the activation bytes)
    W : (a_hi + a_lo) * w_hi     weights rounded to fp16                              (two MMAs)
    AW: a_hi * w_hi              (one MMA)
per layer, and reports max |dSE3| of the final pose against the unrounded fp32 oracle on seeded synthetic scenes.
Usage: python tools/precision_study.py [--objects 6] [--iters 4] [--modes A,W,AW] [--keep encoder.convf1,...]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refine_oracle as O            # noqa: E402
from oracle import encoder_oracle as EO          # noqa: E402
from rnnpose_b200 import assets, synthetic as S  # noqa: E402


def split22(x):
    hi = x.half().float()
    lo = (x - hi).half().float()
    return hi + lo


def run(wts, inp, n_iters, n_lm, mode, keep):
    """mode in {None,'A','W','AW','3'}; layers whose name is in `keep` stay at three terms."""
    orig = F.conv2d

    def conv2d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
        name = names.get(id(w))
        m = mode
        if name is None or m is None:
            return orig(x, w, b, stride, padding, dilation, groups)
        if name in keep:
            m = "3"
        xa = x.half().float() if m in ("A", "AW") else split22(x)
        wa = w.half().float() if m in ("W", "AW") else split22(w)
        return orig(xa.double(), wa.double(), b.double() if b is not None else None, stride, padding, dilation, groups).float()

    names = {id(v): k[:-len(".weight")] for k, v in wts.items() if k.endswith(".weight")}
    F.conv2d = conv2d
    try:
        out = []
        for b in range(inp["depth"].shape[0]):
            r = O.refine_inner_loop(wts, inp["fmap1"][b:b + 1], inp["fmap2"][b:b + 1], inp["context"][b:b + 1],
                                    inp["geofea1"][b:b + 1], inp["geofea2"][b:b + 1], inp["depth"][b:b + 1], inp["K"][b:b + 1],
                                    inp["G0"][b:b + 1], 1.0, n_iters, n_lm)
            out.append(r["G"])
        return torch.cat(out)
    finally:
        F.conv2d = orig


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=6)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--lm", type=int, default=3)
    ap.add_argument("--H", type=int, default=240)
    ap.add_argument("--W", type=int, default=320)
    ap.add_argument("--modes", default="3,A,W,AW")
    ap.add_argument("--keep", default="")
    ap.add_argument("--occlude", action="store_true")
    a = ap.parse_args()
    torch.manual_seed(0)
    wts = assets.load_update_weights()
    ew = assets.load_encoder_weights()
    idx = list(range(a.first, a.first + a.objects))
    mb = S.make_batch(idx, a.H, a.W, occlude=a.occlude, with_images=True)
    with torch.no_grad():
        f1, f2 = EO.image_encoder(ew, mb["syn_img"], mb["obs_img"])
    inp = dict(mb)
    inp["fmap1"], inp["fmap2"] = f1, f2
    inp["G0"] = torch.eye(4)[None].repeat(a.objects, 1, 1)
    keep = set(k for k in a.keep.split(",") if k)
    with torch.no_grad():
        ref = run(wts, inp, a.iters, a.lm, None, keep)
        print(f"objects {idx}, {a.H}x{a.W}, {a.iters}x{a.lm}; keep(3 terms)={sorted(keep)}")
        print("pose change of the loop itself: max |G - G0| =", float((ref - inp['G0']).abs().max()))
        for m in a.modes.split(","):
            G = run(wts, inp, a.iters, a.lm, m, keep)
            d = (G - ref).abs().amax(dim=(1, 2))
            print(f"mode {m:>2}: max |dSE3| per object = {[f'{x:.2e}' for x in d.tolist()]}  max {float(d.max()):.3e}")


if __name__ == "__main__":
    main()

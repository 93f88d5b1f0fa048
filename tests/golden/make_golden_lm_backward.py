"""Generates tests/golden/lm_backward.npz by EXECUTING the reference's SE3Sequence.reprojction_optim(num_iters=1)
(geometry/transformation.py:265-316) under autograd on CPU and back-propagating a fixed linear loss of the output pose:
gradients with respect to target and weight flow through the reference's custom Cholesky backward (geometry/cholesky.py:19-28)
and its se3 exponential.  dL/d(delta) is tapped on the solver's output.  Inputs are those of tests/golden/lm.npz.
Case "clamp": almost no damping (EP_LMBDA = 1e-4) and targets shifted by 400 px, so that components of the raw update leave [-1, 1]
and the clamp blocks their gradient.  Run in the build container only:  python tests/golden/make_golden_lm_backward.py"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_harness as RH  # noqa: E402

RH.install_stubs()


def main():
    import geometry.transformation as GT
    from config.default import get_cfg
    from geometry.transformation import SE3Sequence
    g = np.load(os.path.join(HERE, "lm.npz"))
    T = lambda a: torch.from_numpy(np.asarray(a))
    out = {}
    orig = GT.cholesky_solve
    for name, ep, shift in (("plain", 100.0, 0.0), ("clamp", 1e-4, 400.0)):
        get_cfg("LM").EP_LMBDA = ep
        depth = T(g["depth"])[:, None]                                     # [B,1,H,W] = syn_depth + 1e-5
        target = (T(g["target"])[:, None] + shift).clone().requires_grad_(True)       # [B,1,H,W,2]
        weight = T(g["weight"])[:, None, :, :, None].clone().requires_grad_(True)     # [B,1,H,W,1]
        K = T(g["K"]); G = T(g["G_in"])[:, None]
        taps = []

        def tap(Hm, b):
            x = orig(Hm, b)
            x.retain_grad()
            taps.append(x)
            return x
        GT.cholesky_solve = tap
        Tn = SE3Sequence(matrix=G.clone()).reprojction_optim(target, weight, depth, K, num_iters=1)
        GT.cholesky_solve = orig
        Wr = torch.randn(Tn.G.shape, generator=torch.Generator().manual_seed(3))
        (Tn.G * Wr).sum().backward()
        out[f"{name}_ep"] = np.array(ep); out[f"{name}_shift"] = np.array(shift)
        out[f"{name}_delta"] = taps[0].detach()[:, 0].numpy()
        out[f"{name}_grad_delta"] = taps[0].grad[:, 0].numpy()
        out[f"{name}_grad_target"] = target.grad[:, 0].numpy()
        out[f"{name}_grad_weight"] = weight.grad[:, 0, :, :, 0].numpy()
        print(name, "delta", out[f"{name}_delta"], "\n  |grad_target|max", np.abs(out[f"{name}_grad_target"]).max(),
              "|grad_weight|max", np.abs(out[f"{name}_grad_weight"]).max())
    get_cfg("LM").EP_LMBDA = 100.0
    path = os.path.join(HERE, "lm_backward.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

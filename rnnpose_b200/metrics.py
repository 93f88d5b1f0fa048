"""Per-object pose metrics on torch tensors (any device): the quantities the reference's evaluator
computes per object (reference utils/eval_metric.py:102-192,306-339; utils/geometric.py:36-40) and
that the multi-GPU harness all-gathers (reference tools/train.py:724-741).  On CUDA tensors the arithmetic
runs in the library's own kernel (csrc/metrics.cu through b200pose_pose_metrics: SURVEY.md section 8(f)-3, the
replacement for the reference's thirdparty/nn extension); the torch formulation below is kept for CPU tensors
(host-side gloo tests) and as a second checker of that kernel."""
from __future__ import annotations

import math

import torch

METRIC_NAMES = ("add", "adds", "ang_err_deg", "trans_err", "proj2d_px", "ang_trace_deg",
                "add_lt_0p1d", "adds_lt_0p1d", "add_lt_0p02d", "adds_lt_0p02d", "add_lt_0p05d", "adds_lt_0p05d",
                "proj2d_lt_5px", "cm5deg5", "obj_index", "diameter")
COL = {n: i for i, n in enumerate(METRIC_NAMES)}
LINEMOD_K = ((572.4114, 0.0, 325.2611), (0.0, 573.57043, 242.04899), (0.0, 0.0, 1.0))   # data/linemod/linemod_config.py:23-25


def pose_metrics(T_pred: torch.Tensor, T_gt: torch.Tensor, pts: torch.Tensor, diameter: torch.Tensor,
                 obj_index: torch.Tensor, K: torch.Tensor = None) -> torch.Tensor:
    """T_pred, T_gt [B,4,4]; pts [B,N,3] model points; diameter [B]; K [3,3] or [B,3,3] (default linemod_K);
    returns [B, 16] float32 in the order of METRIC_NAMES."""
    if T_pred.is_cuda:
        from . import ops
        out = ops.pose_metrics(T_pred, T_gt, pts, diameter, K)
        out[:, COL["obj_index"]] = obj_index.to(out.dtype)
        return out
    return pose_metrics_torch(T_pred, T_gt, pts, diameter, obj_index, K)


def pose_metrics_torch(T_pred: torch.Tensor, T_gt: torch.Tensor, pts: torch.Tensor, diameter: torch.Tensor,
                       obj_index: torch.Tensor, K: torch.Tensor = None) -> torch.Tensor:
    """The same quantities with torch ops (any device)."""
    B = pts.shape[0]
    if K is None:
        K = torch.tensor(LINEMOD_K, dtype=pts.dtype, device=pts.device)
    K = K.to(pts)
    if K.dim() == 2:
        K = K[None].expand(B, 3, 3)
    Rp, tp = T_pred[:, :3, :3], T_pred[:, :3, 3]
    Rg, tg = T_gt[:, :3, :3], T_gt[:, :3, 3]
    pp = torch.einsum("bij,bnj->bni", Rp, pts) + tp[:, None]
    pg = torch.einsum("bij,bnj->bni", Rg, pts) + tg[:, None]
    add = (pp - pg).norm(dim=-1).mean(dim=1)                                   # eval_metric.py:173-174
    # eval_metric.py:167-171: for every ground-truth point the nearest predicted point (dim 1 = predicted points)
    adds = torch.cdist(pp, pg, compute_mode="donot_use_mm_for_euclid_dist").min(dim=1).values.mean(dim=1)
    up = torch.einsum("bij,bnj->bni", K, pp); ug = torch.einsum("bij,bnj->bni", K, pg)
    proj = (up[..., :2] / up[..., 2:] - ug[..., :2] / ug[..., 2:]).norm(dim=-1).mean(dim=1)   # eval_metric.py:102-110
    n = (Rg - Rp).reshape(B, -1).norm(dim=1)
    ang = 2 * torch.asin(torch.clamp(n / math.sqrt(8.0), max=1.0)) * (180.0 / math.pi)   # geometric.py:36-40
    trace = torch.clamp((Rp * Rg).sum(dim=(1, 2)), max=3.0)                    # eval_metric.py:184-186
    ang_tr = torch.acos((trace - 1.0) / 2.0) * (180.0 / math.pi)
    trans = (tp - tg).norm(dim=1)
    d = diameter.to(add)
    f = lambda c: c.to(add)
    cm5 = f((trans * 100 < 5) & (ang_tr < 5))                                  # eval_metric.py:181-192
    return torch.stack([add, adds, ang, trans, proj, ang_tr, f(add < 0.1 * d), f(adds < 0.1 * d), f(add < 0.02 * d),
                        f(adds < 0.02 * d), f(add < 0.05 * d), f(adds < 0.05 * d), f(proj < 5), cm5,
                        obj_index.to(add), d], dim=1).float()

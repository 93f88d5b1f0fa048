"""Per-object pose metrics on torch tensors (any device): the quantities the reference's evaluator
accumulates per object (reference utils/eval_metric.py:161-192,306-339; utils/geometric.py:36-40) and
that the multi-GPU harness all-gathers (reference tools/train.py:724-741).  On CUDA tensors the arithmetic
runs in the library's own kernel (csrc/metrics.cu through b200pose_pose_metrics: SURVEY.md section 8(f)-3, the
replacement for the reference's thirdparty/nn extension); the torch formulation below is kept for CPU tensors
(host-side gloo tests) and as the checker of that kernel."""
from __future__ import annotations

import math

import torch

METRIC_NAMES = ("add", "adds", "ang_err_deg", "trans_err", "add_lt_0p1d", "adds_lt_0p1d", "cm5deg5", "obj_index")


def pose_metrics(T_pred: torch.Tensor, T_gt: torch.Tensor, pts: torch.Tensor, diameter: torch.Tensor,
                 obj_index: torch.Tensor) -> torch.Tensor:
    """T_pred, T_gt [B,4,4]; pts [B,N,3] model points; diameter [B]; returns [B, 8] float32 in the
    order of METRIC_NAMES."""
    if T_pred.is_cuda:
        from . import ops
        out = ops.pose_metrics(T_pred, T_gt, pts, diameter)
        out[:, 7] = obj_index.to(out.dtype)
        return out
    return pose_metrics_torch(T_pred, T_gt, pts, diameter, obj_index)


def pose_metrics_torch(T_pred: torch.Tensor, T_gt: torch.Tensor, pts: torch.Tensor, diameter: torch.Tensor,
                       obj_index: torch.Tensor) -> torch.Tensor:
    """The same quantities with torch ops (any device)."""
    Rp, tp = T_pred[:, :3, :3], T_pred[:, :3, 3]
    Rg, tg = T_gt[:, :3, :3], T_gt[:, :3, 3]
    pp = torch.einsum("bij,bnj->bni", Rp, pts) + tp[:, None]
    pg = torch.einsum("bij,bnj->bni", Rg, pts) + tg[:, None]
    add = (pp - pg).norm(dim=-1).mean(dim=1)                                   # eval_metric.py:173-174
    adds = torch.cdist(pp, pg, compute_mode="donot_use_mm_for_euclid_dist").min(dim=2).values.mean(dim=1)                   # eval_metric.py:167-171
    n = (Rg - Rp).reshape(Rp.shape[0], -1).norm(dim=1)
    ang = 2 * torch.asin(torch.clamp(n / math.sqrt(8.0), max=1.0)) * (180.0 / math.pi)   # geometric.py:36-40
    trans = (tp - tg).norm(dim=1)
    thr = 0.1 * diameter
    cm5 = ((trans * 100 < 5) & (ang < 5)).float()                              # eval_metric.py:181-192
    return torch.stack([add, adds, ang, trans, (add < thr).float(), (adds < thr).float(), cm5,
                        obj_index.float()], dim=1).float()

"""Host-side mirror of the reference's pose container.

``SE3Sequence`` carries ``G`` of shape [B, K, 4, 4] (K = 1 on this path) exactly like the reference's
``geometry.transformation.SE3Sequence`` (reference geometry/transformation.py:65-225,230-320), with the members
the callers of the hot path touch: ``.G``, ``matrix()``, ``inv()``, ``__mul__``, ``copy()``, ``identity()``,
``shape()`` (consumers: reference utils/eval_metric.py:314-316, tools/eval.py:545, model/RNNPose.py:207-208).
The heavy members (``transform``, ``reprojction_optim``) live in the CUDA library and are reached through
:mod:`rnnpose_b200.ops`; ``reprojction_optim`` is provided here with the reference's signature for drop-in use.
"""
from __future__ import annotations

import torch


class SE3Sequence:
    def __init__(self, upsilon=None, matrix=None, so3=None, translation=None, eq="aijk,ai...k->ai...j", internal="matrix"):
        if internal != "matrix" or matrix is None:
            raise NotImplementedError("only the matrix representation is used on the refinement path")
        self.G = matrix
        self.eq = eq
        self.internal = internal

    # reference transformation.py:95-100
    def __mul__(self, other: "SE3Sequence") -> "SE3Sequence":
        return SE3Sequence(matrix=torch.matmul(self.G, other.G))

    # reference geometry/se3.py:194-209
    def inv(self) -> "SE3Sequence":
        G = self.G
        Rt = G[..., :3, :3].transpose(-1, -2)
        t = -torch.matmul(Rt, G[..., :3, 3:])
        out = torch.zeros_like(G)
        out[..., :3, :3] = Rt
        out[..., :3, 3:] = t
        out[..., 3, 3] = 1
        return SE3Sequence(matrix=out)

    def copy(self, stop_gradients: bool = False) -> "SE3Sequence":
        return SE3Sequence(matrix=self.G.detach() if stop_gradients else self.G)

    def identity(self) -> "SE3Sequence":
        B = self.G.shape[0]
        eye = torch.eye(4, dtype=self.G.dtype, device=self.G.device).repeat(B, 1, 1, 1)
        return SE3Sequence(matrix=eye)

    def identity_(self) -> None:
        self.G = torch.eye(4, device=self.G.device, dtype=self.G.dtype).repeat(*self.G.shape[:-2], 1, 1)

    def matrix(self, fill: bool = True) -> torch.Tensor:
        return self.G

    def shape(self):
        return (self.G.shape[0], self.G.shape[1])

    def reprojction_optim(self, target, weight, depth, intrinsics, num_iters=2, depth_img_coords=None,
                          ep_lmbda: float = 100.0, lm_lmbda: float = 1e-4) -> "SE3Sequence":
        """Reference signature (geometry/transformation.py:265-271): target [B,1,H,W,2], weight [B,1,H,W,1],
        depth [B,1,H,W] (= syn_depth + EPS), intrinsics [B,3,3].  Runs b200pose_lm_solve on the device."""
        from . import ops
        if depth_img_coords is not None:
            raise NotImplementedError("depth_img_coords is unused on the refinement path")
        G = self.G[:, 0].contiguous().clone()
        ops.lm_solve(depth[:, 0].contiguous().float(), target[:, 0].contiguous().float(),
                     weight[:, 0, :, :, 0].contiguous().float(), intrinsics.contiguous().float(), G, num_iters,
                     ep_lmbda=ep_lmbda, lm_lmbda=lm_lmbda)
        self.G = G[:, None]
        return SE3Sequence(matrix=self.G)

"""TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs; never by the product path).

CPU restatement of the reference's image feature encoder: ImageFeaEncoder.forward (reference model/CFNet.py:39-49) =
BasicEncoder(output_dim=256, norm_fn='instance') of thirdparty/raft/extractor.py:118-232 with ResidualBlock (:6-57), in
plain fp32 torch ops (F.conv2d + an explicit InstanceNorm).  The reference runs it under fp16 autocast on its GPU path; the
oracle is the CPU fp32 result.  Pinned by tests/golden/encoder.npz, the executed reference
(tests/golden/make_golden_encoder.py)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def instance_norm(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.InstanceNorm2d defaults (extractor.py:28-31,128): per (sample, channel) over H x W, biased variance, no affine."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def _conv(w: Dict[str, torch.Tensor], key: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv2d(x, w[key + ".weight"], w[key + ".bias"], stride=stride, padding=padding)


def residual_block(w, prefix: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """extractor.py:46-57."""
    y = torch.relu(instance_norm(_conv(w, prefix + ".conv1", x, stride, 1)))
    y = torch.relu(instance_norm(_conv(w, prefix + ".conv2", y, 1, 1)))
    if stride != 1:
        x = instance_norm(_conv(w, prefix + ".downsample.0", x, stride, 0))
    return torch.relu(x + y)


def image_encoder(w: Dict[str, torch.Tensor], image1: torch.Tensor, image2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """w: state dict of weights/img_fea_enc.pth ('fnet.*').  image1, image2 [B,3,H,W] -> fmaps [B,256,H/8,W/8]."""
    B = image1.shape[0]
    x = torch.cat([2 * (image1 / 255.0) - 1.0, 2 * (image2 / 255.0) - 1.0], dim=0)     # CFNet.py:42-43, extractor.py:196-198
    x = torch.relu(instance_norm(_conv(w, "fnet.conv1", x, 2, 3)))                       # :200-202
    for layer, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):                  # :205-208
        x = residual_block(w, f"fnet.{layer}.0", x, stride)
        x = residual_block(w, f"fnet.{layer}.1", x, 1)
    x = _conv(w, "fnet.conv2", x)                                                        # :217
    return x[:B], x[B:]

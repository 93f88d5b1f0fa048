"""Host-side logic of the host-buffer entry (no GPU needed): the gather of the context texels the 1/8 resample reads
(b200pose_context_gather_texels; reference op: F.interpolate(context_fea, scale_factor=1/8, mode='bilinear',
align_corners=True), model/CFNet.py:129)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rnnpose_b200 import ops


def taps(n_in, n_out):
    s = np.float32(n_in - 1) / np.float32(n_out - 1)
    f = (s * np.arange(n_out, dtype=np.float32)).astype(np.float32)
    i0 = np.minimum(f.astype(np.int32), n_in - 1)
    return f, i0, np.minimum(i0 + 1, n_in - 1)


@pytest.mark.parametrize("B,H,W,threads", [(1, 240, 320, 1), (2, 64, 96, 3), (3, 72, 104, 16), (1, 480, 640, 0), (5, 16, 24, 7)])
def test_context_gather_texels(B, H, W, threads):
    g = torch.Generator().manual_seed(H * W + B)
    ctx = torch.randn(B, 256, H, W, generator=g)
    tex = ops.context_gather_texels(ctx, threads=threads)
    h, w = H // 8, W // 8
    fy, y0, y1 = taps(H, h)
    fx, x0, x1 = taps(W, w)
    c = ctx.numpy()
    ref = np.stack([c[:, :, y0][:, :, :, x0], c[:, :, y0][:, :, :, x1], c[:, :, y1][:, :, :, x0], c[:, :, y1][:, :, :, x1]], -1)
    assert np.array_equal(ref.reshape(B, 256, h * w, 4), tex.numpy())
    # the texels are everything the reference's resample needs
    ly = torch.from_numpy(fy - y0.astype(np.float32)).view(1, 1, h, 1)
    lx = torch.from_numpy(fx - x0.astype(np.float32)).view(1, 1, 1, w)
    t = tex.view(B, 256, h, w, 4)
    out = (t[..., 0] * (1 - lx) + t[..., 1] * lx) * (1 - ly) + (t[..., 2] * (1 - lx) + t[..., 3] * lx) * ly
    want = F.interpolate(ctx, scale_factor=1 / 8, mode="bilinear", align_corners=True)
    assert (out - want).abs().max().item() < 2e-6


def test_context_gather_texels_rejects_bad_arguments():
    with pytest.raises(ValueError):
        ops.context_gather_texels(torch.zeros(1, 128, 64, 64))
    with pytest.raises(ValueError):
        ops.context_gather_texels(torch.zeros(1, 256, 64, 64, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        ops.context_gather_texels(torch.zeros(1, 256, 8, 8))           # h = w = 1: the resample is undefined

"""Every convolution of the update block, one at a time, on both code paths (exact fp32 FFMA and the
tcgen05 fp16-hi/lo tensor-core path) against torch's fp32 conv2d on the CPU (what the reference runs:
thirdparty/raft/update.py nn.Conv2d layers with the shipped weights)."""
import pytest
import torch
import torch.nn.functional as F

from rnnpose_b200 import synthetic as S
from tests.util import load_update_weights

pytestmark = pytest.mark.gpu

# layer id -> (weight keys (fused along Cout), padding)
LAYERS = {
    0: (["encoder.convc1"], 0), 1: (["encoder.convc2"], 1), 3: (["encoder.convf2"], 1), 4: (["encoder.conv"], 1),
    5: (["gru.convz1", "gru.convr1"], (0, 2)), 6: (["gru.convq1"], (0, 2)),
    7: (["gru.convz2", "gru.convr2"], (2, 0)), 8: (["gru.convq2"], (2, 0)),
    9: (["flow_head.conv1", "mask.0"], 1), 10: (["mask.2"], 0),
}


def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def to_pxc(x):
    B, C, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(B * h * w, C).contiguous()


def from_pxc(x, B, h, w):
    return x.view(B, h, w, -1).permute(0, 3, 1, 2).contiguous()


@pytest.fixture(scope="module")
def ops():
    from rnnpose_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def packed(ops):
    return ops.pack_weights(load_update_weights(), dev())


@pytest.mark.parametrize("flags", [0, 1], ids=["fp32", "tcgen05"])
@pytest.mark.parametrize("shape", [(2, 9, 12), (1, 30, 40), (3, 16, 20)], ids=["9x12", "30x40", "16x20"])
@pytest.mark.parametrize("layer", sorted(LAYERS))
def test_conv_layer(ops, packed, layer, shape, flags):
    wts = load_update_weights()
    keys, pad = LAYERS[layer]
    W = torch.cat([wts[k + ".weight"] for k in keys], 0)
    bias = torch.cat([wts[k + ".bias"] for k in keys], 0)
    B, h, w = shape
    cin0, cin1, cout, kh, kw = ops.conv_layer_info(layer)
    assert W.shape == (cout, cin0 + cin1, kh, kw)
    x = S.hash_features((B, cin0 + cin1, h, w), 100 + layer, 1.5)
    ref = F.conv2d(x, W, bias, padding=pad)
    in0 = to_pxc(x[:, :cin0]).to(dev())
    in1 = to_pxc(x[:, cin0:]).to(dev()) if cin1 else None
    out = ops.conv_layer(packed, layer, in0, in1, B, h, w, flags=flags)
    torch.cuda.synchronize()
    got = from_pxc(out[:, :cout].contiguous(), B, h, w).cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    # fp32 path: plain fp32 rounding noise.  tcgen05 path: operands are exact to 2^-22, but the tensor core adds each
    # 16-deep partial dot product into the fp32 TMEM accumulator with truncation, so the error grows with the number
    # of accumulation steps (3 * K/16, up to 432 here): observed <= 1.5e-5 of the output scale.
    tol = (3e-6 if flags == 0 else 4e-5) * scale + 1e-5
    assert err <= tol, f"layer {layer} flags {flags}: max err {err} (scale {scale})"


@pytest.mark.parametrize("flags", [0, 1], ids=["fp32", "tcgen05"])
def test_convf1_via_im2col(ops, packed, flags):
    """encoder.convf1 (7x7, Cin=2) is lowered to a 98-wide im2col + 1x1 GEMM (layer id 2)."""
    wts = load_update_weights()
    B, h, w = 2, 11, 13
    flow = S.hash_features((B, 2, h, w), 7, 4.0)
    ref = F.conv2d(flow, wts["encoder.convf1.weight"], wts["encoder.convf1.bias"], padding=3)
    col = F.unfold(flow, 7, padding=3).view(B, 2, 49, h, w).permute(0, 2, 1, 3, 4).reshape(B, 98, h, w)   # k = tap*2 + c
    colp = torch.zeros(B * h * w, 112)                      # pipeline pitch of the im2col buffer (16-byte rows)
    colp[:, :98] = to_pxc(col)
    out = ops.conv_layer(packed, 2, colp.to(dev()), None, B, h, w, flags=flags)
    got = from_pxc(out[:, :128].contiguous(), B, h, w).cpu()
    assert (got - ref).abs().max().item() <= (3e-6 if flags == 0 else 4e-5) * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("shape", [(2, 30, 40), (1, 16, 20), (2, 60, 80)], ids=["30x40", "16x20", "60x80"])
def test_corr_pyramid_tensor_core_vs_oracle(ops, shape):
    """All-pairs volume on tcgen05 (fp16 hi/lo) + pooling against the fp32 restatement of CorrBlock.__init__."""
    from oracle import refine_oracle as O
    B, h, w = shape
    f1 = S.hash_features((B, 256, h, w), 1).to(dev()); f2 = S.hash_features((B, 256, h, w), 2).to(dev())
    lv = ops.pyramid_level_views(ops.corr_pyramid_tc(f1, f2), B, h, w)
    ref = O.corr_pyramid(f1.cpu(), f2.cpu())
    for l in range(4):
        assert lv[l].shape == ref[l].shape
        err = (lv[l].cpu() - ref[l]).abs().max().item()
        assert err <= 2e-5 * ref[l].abs().max().item() + 1e-5, (l, err)


# Second-generation tensor-core kernel (conv_umma2_kernel): B200POSE_CONV_MODE bit 0 = CTA pairs (cta_group::2, only taken
# when the problem has at least one tile per SM), bit 1 = vertical-tap reuse of the activation box.  Shapes: few tiles
# (every tile split), 320 M tiles (B=32 at 30x40: several persistent rounds + split tail) and an odd tile count (153: the
# last pair has a dummy half).
_REF_CACHE = {}


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["pair", "vreuse", "pair+vreuse"])
@pytest.mark.parametrize("layer", sorted(LAYERS))
@pytest.mark.parametrize("shape", [(2, 9, 12), (32, 30, 40), (51, 16, 20)], ids=["2x9x12", "32x30x40", "51x16x20"])
def test_conv_layer_second_generation(ops, packed, layer, shape, mode):
    wts = load_update_weights()
    keys, pad = LAYERS[layer]
    W = torch.cat([wts[k + ".weight"] for k in keys], 0)
    bias = torch.cat([wts[k + ".bias"] for k in keys], 0)
    B, h, w = shape
    cin0, cin1, cout, kh, kw = ops.conv_layer_info(layer)
    if (layer, shape) not in _REF_CACHE:                    # the CPU reference is shared by the three kernel variants
        x = S.hash_features((B, cin0 + cin1, h, w), 200 + layer, 1.5)
        _REF_CACHE.clear()
        _REF_CACHE[(layer, shape)] = (x, F.conv2d(x, W, bias, padding=pad))
    x, ref = _REF_CACHE[(layer, shape)]
    in0 = to_pxc(x[:, :cin0]).to(dev())
    in1 = to_pxc(x[:, cin0:]).to(dev()) if cin1 else None
    with ops.options(conv_mode=mode):
        out = ops.conv_layer(packed, layer, in0, in1, B, h, w, flags=1)
        torch.cuda.synchronize()
    got = from_pxc(out[:, :cout].contiguous(), B, h, w).cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= 4e-5 * scale + 1e-5, f"layer {layer} mode {mode}: max err {err} (scale {scale})"

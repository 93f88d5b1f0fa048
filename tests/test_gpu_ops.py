"""Parity of every CUDA entry point (through the C-ABI) with the oracle and with the golden vectors of
the executed reference.  Run on the B200 box:  python -m pytest tests -m gpu"""
import numpy as np
import pytest
import torch

from oracle import refine_oracle as O
from rnnpose_b200 import synthetic as S
from tests.util import golden, load_update_weights

pytestmark = pytest.mark.gpu


def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.asarray(a))


def to_pxc(x):      # [B,C,h,w] -> [B*h*w, C]
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


@pytest.mark.parametrize("variant", [(2, 1), (1, 1), (0, 0)], ids=["row-lookup+pool3", "window-lookup+pool3", "round1-kernels"])
def test_corr_pyramid_and_lookup_golden(ops, libopt, variant):
    libopt("lookup_mode", variant[0]); libopt("pool_mode", variant[1])
    g = golden("corr_lookup.npz")
    B, D, h, w, s1, s2 = [int(v) for v in g["meta"]]
    f1 = S.hash_features((B, D, h, w), s1).to(dev()); f2 = S.hash_features((B, D, h, w), s2).to(dev())
    pyr = ops.corr_pyramid(f1, f2)
    lv = ops.pyramid_level_views(pyr, B, h, w)
    assert [tuple(v.shape[-2:]) for v in lv] == [(17, 22), (8, 11), (4, 5), (2, 2)]
    for l, key in ((0, "pyr0"), (1, "pyr1"), (3, "pyr3")):
        torch.testing.assert_close(lv[l].cpu(), T(g[key]), rtol=1e-5, atol=3e-6)
    ref = O.corr_pyramid(f1.cpu(), f2.cpu())
    for l in range(4):
        torch.testing.assert_close(lv[l].cpu(), ref[l], rtol=1e-5, atol=3e-6)
    coords = T(g["coords"]).to(dev())
    out = ops.corr_lookup(pyr, to_pxc(coords), B, h, w)
    assert out.shape == (B * h * w, 328)
    assert torch.all(out[:, 324:] == 0)
    torch.testing.assert_close(from_pxc(out[:, :324].contiguous(), B, h, w).cpu(), T(g["out"]), rtol=1e-5, atol=5e-6)


def test_corr_lookup_window_kernel_edge_coordinates(ops):
    """The shared-memory window lookup against the round-1 per-sample kernel and the oracle on coordinates that stress the
    window logic: integers, just below integers, far outside the map on every side, huge and non-finite values."""
    B, D, h, w = 1, 64, 16, 20
    f1 = S.hash_features((B, D, h, w), 61).to(dev()); f2 = S.hash_features((B, D, h, w), 62).to(dev())
    pyr = ops.corr_pyramid(f1, f2)
    P = h * w
    base = S.hash_features((B, 2, h, w), 63, 12.0) + torch.tensor([10.0, 8.0]).view(1, 2, 1, 1)
    c = to_pxc(base).clone()
    c[0] = torch.tensor([3.0, 5.0]); c[1] = torch.tensor([2.9999998, 4.9999995]); c[2] = torch.tensor([-30.0, 7.0])
    c[3] = torch.tensor([7.0, 400.0]); c[4] = torch.tensor([1e9, -1e9]); c[5] = torch.tensor([19.0, 15.0])
    c[6] = torch.tensor([-4.5, -4.5]); c[7] = torch.tensor([23.5, 19.5]); c[8] = torch.tensor([0.0, 0.0])
    cd = c.to(dev())
    with ops.options(lookup_mode=0):
        old = ops.corr_lookup(pyr, cd, B, h, w).cpu()
    with ops.options(lookup_mode=1):
        new = ops.corr_lookup(pyr, cd, B, h, w).cpu()
    assert torch.all(new[:, 324:] == 0)
    torch.testing.assert_close(new, old, rtol=1e-5, atol=5e-6)
    with ops.options(lookup_mode=2):                                   # one thread per window row
        row = ops.corr_lookup(pyr, cd, B, h, w).cpu()
    assert torch.all(row[:, 324:] == 0)
    torch.testing.assert_close(row, old, rtol=1e-5, atol=5e-6)
    ref = O.corr_lookup(O.corr_pyramid(f1.cpu(), f2.cpu()), from_pxc(c, B, h, w))
    torch.testing.assert_close(from_pxc(new[:, :324].contiguous(), B, h, w), ref, rtol=1e-5, atol=5e-6)
    # non-finite coordinates: the reference's grid_sample yields zeros there; both kernels must agree and stay finite
    c2 = c.clone(); c2[9] = torch.tensor([float("nan"), 3.0]); c2[10] = torch.tensor([float("inf"), 3.0])
    for mode in (1, 2):
        with ops.options(lookup_mode=mode):
            nf = ops.corr_lookup(pyr, c2.to(dev()), B, h, w).cpu()
        assert torch.all(nf[9:11] == 0) and torch.isfinite(nf).all()


def test_corr_pyramid_full_size_vs_oracle(ops):
    B, D, h, w = 2, 256, 30, 40
    f1 = S.hash_features((B, D, h, w), 11).to(dev()); f2 = S.hash_features((B, D, h, w), 12).to(dev())
    lv = ops.pyramid_level_views(ops.corr_pyramid(f1, f2), B, h, w)
    ref = O.corr_pyramid(f1.cpu(), f2.cpu())
    assert [tuple(v.shape[-2:]) for v in lv] == [(30, 40), (15, 20), (7, 10), (3, 5)]
    for l in range(4):
        torch.testing.assert_close(lv[l].cpu(), ref[l], rtol=2e-5, atol=2e-5)


def test_context_init(ops):
    B, H, W = 2, 128, 160
    ctx = S.hash_features((B, 256, H, W), 21, 0.5).to(dev())
    net, xbuf = ops.context_init(ctx)
    rnet, rinp = O.context_init(ctx.cpu(), W // 8)
    torch.testing.assert_close(from_pxc(net, B, H // 8, W // 8).cpu(), rnet, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(from_pxc(xbuf[:, :128].contiguous(), B, H // 8, W // 8).cpu(), rinp, rtol=1e-5, atol=2e-6)


def test_flow_init(ops):
    mb = S.make_batch([0, 3], 128, 160, with_images=False)
    depth = mb["depth"][:, 0].contiguous()
    G = O.se3_exp(torch.tensor([[0.01, -0.02, 0.015, 0.02, -0.01, 0.03], [-0.03, 0.01, 0.0, -0.015, 0.025, -0.02]]))
    c1, fl = ops.flow_init(depth.to(dev()), mb["K"].to(dev()), G.to(dev()))
    ref = O.flow_init_lowres(depth + O.EPS_DEPTH, mb["K"], G)
    B, h, w = 2, 16, 20
    torch.testing.assert_close(from_pxc(fl, B, h, w).cpu(), ref, rtol=1e-4, atol=2e-5)
    u, v = O.pixel_grid(h, w)
    torch.testing.assert_close(from_pxc(c1, B, h, w).cpu(), ref + torch.stack([u, v])[None], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("flags", [0, 1], ids=["fp32", "tcgen05"])
def test_update_block_golden(ops, packed, flags):
    g = golden("update_block.npz")
    B, h, w, s1, s2, s3, s4 = [int(v) for v in g["meta"]]
    net = torch.tanh(S.hash_features((B, 128, h, w), s1)); inp = torch.relu(S.hash_features((B, 128, h, w), s2))
    corr = S.hash_features((B, 324, h, w), s3, 2.0); flow = S.hash_features((B, 2, h, w), s4, 4.0)
    P = B * h * w
    netd = to_pxc(net).to(dev())
    xbuf = torch.zeros(P, 256, device=dev()); xbuf[:, :128] = to_pxc(inp).to(dev())
    corrd = torch.zeros(P, 328, device=dev()); corrd[:, :324] = to_pxc(corr).to(dev())
    u, v = O.pixel_grid(h, w)
    flowd = to_pxc(flow).to(dev())
    coords1 = (to_pxc(torch.stack([u, v])[None].expand(B, 2, h, w)).to(dev()) + flowd).contiguous()
    mask, dflow = ops.update_block(packed, netd, xbuf, corrd, coords1, flowd, B, h, w, flags=flags)
    torch.testing.assert_close(from_pxc(netd, B, h, w).cpu(), T(g["net_out"]), rtol=1e-4, atol=(3e-5 if flags == 0 else 2e-4))
    torch.testing.assert_close(from_pxc(mask, B, h, w).cpu(), T(g["mask"]), rtol=1e-4, atol=(3e-5 if flags == 0 else 2e-4))
    torch.testing.assert_close(from_pxc(dflow, B, h, w).cpu(), T(g["dflow"]), rtol=1e-4, atol=(3e-5 if flags == 0 else 2e-4))
    # flow out = (coords1 + dflow) - coords0
    torch.testing.assert_close(from_pxc(flowd, B, h, w).cpu(), flow + T(g["dflow"]), rtol=1e-4,
                               atol=(5e-5 if flags == 0 else 5e-4))


@pytest.mark.parametrize("flags", [0, 1], ids=["fp32", "tcgen05"])
def test_update_block_ragged_tile(ops, packed, flags):
    """P not a multiple of the 128-pixel tile and a 1-row map: exercises the M-tail and the halo predicates."""
    wts = load_update_weights()
    for (B, h, w) in ((1, 1, 5), (3, 7, 13)):
        net = torch.tanh(S.hash_features((B, 128, h, w), 31)); inp = torch.relu(S.hash_features((B, 128, h, w), 32))
        corr = S.hash_features((B, 324, h, w), 33, 2.0); flow = S.hash_features((B, 2, h, w), 34, 4.0)
        rn, rm, rd = O.update_block(wts, net, inp, corr, flow)
        P = B * h * w
        netd = to_pxc(net).to(dev())
        xbuf = torch.zeros(P, 256, device=dev()); xbuf[:, :128] = to_pxc(inp).to(dev())
        corrd = torch.zeros(P, 328, device=dev()); corrd[:, :324] = to_pxc(corr).to(dev())
        flowd = to_pxc(flow).to(dev())
        coords1 = flowd.clone()
        mask, dflow = ops.update_block(packed, netd, xbuf, corrd, coords1, flowd, B, h, w, flags=flags)
        torch.testing.assert_close(from_pxc(netd, B, h, w).cpu(), rn, rtol=1e-4, atol=(3e-5 if flags == 0 else 2e-4))
        torch.testing.assert_close(from_pxc(mask, B, h, w).cpu(), rm, rtol=1e-4, atol=(3e-5 if flags == 0 else 2e-4))
        torch.testing.assert_close(from_pxc(dflow, B, h, w).cpu(), rd, rtol=1e-4, atol=(3e-5 if flags == 0 else 2e-4))


def test_upsample_golden(ops):
    g = golden("upsample.npz")
    B, h, w, s1, s2 = [int(v) for v in g["meta"]]
    fl = S.hash_features((B, 2, h, w), s1, 3.0); mk = S.hash_features((B, 576, h, w), s2, 2.0)
    fu, tgt, wgt = ops.upsample_weight(to_pxc(fl).to(dev()), to_pxc(mk).to(dev()), None, None, None, 1.0, B, 8 * h, 8 * w)
    assert wgt is None
    torch.testing.assert_close(fu.cpu(), T(g["out"]), rtol=1e-5, atol=1e-5)
    u, v = O.pixel_grid(8 * h, 8 * w)
    torch.testing.assert_close(tgt.cpu(), torch.stack([T(g["out"])[:, 0] + u, T(g["out"])[:, 1] + v], -1), rtol=1e-5, atol=2e-5)


def test_upsample_weight_vs_oracle(ops):
    B, C, h, w = 2, 32, 6, 9
    H, W = 8 * h, 8 * w
    fl = S.hash_features((B, 2, h, w), 41, 2.5); fl[0, :, 0, 0] = torch.tensor([-9.0, -7.0]); fl[1, :, -1, -1] = 11.0
    mk = S.hash_features((B, 576, h, w), 42, 2.0)
    g1 = S.hash_features((B, C, H, W), 43); g1 = g1 / g1.norm(dim=1, keepdim=True)
    g2 = S.hash_features((B, C, H, W), 44); g2 = g2 / g2.norm(dim=1, keepdim=True)
    depth = (S.hash_features((B, H, W), 45) > -0.3).float() * 0.9
    fu, tgt, wgt = ops.upsample_weight(to_pxc(fl).to(dev()), to_pxc(mk).to(dev()), g1.to(dev()), g2.to(dev()),
                                       depth.to(dev()), 0.7, B, H, W)
    rfu = O.convex_upsample(fl, mk)
    u, v = O.pixel_grid(H, W)
    rt = torch.stack([rfu[:, 0] + u, rfu[:, 1] + v], -1)
    torch.testing.assert_close(fu.cpu(), rfu, rtol=1e-5, atol=1e-5)
    rw = O.corr_weight(g1, g2, rt, depth, 0.7)
    torch.testing.assert_close(wgt.cpu(), rw, rtol=1e-4, atol=1e-5)
    assert torch.all(wgt.cpu()[depth == 0] == 0)


def test_upsample_weight_kernel_builds_agree(ops, libopt):
    """The dense kernel's builds (option upsample_variant: compile-time plane stride, staged channel batches) are the same
    arithmetic in the same order: bit-identical outputs at the crop size the specialised builds exist for, and against the
    oracle there."""
    B, C, H, W = 2, 32, 240, 320
    h, w = H // 8, W // 8
    fl = S.hash_features((B, 2, h, w), 141, 2.5); mk = S.hash_features((B, 576, h, w), 142, 2.0)
    g1 = S.hash_features((B, C, H, W), 143); g1 = g1 / g1.norm(dim=1, keepdim=True)
    g2 = S.hash_features((B, C, H, W), 144); g2 = g2 / g2.norm(dim=1, keepdim=True)
    depth = (S.hash_features((B, H, W), 145) > 0.2).float() * 0.9
    args = (to_pxc(fl).to(dev()), to_pxc(mk).to(dev()), g1.to(dev()), g2.to(dev()), depth.to(dev()), 0.7, B, H, W)
    libopt("upsample_variant", 0)
    fu0, tgt0, wgt0 = [t.clone() for t in ops.upsample_weight(*args)]
    rfu = O.convex_upsample(fl, mk)
    u, v = O.pixel_grid(H, W)
    rw = O.corr_weight(g1, g2, torch.stack([rfu[:, 0] + u, rfu[:, 1] + v], -1), depth, 0.7)
    torch.testing.assert_close(wgt0.cpu(), rw, rtol=1e-4, atol=2e-5)
    for var in range(1, 7):
        libopt("upsample_variant", var)
        fu, tgt, wgt = ops.upsample_weight(*args)
        assert torch.equal(fu, fu0) and torch.equal(tgt, tgt0) and torch.equal(wgt, wgt0), f"variant {var}"


def test_weight_golden_via_identity_mask(ops):
    """G5 fixture: drive the fused kernel with a mask that selects the centre tap (softmax -> one-hot) so
    that target = grid + 8*flow, then compare the weight with the executed reference."""
    g = golden("weight.npz")
    B, C, H, W, s1, s2, s3, s4 = [int(v) for v in g["meta"]]
    h, w = H // 8, W // 8
    g1 = S.hash_features((B, C, H, W), s1); g1 = g1 / g1.norm(dim=1, keepdim=True)
    g2 = S.hash_features((B, C, H, W), s2); g2 = g2 / g2.norm(dim=1, keepdim=True)
    depth = (S.hash_features((B, 1, H, W), s4) > -0.3).float()[:, 0] * 0.9
    # oracle weight on the kernel's own target (flow piecewise constant over 8x8 blocks)
    fl = S.hash_features((B, 2, h, w), 77, 0.6)
    mk = torch.full((B, 9, 64, h, w), -200.0); mk[:, 4] = 200.0
    fu, tgt, wgt = ops.upsample_weight(to_pxc(fl).to(dev()), to_pxc(mk.reshape(B, 576, h, w)).to(dev()), g1.to(dev()),
                                       g2.to(dev()), depth.to(dev()), float(g["sigma"]), B, H, W)
    rw = O.corr_weight(g1, g2, tgt.cpu(), depth, float(g["sigma"]))
    torch.testing.assert_close(wgt.cpu(), rw, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n", [1, 3])
def test_lm_golden(ops, n):
    g = golden("lm.npz")
    depth = T(g["depth"])                               # = syn_depth + EPS, as reprojction_optim receives it
    tgt, wgt, K = T(g["target"]), T(g["weight"]), T(g["K"])
    G = T(g["G_in"]).clone().to(dev())
    G, Ho, bo, do = ops.lm_solve(depth.to(dev()), tgt.to(dev()), wgt.to(dev()), K.to(dev()), G, n, taps=True)
    eye = torch.eye(6, dtype=torch.float64)
    for it in range(n):
        Hd = Ho[it].cpu() + 100.0 * eye + 1e-4 * Ho[it].cpu() * eye
        Href = T(g[f"Hd_{n}"][it]); bref = T(g[f"b_{n}"][it])
        # step 0 sees bit-identical inputs; later steps inherit the ~1e-7 fp32 rounding of exp(delta) G, which
        # shows up relative to the matrix norm in the cancelling off-diagonal sums
        slack = 0.0 if it == 0 else 2e-6
        torch.testing.assert_close(Hd, Href, rtol=1e-7, atol=1e-3 + slack * Href.abs().max().item())
        # b = J^T W r shrinks towards 0 as the steps converge: its error scales with |H| * |dG|, not with |b|
        torch.testing.assert_close(bo[it].cpu(), bref, rtol=1e-6, atol=1e-2 + slack * Href.abs().max().item())
    torch.testing.assert_close(G.cpu(), T(g[f"G_out_{n}"]), rtol=0, atol=5e-6)


def test_lm_near_plane_and_nan(ops):
    g = golden("lm.npz")
    depth = T(g["depth"])
    tgt, wgt, K = T(g["target"]), T(g["weight"]), T(g["K"])
    G = T(g["G_in_near"]).clone().to(dev())
    G = ops.lm_solve(depth.to(dev()), tgt.to(dev()), wgt.to(dev()), K.to(dev()), G, 2)
    torch.testing.assert_close(G.cpu(), T(g["G_out_near"]), rtol=0, atol=1e-5)
    # NaN weight in sample 0: zero update for sample 0 (cholesky.py:42-45), sample 1 unaffected
    w2 = wgt.clone(); w2[0, 5, 5] = float("nan")
    G0 = T(g["G_in"]).clone()
    Gd, Ho, bo, do = ops.lm_solve(depth.to(dev()), tgt.to(dev()), w2.to(dev()), K.to(dev()), G0.clone().to(dev()), 1, taps=True)
    assert torch.all(do[0, 0] == 0)
    torch.testing.assert_close(Gd[0].cpu(), G0[0], rtol=0, atol=1e-7)
    torch.testing.assert_close(Gd[1].cpu(), T(g["G_out_1"])[1], rtol=0, atol=5e-6)


def test_lm_vs_oracle_random(ops):
    mb = S.make_batch([1, 2, 4], 128, 160, with_images=False)
    depth = mb["depth"][:, 0].contiguous()
    B, H, W = depth.shape
    u, v = O.pixel_grid(H, W)
    tgt = torch.stack([u, v], -1)[None].repeat(B, 1, 1, 1) + S.hash_features((B, H, W, 2), 51, 1.5)
    wgt = torch.rand(B, H, W, generator=torch.Generator().manual_seed(5)) * (depth > 0)
    G = O.se3_exp(S.hash_features((B, 6), 52, 0.02))
    Gr = G.clone()
    for _ in range(3):
        Gr, _, Hm, bv = O.lm_step(depth + O.EPS_DEPTH, tgt, wgt, mb["K"], Gr)
    Gd = ops.lm_solve(depth.to(dev()), tgt.to(dev()), wgt.to(dev()), mb["K"].to(dev()), G.clone().to(dev()), 3,
                      depth_offset=1e-5)
    torch.testing.assert_close(Gd.cpu(), Gr, rtol=0, atol=5e-6)


# ----------------------------------------------------------------------------- f4: backward of one LM step
@pytest.mark.parametrize("case", ["plain", "clamp"])
def test_lm_backward_reference_autograd_golden(ops, case):
    """b200pose_lm_backward against the reference's autograd through reprojction_optim(num_iters=1) (lm_backward.npz): plain
    damping, and a case whose raw update leaves [-1, 1] so that the clamp blocks part of the gradient."""
    g = golden("lm.npz"); gb = golden("lm_backward.npz")
    d = dev()
    target = (T(g["target"]) + float(gb[f"{case}_shift"])).contiguous()
    ep = float(gb[f"{case}_ep"])
    gt, gw = ops.lm_backward(T(g["depth"]).to(d), target.to(d), T(g["weight"]).contiguous().to(d), T(g["K"]).to(d), T(g["G_in"]).contiguous().to(d),
                             T(gb[f"{case}_grad_delta"]).contiguous().to(d), ep_lmbda=ep)
    rt, rw = T(gb[f"{case}_grad_target"]), T(gb[f"{case}_grad_weight"])
    torch.testing.assert_close(gt.cpu(), rt, rtol=1e-4, atol=2e-6 * rt.abs().max().item())
    torch.testing.assert_close(gw.cpu(), rw, rtol=1e-4, atol=2e-6 * rw.abs().max().item())
    # the forward the gradient belongs to: delta of the same step
    G = T(g["G_in"]).contiguous().to(d).clone()
    _, _, _, delta = ops.lm_solve(T(g["depth"]).to(d), target.to(d), T(g["weight"]).contiguous().to(d), T(g["K"]).to(d), G, 1, ep_lmbda=ep, taps=True)
    torch.testing.assert_close(delta[0].cpu(), T(gb[f"{case}_delta"]), rtol=1e-5, atol=1e-6)


def test_lm_step_autograd_function(ops):
    """ops.lm_step_autograd / ops.LMStep (forward = one LM step, backward = b200pose_lm_backward) under torch autograd against the oracle's
    hand-written gradient for a loss on delta."""
    g = golden("lm.npz")
    d = dev()
    depth, K, G = T(g["depth"]).to(d), T(g["K"]).to(d), T(g["G_in"]).contiguous().to(d)
    target = T(g["target"]).contiguous().to(d).requires_grad_(True)
    weight = T(g["weight"]).contiguous().to(d).requires_grad_(True)
    delta, Gn = ops.lm_step_autograd(depth, target, weight, K, G)
    coef = torch.tensor([[1.0, -2.0, 0.5, 3.0, -1.0, 2.0], [0.3, 0.1, -0.7, 1.5, 2.5, -0.2]], device=d)
    (delta * coef).sum().backward()
    rt, rw = O.lm_step_backward(T(g["depth"]), T(g["target"]), T(g["weight"]), T(g["K"]), T(g["G_in"]), coef.cpu())
    torch.testing.assert_close(target.grad.cpu(), rt, rtol=1e-4, atol=2e-6 * rt.abs().max().item())
    torch.testing.assert_close(weight.grad.cpu(), rw, rtol=1e-4, atol=2e-6 * rw.abs().max().item())
    Gr, dr, _, _ = O.lm_step(T(g["depth"]), T(g["target"]), T(g["weight"]), T(g["K"]), T(g["G_in"]))
    torch.testing.assert_close(Gn.cpu(), Gr, rtol=0, atol=2e-6)


# ----------------------------------------------------------------------------- GPU pins of the small branches
def test_se3_retract_golden_both_branches(ops):
    """geometry/se3.py _se3_matrix_expm executed (expm.npz): Taylor branch (theta < 1e-4), Rodrigues branch, the threshold."""
    g = golden("expm.npz")
    xi = T(g["xi"]).float()
    G = torch.eye(4)[None].repeat(xi.shape[0], 1, 1).contiguous().to(dev())
    out = ops.se3_retract(xi.to(dev()), G).cpu()
    torch.testing.assert_close(out, T(g["G"]), rtol=1e-6, atol=1e-7)
    # composition: exp(d) G0 against the oracle
    G0 = O.se3_exp(torch.tensor([[0.2, -0.1, 0.3, 0.5, 0.1, -0.4]])).repeat(xi.shape[0], 1, 1)
    out2 = ops.se3_retract(xi.to(dev()), G0.clone().to(dev())).cpu()
    torch.testing.assert_close(out2, torch.matmul(O.se3_exp(xi), G0), rtol=1e-6, atol=2e-7)


def test_cholesky_solve_golden(ops):
    """geometry/cholesky.py solve executed (cholesky.npz): the module's own 3x3 test system (embedded in a 6x6 block
    system; its answer hits the +-1 clamp), eight random 6x6 systems, and NaN -> 0."""
    g = golden("cholesky.npz")
    H3 = torch.eye(6, dtype=torch.float64); H3[:3, :3] = T(g["H3"])
    b3 = torch.zeros(6, dtype=torch.float64); b3[:3] = T(g["b3"])
    x3 = ops.cholesky_solve(H3[None].contiguous().to(dev()), b3[None].contiguous().to(dev())).cpu()[0]
    torch.testing.assert_close(x3[:3], T(g["x3"]).float(), rtol=1e-6, atol=1e-7)
    assert x3[0] == -1.0 and x3[1] == 1.0 and torch.all(x3[3:] == 0)
    x6 = ops.cholesky_solve(T(g["H6"]).contiguous().to(dev()), T(g["b6"]).contiguous().to(dev())).cpu()
    torch.testing.assert_close(x6, T(g["x6"]).float(), rtol=1e-6, atol=1e-7)
    Hn = T(g["H6"]).clone(); Hn[0, 2, 2] = float("nan")
    xn = ops.cholesky_solve(Hn.contiguous().to(dev()), T(g["b6"]).contiguous().to(dev())).cpu()
    assert torch.all(xn[0] == 0) and torch.equal(xn[1:], x6[1:])


def test_cfnet_sequence_state_carry_golden(ops, packed):
    """Three consecutive GRU_CFUpdator.forward calls of the executed reference (cfnet_seq.npz; update_corr_fn only on the
    first): the hidden state and the context features must carry over between the per-operator calls on the GPU."""
    g = golden("cfnet_seq.npz")
    B, H, W, s1, s2, s3, s4 = [int(v) for v in g["meta"]]
    h, w = H // 8, W // 8
    f1 = S.hash_features((B, 256, h, w), s1).to(dev()); f2 = S.hash_features((B, 256, h, w), s2).to(dev())
    ctx = S.hash_features((B, 256, H, W), s3, 0.1).to(dev())
    u, v = O.pixel_grid(h, w)
    grid = to_pxc(torch.stack([u, v])[None].expand(B, 2, h, w)).to(dev())
    for flags in (0, 1):
        pyr = ops.corr_pyramid(f1, f2)                               # update_corr_fn == True part (CFNet.py:115-133)
        net, xbuf = ops.context_init(ctx)
        for it in range(3):
            fi = S.hash_features((B, 2, H, W), s4 + it, 6.0)
            fl_lr = O.downsample_align_corners(fi / 8.0, 8)            # CFNet.py:138-142 (input preparation)
            flow = to_pxc(fl_lr).to(dev()).contiguous()
            coords1 = (grid + flow).contiguous()
            corr = ops.corr_lookup(pyr, coords1, B, h, w)
            mask, _ = ops.update_block(packed, net, xbuf, corr, coords1, flow, B, h, w, flags=flags)
            fu, _, _ = ops.upsample_weight(flow, mask, None, None, None, 1.0, B, H, W)
            tol = 1.0 if flags == 0 else 4.0
            torch.testing.assert_close(fu.cpu(), T(g["flow_up"][it]), rtol=1e-4, atol=2e-4 * tol)
            torch.testing.assert_close(from_pxc(net, B, h, w).cpu(), T(g["net"][it]), rtol=1e-4, atol=2e-5 * tol * 2)


# ----------------------------------------------------------------------------- f3: pose metrics kernel
def _metric_cases():
    """tests/golden/metrics.npz: the reference's LineMODEvaluator executed (make_golden_metrics.py)."""
    g = golden("metrics.npz")
    rows = g["rows"]
    by_set = {}
    for r in rows:
        by_set.setdefault(int(r[0]), []).append(r)
    return g, by_set


def test_pose_metrics_reference_golden(ops):
    """ADD, ADD-S (ground-truth query -> nearest predicted point, eval_metric.py:167-171), ADD2/ADD5, 2-D projection and
    5cm5deg against the executed reference: distances within 1e-6 d, every flag identical (the golden holds pairs 0.5 %
    either side of each threshold)."""
    from rnnpose_b200 import metrics as M
    g, by_set = _metric_cases()
    K = T(g["K"]).float()
    for si, rows in by_set.items():
        pts = T(g[f"pts{si}"]).float()
        n = len(rows)
        Tp = torch.eye(4).repeat(n, 1, 1); Tg = torch.eye(4).repeat(n, 1, 1)
        for k, r in enumerate(rows):
            Tp[k, :3] = T(g[f"pose_pred_{si}_{int(r[1])}"]); Tg[k, :3] = T(g[f"pose_gt_{si}_{int(r[1])}"])
        R = torch.from_numpy(np.asarray(rows))
        d = R[:, 2]
        out = M.pose_metrics(Tp.to(dev()), Tg.to(dev()), pts[None].repeat(n, 1, 1).to(dev()), d.float().to(dev()),
                             torch.arange(n).to(dev()), K.to(dev())).cpu().double()
        assert out.shape == (n, 16)
        assert ((out[:, 0] - R[:, 3]).abs() / d).max() < 1e-6, "ADD"
        assert ((out[:, 1] - R[:, 4]).abs() / d).max() < 1e-6, "ADD-S"
        torch.testing.assert_close(out[:, 4], R[:, 5], rtol=2e-5, atol=2e-4)          # projection error, pixels
        fin = torch.isfinite(R[:, 6]) & (R[:, 6] > 0.5) & (R[:, 6] < 179.5)
        torch.testing.assert_close(out[fin, 5], R[fin, 6], rtol=1e-4, atol=3e-2)      # acos near 0 / 180 deg is ill-conditioned
        ok = R[:, 7] < 170                                                             # asin at 1 is ill-conditioned
        torch.testing.assert_close(out[ok, 2], R[ok, 7], rtol=1e-4, atol=2e-3)
        torch.testing.assert_close(out[~ok, 2], R[~ok, 7], rtol=0, atol=0.1)
        torch.testing.assert_close(out[:, 3], R[:, 8], rtol=1e-5, atol=1e-7)
        for col, gcol in ((6, 9), (7, 10), (8, 11), (9, 12), (10, 13), (11, 14), (12, 15), (13, 16)):
            assert torch.equal(out[:, col], R[:, gcol]), f"flag column {col} (set {si})"
        assert out[:, 14].tolist() == [float(i) for i in range(n)]
        torch.testing.assert_close(out[:, 15], d, rtol=1e-6, atol=0)


@pytest.mark.parametrize("n_pts", [1, 50, 256, 300, 2500])
def test_pose_metrics_vs_oracle(ops, n_pts):
    from rnnpose_b200 import metrics as M
    g = torch.Generator().manual_seed(n_pts)
    B = 5
    xi = torch.randn(B, 6, generator=g) * 0.1
    Tp = O.se3_exp(xi); Tg = O.se3_exp(xi + 0.02 * torch.randn(B, 6, generator=g))
    Tg[0] = Tp[0]                                                      # an exact pose: every metric 0, every flag 1
    Tp[:, 2, 3] += 0.8; Tg[:, 2, 3] += 0.8                             # in front of the camera (projection)
    pts = torch.randn(n_pts, 3, generator=g) * 0.05
    diam = torch.tensor([0.15, 0.15, 0.01, 0.15, 0.3])
    out = ops.pose_metrics(Tp.cuda(), Tg.cuda(), pts[None].repeat(B, 1, 1).cuda(), diam.cuda()).cpu()
    Rp, tp, Rg, tg = Tp[:, :3, :3], Tp[:, :3, 3], Tg[:, :3, :3], Tg[:, :3, 3]
    torch.testing.assert_close(out[:, 0], O.add_metric(Rp, tp, Rg, tg, pts, False), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(out[:, 1], O.add_metric(Rp, tp, Rg, tg, pts, True), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(out[:, 2], O.rotation_angle_deg(Rp, Rg), rtol=1e-4, atol=2e-3)
    torch.testing.assert_close(out[:, 3], (tp - tg).norm(dim=1), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(out[:, 4], O.projection_2d(Rp, tp, Rg, tg, pts, torch.tensor(ops.LINEMOD_K)), rtol=1e-4, atol=1e-3)
    assert out[0, :5].abs().max() < 1e-4 and out[0, 6:14].tolist() == [1.0] * 8
    ref = M.pose_metrics_torch(Tp, Tg, pts[None].repeat(B, 1, 1), diam, torch.zeros(B))
    for col, val, frac in ((6, 0, 0.1), (7, 1, 0.1), (8, 0, 0.02), (9, 1, 0.02), (10, 0, 0.05), (11, 1, 0.05)):
        margin = (ref[:, val] - frac * diam).abs() > 1e-6              # flags can only differ on a threshold tie
        assert torch.equal(out[margin, col], ref[margin, col])
    assert torch.equal(out[:, 13], ref[:, 13])
    full = M.pose_metrics(Tp.cuda(), Tg.cuda(), pts[None].repeat(B, 1, 1).cuda(), diam.cuda(), torch.arange(B).cuda()).cpu()
    assert full[:, 14].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0] and torch.equal(full[:, :14], out[:, :14])


def test_pose_metrics_error_codes(ops):
    from rnnpose_b200 import _lib
    L = _lib.lib()
    t = torch.zeros(64, device="cuda")
    p = t.data_ptr()
    assert L.b200pose_pose_metrics(0, p, p, p, p, 1, 4, p, p, 1024, 0) == -1
    assert L.b200pose_pose_metrics(p, p, p, p, 0, 1, 4, p, p, 1024, 0) == -1      # K is required
    assert L.b200pose_pose_metrics(p, p, p, p, p, 0, 4, p, p, 1024, 0) == -2
    assert L.b200pose_pose_metrics(p, p, p, p, p, 1, 4, p, p, 0, 0) == -3

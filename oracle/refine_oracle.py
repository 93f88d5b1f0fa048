"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (torch-CPU tensors, fp32 with an fp64 LM step, exactly like the reference's CPU
path) of the RNNPose recurrent pose-refinement inner loop, ``model/PoseRefiner.py:315-365`` of the
reference and everything it calls.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product path
(``rnnpose_b200``) never does and has no CPU fallback.

Parity pin: ``tests/golden/*.npz`` hold outputs of the UNMODIFIED reference executed in the build
container (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks every function
here against them.  The reference has no tests of its own on this path.

Unlike the reference (B=1 only, SURVEY finding 1) every function is batched over B; samples are
independent, so a batch equals B separate reference calls.

Every function cites the reference file:line it restates (paths relative to /root/reference).
The arithmetic library underneath (conv2d / matmul) is PyTorch, as in the reference
(``docker/freeze.yml:133`` pins pytorch 1.7.1); resampling, pooling, the Jacobians, the 6x6 solve
and the exponential map are written out explicitly from the formulas rather than delegated to
``grid_sample`` / ``interpolate`` / ``torch.cholesky`` so that the restatement is independent.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

CORR_LEVELS = 4      # model/CFNet.py:59
CORR_RADIUS = 4      # model/CFNet.py:60
EPS_DEPTH = 1e-5     # model/PoseRefiner.py:21  (EPS)
MIN_DEPTH_VALID = 0.1   # geometry/transformation.py:16
MIN_DEPTH_PROJ = 0.01   # geometry/projective_ops.py:9
MIN_THETA = 1e-4        # geometry/se3.py:10
LM_LMBDA = 1e-4         # config/default.py:54
EP_LMBDA = 100.0        # config/default.py:55


# ----------------------------------------------------------------------------- helpers
def _bilinear_zero(img: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Bilinear sample of img [N,Hs,Ws] at pixel coordinates x,y [N,...] with zero padding.
    Pixel-exact coordinates (what grid_sample(align_corners=True) computes after
    ``bilinear_sampler`` un-does its own normalisation, thirdparty/raft/utils/utils.py:57-65)."""
    N, Hs, Ws = img.shape
    x0 = torch.floor(x); y0 = torch.floor(y)
    fx = x - x0; fy = y - y0
    x0 = x0.long(); y0 = y0.long()
    flat = img.reshape(N, -1)
    shp = x.shape

    def tap(xi, yi):
        ok = (xi >= 0) & (xi < Ws) & (yi >= 0) & (yi < Hs)
        idx = (yi.clamp(0, Hs - 1) * Ws + xi.clamp(0, Ws - 1)).reshape(N, -1)
        v = torch.gather(flat, 1, idx).reshape(shp)
        return v * ok.to(v.dtype)

    v00 = tap(x0, y0); v01 = tap(x0 + 1, y0); v10 = tap(x0, y0 + 1); v11 = tap(x0 + 1, y0 + 1)
    return (v00 * (1 - fx) * (1 - fy) + v01 * fx * (1 - fy) + v10 * (1 - fx) * fy + v11 * fx * fy)


def pixel_grid(H: int, W: int, device=None, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """u (column) and v (row) coordinate planes [H,W] (geometry/projective_ops.py:25-44)."""
    v, u = torch.meshgrid(torch.arange(H, device=device, dtype=dtype),
                          torch.arange(W, device=device, dtype=dtype), indexing="ij")
    return u, v


# ----------------------------------------------------------------------------- a1: correlation pyramid
def corr_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, levels: int = CORR_LEVELS) -> List[torch.Tensor]:
    """All-pairs correlation volume and its average-pooled pyramid.
    thirdparty/raft/corr.py:60-67 (volume, / sqrt(D)) and :28-34 (2x2 avg-pool with floor).
    Returns levels tensors [B, P, h/2^l, w/2^l] (P = h*w source pixels)."""
    B, D, h, w = fmap1.shape
    f1 = fmap1.reshape(B, D, h * w).float()
    f2 = fmap2.reshape(B, D, h * w).float()
    c = torch.matmul(f1.transpose(1, 2), f2) / math.sqrt(D)
    c = c.reshape(B, h * w, h, w)
    pyr = [c]
    for _ in range(levels - 1):
        hh, ww = c.shape[-2] // 2, c.shape[-1] // 2
        c = c[..., :2 * hh, :2 * ww]
        c = (c[..., 0::2, 0::2] + c[..., 0::2, 1::2] + c[..., 1::2, 0::2] + c[..., 1::2, 1::2]) * 0.25
        pyr.append(c)
    return pyr


# ----------------------------------------------------------------------------- a2: lookup
def corr_lookup(pyr: List[torch.Tensor], coords: torch.Tensor, radius: int = CORR_RADIUS) -> torch.Tensor:
    """9x9 bilinear window per level around coords/2^l.  thirdparty/raft/corr.py:36-57.
    coords [B,2,h,w] as (x,y).  Output [B, levels*81, h, w], channel = l*81 + i*9 + j where the
    SLOW index i offsets x and the fast index j offsets y (the reference's meshgrid(dy,dx) stacked
    into the (x,y) slots, corr.py:44-50)."""
    B, _, h, w = coords.shape
    P = h * w
    r = radius
    n = 2 * r + 1
    d = torch.arange(-r, r + 1, dtype=coords.dtype, device=coords.device)
    di = d.view(n, 1).expand(n, n)     # varies with slow index i -> added to x
    dj = d.view(1, n).expand(n, n)     # varies with fast index j -> added to y
    cx = coords[:, 0].reshape(B * P, 1, 1)
    cy = coords[:, 1].reshape(B * P, 1, 1)
    outs = []
    for l, c in enumerate(pyr):
        hl, wl = c.shape[-2:]
        img = c.reshape(B * P, hl, wl)
        x = cx / (2 ** l) + di
        y = cy / (2 ** l) + dj
        s = _bilinear_zero(img, x, y)              # [B*P, 9, 9]
        outs.append(s.reshape(B, h, w, n * n))
    out = torch.cat(outs, dim=-1)
    return out.permute(0, 3, 1, 2).contiguous().float()


# ----------------------------------------------------------------------------- a3-a5: update block
def update_block(wts: Dict[str, torch.Tensor], net: torch.Tensor, inp: torch.Tensor,
                 corr: torch.Tensor, flow: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """BasicUpdateBlock.forward, thirdparty/raft/update.py:179-188 with BasicMotionEncoder
    (:89-97), SepConvGRU (:45-60), FlowHead (:13-14) and the mask head (:172-176,187).
    ``wts`` uses the state-dict keys of ``cf_net.update_block``."""
    def conv(x, name, pad):
        return F.conv2d(x, wts[name + ".weight"], wts[name + ".bias"], padding=pad)

    cor = F.relu(conv(corr, "encoder.convc1", 0))
    cor = F.relu(conv(cor, "encoder.convc2", 1))
    flo = F.relu(conv(flow, "encoder.convf1", 3))
    flo = F.relu(conv(flo, "encoder.convf2", 1))
    out = F.relu(conv(torch.cat([cor, flo], 1), "encoder.conv", 1))
    motion = torch.cat([out, flow], 1)
    x = torch.cat([inp, motion], 1)
    h = net
    for sfx, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(conv(hx, "gru.convz" + sfx, pad))
        r = torch.sigmoid(conv(hx, "gru.convr" + sfx, pad))
        q = torch.tanh(conv(torch.cat([r * h, x], 1), "gru.convq" + sfx, pad))
        h = (1 - z) * h + z * q
    dflow = conv(F.relu(conv(h, "flow_head.conv1", 1)), "flow_head.conv2", 1)
    mask = 0.25 * conv(F.relu(conv(h, "mask.0", 1)), "mask.2", 0)
    return h, mask, dflow


# ----------------------------------------------------------------------------- a7: convex upsampling
def convex_upsample(flow: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """GRU_CFUpdator.upsample_flow, model/CFNet.py:95-106.  flow [B,2,h,w], mask [B,576,h,w]
    -> [B,2,8h,8w]:  out[c,8y+i,8x+j] = sum_k softmax_k(mask[k*64+i*8+j,y,x]) * 8*flow[c,y+k//3-1,x+k%3-1]."""
    B, _, h, w = flow.shape
    m = torch.softmax(mask.reshape(B, 9, 8, 8, h, w), dim=1)
    fp = F.pad(8.0 * flow, (1, 1, 1, 1))
    out = torch.zeros(B, 2, 8, 8, h, w, dtype=flow.dtype, device=flow.device)
    for k in range(9):
        ky, kx = k // 3, k % 3
        nb = fp[:, :, ky:ky + h, kx:kx + w]                 # flow[y+ky-1, x+kx-1]
        out = out + m[:, k].unsqueeze(1) * nb.reshape(B, 2, 1, 1, h, w)
    return out.permute(0, 1, 4, 2, 5, 3).reshape(B, 2, 8 * h, 8 * w)


# ----------------------------------------------------------------------------- a6 helpers
def downsample_align_corners(x: torch.Tensor, ds: int = 8) -> torch.Tensor:
    """F.interpolate(x, scale_factor=1/ds, mode='bilinear', align_corners=True) restated:
    src = dst * (In-1)/(Out-1)  (model/CFNet.py:129,142)."""
    B, C, H, W = x.shape
    h, w = H // ds, W // ds
    sy = (H - 1) / (h - 1) if h > 1 else 0.0
    sx = (W - 1) / (w - 1) if w > 1 else 0.0
    yy = torch.arange(h, dtype=torch.float32, device=x.device) * torch.tensor(sy, dtype=torch.float32)
    xx = torch.arange(w, dtype=torch.float32, device=x.device) * torch.tensor(sx, dtype=torch.float32)
    y0 = yy.floor().long().clamp(max=H - 1); x0 = xx.floor().long().clamp(max=W - 1)
    y1 = (y0 + 1).clamp(max=H - 1); x1 = (x0 + 1).clamp(max=W - 1)
    ly = (yy - y0.float()).view(1, 1, h, 1); lx = (xx - x0.float()).view(1, 1, 1, w)
    top = x[:, :, y0][:, :, :, x0] * (1 - lx) + x[:, :, y0][:, :, :, x1] * lx
    bot = x[:, :, y1][:, :, :, x0] * (1 - lx) + x[:, :, y1][:, :, :, x1] * lx
    return top * (1 - ly) + bot * ly


def context_init(context_fea: torch.Tensor, w_low: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Hidden state / input features from the (already x0.1) context map, model/CFNet.py:124-133
    (ds = context width // fmap width)."""
    cnet = downsample_align_corners(context_fea, context_fea.shape[-1] // w_low)
    net, inp = torch.split(cnet, [128, 128], dim=1)
    return torch.tanh(net), torch.relu(inp)


# ----------------------------------------------------------------------------- a8: geometry
def _intr(K: torch.Tensor):
    return (K[:, 0, 0].view(-1, 1, 1), K[:, 1, 1].view(-1, 1, 1), K[:, 0, 2].view(-1, 1, 1), K[:, 1, 2].view(-1, 1, 1))


def backproject(depth: torch.Tensor, K: torch.Tensor) -> torch.Tensor:
    """geometry/projective_ops.py:68-99.  depth [B,H,W] -> X0 [B,H,W,3]."""
    B, H, W = depth.shape
    u, v = pixel_grid(H, W, depth.device)
    fx, fy, cx, cy = _intr(K)
    X = depth * (u - cx) / fx
    Y = depth * (v - cy) / fy
    return torch.stack([X, Y, depth], dim=-1)


def se3_apply(G: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """X1 = G[:3,:] [X;1]  (geometry/transformation.py:78-85, eq 'aijk,ai...k->ai...j').  G [B,4,4]."""
    Xh = torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)
    return torch.einsum("bjk,bhwk->bhwj", G[:, :3, :].to(X.dtype), Xh)


def project(X: torch.Tensor, K: torch.Tensor) -> torch.Tensor:
    """geometry/projective_ops.py:103-114 (Z clamped at MIN_DEPTH=0.01)."""
    fx, fy, cx, cy = _intr(K)
    Z = torch.clamp(X[..., 2], min=MIN_DEPTH_PROJ)
    return torch.stack([fx * (X[..., 0] / Z) + cx, fy * (X[..., 1] / Z) + cy], dim=-1)


def transform(depth: torch.Tensor, K: torch.Tensor, G: torch.Tensor):
    """SE3.transform, geometry/transformation.py:184-198. depth [B,H,W] (already +EPS), returns
    reprojected coords [B,H,W,2] and the validity mask [B,H,W]."""
    X0 = backproject(depth, K)
    X1 = se3_apply(G, X0)
    vmask = ((X0[..., 2] > MIN_DEPTH_VALID) & (X1[..., 2] > MIN_DEPTH_VALID)).float()
    return project(X1, K), vmask


def flow_init_lowres(depth_eps: torch.Tensor, K: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    """flow_init of model/PoseRefiner.py:324-328, divided by 8 and 1/8-resampled with
    align_corners=True as in model/CFNet.py:138-142.  Returns [B,2,h,w]."""
    B, H, W = depth_eps.shape
    x1, _ = transform(depth_eps, K, G)
    u, v = pixel_grid(H, W, depth_eps.device)
    fl = torch.stack([x1[..., 0] - u, x1[..., 1] - v], dim=1) * (depth_eps > EPS_DEPTH).float().unsqueeze(1)
    return downsample_align_corners(fl / 8.0, 8)


# ----------------------------------------------------------------------------- a9: correspondence weight
def corr_weight(geofea1: torch.Tensor, geofea2: torch.Tensor, target: torch.Tensor,
                syn_depth: torch.Tensor, sigma: float) -> torch.Tensor:
    """model/PoseRefiner.py:342-345 with projective_ops.normalize_coords_grid (:11-23).
    target [B,H,W,2] pixel coords; the normalised grid (align_corners=True convention) is fed to
    grid_sample with its default align_corners=False, i.e. sampled position
    px = ((2*tx/(W-1)-1+1)*W-1)/2.  Returns w [B,H,W]."""
    B, C, H, W = geofea2.shape
    gx = 2 * target[..., 0] / (W - 1) - 1
    gy = 2 * target[..., 1] / (H - 1) - 1
    px = ((gx + 1) * W - 1) / 2
    py = ((gy + 1) * H - 1) / 2
    img = geofea2.reshape(B * C, H, W)
    pxe = px.unsqueeze(1).expand(B, C, H, W).reshape(B * C, H, W)
    pye = py.unsqueeze(1).expand(B, C, H, W).reshape(B * C, H, W)
    warp = _bilinear_zero(img, pxe, pye).reshape(B, C, H, W)
    s = (geofea1 * warp).sum(dim=1)
    return torch.exp(-torch.abs(1 - s) / sigma) * (syn_depth > 0).float()


# ----------------------------------------------------------------------------- a12: se3 exp
def se3_exp(xi: torch.Tensor) -> torch.Tensor:
    """_se3_matrix_expm, geometry/se3.py:228-281.  xi [B,6] = (upsilon, omega) -> [B,4,4]."""
    dt = xi.dtype
    v, w = xi[:, :3], xi[:, 3:]
    th2 = (w * w).sum(dim=1).view(-1, 1, 1)
    th = torch.sqrt(th2)
    th4 = th2 * th2
    z = torch.zeros_like(w[:, 0])
    wx = torch.stack([torch.stack([z, -w[:, 2], w[:, 1]], -1),
                      torch.stack([w[:, 2], z, -w[:, 0]], -1),
                      torch.stack([-w[:, 1], w[:, 0], z], -1)], dim=-2)
    wx2 = torch.matmul(wx, wx)
    I = torch.eye(3, dtype=dt, device=xi.device).expand(xi.shape[0], 3, 3)
    R1 = I + (1.0 - th2 / 6.0 + th4 / 120.0) * wx + (0.5 - th2 / 12.0 + th4 / 720.0) * wx2
    V1 = I + (0.5 - th2 / 24.0 + th4 / 720.0) * wx + (1.0 / 6.0 - th2 / 120.0 + th4 / 5040.0) * wx2
    eps = 1e-12
    R2 = I + (torch.sin(th) / (th + eps)) * wx + ((1 - torch.cos(th)) / (th2 + eps)) * wx2
    V2 = I + ((1 - torch.cos(th)) / (th2 + eps)) * wx + ((th - torch.sin(th)) / (th2 * th + eps)) * wx2
    small = th < MIN_THETA
    R = torch.where(small, R1, R2)
    V = torch.where(small, V1, V2)
    t = torch.matmul(V, v.unsqueeze(-1))
    G = torch.zeros(xi.shape[0], 4, 4, dtype=dt, device=xi.device)
    G[:, :3, :3] = R; G[:, :3, 3:] = t; G[:, 3, 3] = 1
    return G


# ----------------------------------------------------------------------------- a11: 6x6 solve
def cholesky_solve6(H: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """geometry/cholesky.py:32-50: fp64 Cholesky solve, NaN -> 0, clamp to [-1,1], cast fp32.
    Written out (no LAPACK) so the restatement is independent.  H [B,6,6], b [B,6] (fp64)."""
    H = H.double(); b = b.double()
    B, n, _ = H.shape
    L = torch.zeros_like(H)
    for j in range(n):
        s = H[:, j, j] - (L[:, j, :j] ** 2).sum(dim=1)
        L[:, j, j] = torch.sqrt(s)
        for i in range(j + 1, n):
            L[:, i, j] = (H[:, i, j] - (L[:, i, :j] * L[:, j, :j]).sum(dim=1)) / L[:, j, j]
    y = torch.zeros_like(b)
    for i in range(n):
        y[:, i] = (b[:, i] - (L[:, i, :i] * y[:, :i]).sum(dim=1)) / L[:, i, i]
    x = torch.zeros_like(b)
    for i in reversed(range(n)):
        x[:, i] = (y[:, i] - (L[:, i + 1:, i] * x[:, i + 1:]).sum(dim=1)) / L[:, i, i]
    x = torch.where(torch.isnan(x), torch.zeros_like(x), x)
    return torch.clamp(x, -1.0, 1.0).float()


# ----------------------------------------------------------------------------- a10: LM normal equations
def lm_jacobian_residual(depth_eps: torch.Tensor, target: torch.Tensor, K: torch.Tensor, G: torch.Tensor):
    """Per-pixel 2x6 Jacobian J (fp64 from fp32 geometry), residual r = target - x1 and validity v.
    geometry/transformation.py:284-292, jac_local_perturb :27-46, project(jacobian=True) projective_ops.py:116-131.
    depth_eps [B,H,W], target [B,H,W,2] -> J [B,H,W,2,6], r [B,H,W,2], v [B,H,W]."""
    X0 = backproject(depth_eps, K)                       # fp32, as the reference
    X1 = se3_apply(G, X0)
    fx, fy, cx, cy = _intr(K)
    X, Y, Zr = X1[..., 0], X1[..., 1], X1[..., 2]
    Z = torch.clamp(Zr, min=MIN_DEPTH_PROJ)
    x1 = torch.stack([fx * (X / Z) + cx, fy * (Y / Z) + cy], dim=-1)
    cut = Z <= MIN_DEPTH_PROJ + 0.01
    zinv1 = torch.where(cut, torch.zeros_like(Z), 1.0 / Z)
    zinv2 = torch.where(cut, torch.zeros_like(Z), 1.0 / Z ** 2)
    o = torch.zeros_like(Z)
    jproj = torch.stack([torch.stack([fx * zinv1, o, -fx * X * zinv2], -1),
                         torch.stack([o, fy * zinv1, -fy * Y * zinv2], -1)], dim=-2)      # [B,H,W,2,3]
    one = torch.ones_like(Z)
    # columns of [I | -[X1]x] with UNclamped X1 (transformation.py:29-45)
    jtran = torch.stack([torch.stack([one, o, o], -1), torch.stack([o, one, o], -1), torch.stack([o, o, one], -1),
                         torch.stack([o, -Zr, Y], -1), torch.stack([Zr, o, -X], -1), torch.stack([-Y, X, o], -1)],
                        dim=-1)                                                               # [B,H,W,3,6]
    J = torch.matmul(jproj.double(), jtran.double())                                          # [B,H,W,2,6]
    v = ((X0[..., 2] > MIN_DEPTH_VALID) & (Zr > MIN_DEPTH_VALID)).double()
    r = target.double() - x1.double()                                                         # fp64 - fp32->fp64
    return J, r, v


def lm_normal_equations(depth_eps: torch.Tensor, target: torch.Tensor, weight: torch.Tensor,
                        K: torch.Tensor, G: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """H = sum v w J^T J, b = sum v w J^T r in fp64 from fp32 geometry (transformation.py:294-297).
    depth_eps/weight [B,H,W], target [B,H,W,2].  Un-damped."""
    J, r, v = lm_jacobian_residual(depth_eps, target, K, G)
    vw = (v * weight.double())[..., None, None]
    Hm = torch.einsum("bhwri,bhwrj->bij", vw * J, J)
    bv = torch.einsum("bhwri,bhwr->bi", vw * J, r)
    return Hm, bv


def lm_step(depth_eps, target, weight, K, G, lm_lmbda=LM_LMBDA, ep_lmbda=EP_LMBDA):
    """One damped Gauss-Newton step and retraction: transformation.py:284-306, se3.py:303-306."""
    Hm, bv = lm_normal_equations(depth_eps, target, weight, K, G)
    eye = torch.eye(6, dtype=Hm.dtype, device=Hm.device)
    Hd = Hm + ep_lmbda * eye + lm_lmbda * Hm * eye
    delta = cholesky_solve6(Hd, bv)
    Gn = torch.matmul(se3_exp(delta), G)
    return Gn, delta, Hm, bv


def lm_step_backward(depth_eps, target, weight, K, G, grad_delta, lm_lmbda=LM_LMBDA, ep_lmbda=EP_LMBDA):
    """Gradients of one LM step with respect to target [B,H,W,2] and weight [B,H,W] given dL/d(delta) [B,6], written out by
    hand (no autograd): geometry/cholesky.py:19-28 (z = H_d^-1 dx, dH_d = -x z^T, db = z, x the raw solution; dx is zero where
    the NaN -> 0 / clamp of :42-45 blocks it), transformation.py:300 (dH = dH_d + lm diag(dH_d)), :294-297 (H, b linear in w;
    b linear in target).  Pinned by tests/golden/lm_backward.npz (the reference's autograd executed)."""
    B, H, W = depth_eps.shape
    J, r, v = lm_jacobian_residual(depth_eps, target, K, G)                       # [B,H,W,2,6], [B,H,W,2], [B,H,W] in fp64
    wv = (v * weight.double())[..., None, None]
    Hm = torch.einsum("bhwri,bhwrj->bij", wv * J, J)
    bv = torch.einsum("bhwri,bhwr->bi", wv * J, r)
    eye = torch.eye(6, dtype=torch.float64)
    Hd = Hm + ep_lmbda * eye + lm_lmbda * Hm * eye
    Lc = torch.linalg.cholesky(Hd)
    x = torch.cholesky_solve(bv[..., None], Lc)[..., 0]
    ok = (~torch.isnan(x)) & (x >= -1.0) & (x <= 1.0)
    dx = torch.where(ok, grad_delta.double(), torch.zeros_like(x))
    z = torch.cholesky_solve(dx[..., None], Lc)[..., 0]
    dHd = -torch.einsum("bi,bj->bij", x, z)
    dH = dHd + lm_lmbda * dHd * eye
    jb = torch.einsum("bhwri,bi->bhwr", J, z)                                     # J_row . dL/db
    q = torch.einsum("bhwri,bij,bhwrj->bhwr", J, dH, J)
    grad_w = v * (q + jb * r).sum(-1)
    grad_t = (v * weight.double())[..., None] * jb
    return grad_t.float(), grad_w.float()


# ----------------------------------------------------------------------------- a13
def se3_inverse(G: torch.Tensor) -> torch.Tensor:
    """geometry/se3.py:194-209."""
    R = G[..., :3, :3].transpose(-1, -2)
    t = -torch.matmul(R, G[..., :3, 3:])
    out = torch.zeros_like(G)
    out[..., :3, :3] = R; out[..., :3, 3:] = t; out[..., 3, 3] = 1
    return out


# ----------------------------------------------------------------------------- a6 + a14: the loop
class RefineState:
    """What GRU_CFUpdator keeps between calls (self.corr_fn, self.net, self.inp; CFNet.py:115-133)."""
    def __init__(self):
        self.pyr = None; self.net = None; self.inp = None


def cf_net_forward(wts, st: RefineState, fmap1, fmap2, flow_init_lr, context_fea, update_corr_fn: bool):
    """GRU_CFUpdator.forward for iters=1, model/CFNet.py:109-173, taking the ALREADY 1/8-resampled
    flow_init/8 (see flow_init_lowres).  Returns flow_up [B,2,H,W] plus the low-res pieces."""
    if update_corr_fn:
        st.pyr = corr_pyramid(fmap1, fmap2)
        st.net, st.inp = context_init(context_fea, fmap1.shape[-1])
    B, _, h, w = fmap1.shape
    u, v = pixel_grid(h, w, fmap1.device)
    coords0 = torch.stack([u, v], 0).unsqueeze(0).expand(B, 2, h, w)
    coords1 = coords0 + flow_init_lr
    corr = corr_lookup(st.pyr, coords1)
    flow = coords1 - coords0
    st.net, up_mask, dflow = update_block(wts, st.net, st.inp, corr, flow)
    coords1 = coords1 + dflow
    flow_up = convex_upsample(coords1 - coords0, up_mask)
    return flow_up, dict(corr=corr, flow_lr=coords1 - coords0, mask=up_mask, dflow=dflow, net=st.net)


def refine_inner_loop(wts: Dict[str, torch.Tensor], fmap1, fmap2, context_fea, geofea1, geofea2,
                      syn_depth, K, G0, sigma: float = 1.0, n_iters: int = 4, n_lm: int = 3,
                      lm_lmbda=LM_LMBDA, ep_lmbda=EP_LMBDA, trace: Optional[list] = None):
    """The hot path: model/PoseRefiner.py:313-365 for one render iteration.
    fmap1/2 [B,256,h,w]; context_fea [B,256,H,W] (x0.1 applied); geofea1/2 [B,32,H,W];
    syn_depth [B,1,H,W]; K [B,3,3]; G0 [B,4,4] = Tij at loop entry.
    Returns dict(G [B,4,4], flows list of [B,2,H,W], weight [B,H,W])."""
    st = RefineState()
    depth = syn_depth[:, 0]
    depth_eps = depth + EPS_DEPTH
    B, H, W = depth.shape
    u, v = pixel_grid(H, W, depth.device)
    G = G0.clone()
    flows = []
    weight = None
    for i in range(n_iters):
        fl_lr = flow_init_lowres(depth_eps, K, G)
        flow_up, low = cf_net_forward(wts, st, fmap1, fmap2, fl_lr, context_fea, update_corr_fn=(i == 0))
        flows.append(flow_up)
        target = torch.stack([flow_up[:, 0] + u, flow_up[:, 1] + v], dim=-1)
        weight = corr_weight(geofea1, geofea2, target, depth, sigma)
        deltas = []
        for _ in range(n_lm):
            G, delta, Hm, bv = lm_step(depth_eps, target, weight, K, G, lm_lmbda, ep_lmbda)
            deltas.append(delta)
        if trace is not None:
            trace.append(dict(flow_up=flow_up, target=target, weight=weight, G=G.clone(), deltas=deltas, **low))
    return dict(G=G, flows=flows, weight=weight)


# ----------------------------------------------------------------------------- metrics (A.8)
def add_metric(R_p, t_p, R_g, t_g, pts, symmetric: bool = False) -> torch.Tensor:
    """ADD / ADD-S mean distance, utils/eval_metric.py:161-179.  R [B,3,3], t [B,3], pts [N,3] -> [B].
    ADD-S (syn=True, :167-171): idxs = find_nearest_point_idx(model_pred, model_targets) -- ref = the PREDICTED points,
    que = the GROUND-TRUTH points (thirdparty/nn/nn_utils.py:6-22), each query takes the first nearest reference point
    (thirdparty/nn/src/nearest_neighborhood.cu:48-80) -- then mean ||model_pred[idxs] - model_targets|| over the
    ground-truth points.  Pinned by tests/golden/metrics.npz (the reference's evaluator executed)."""
    pp = torch.einsum("bij,nj->bni", R_p, pts) + t_p[:, None]
    pg = torch.einsum("bij,nj->bni", R_g, pts) + t_g[:, None]
    if symmetric:
        d2 = ((pp[:, :, None, :] - pg[:, None, :, :]) ** 2).sum(-1)            # [B, pred, gt]
        idx = d2.argmin(dim=1)                                                  # nearest predicted point of each gt point
        d = (torch.gather(pp, 1, idx[..., None].expand(-1, -1, 3)) - pg).norm(dim=-1)
    else:
        d = (pp - pg).norm(dim=-1)
    return d.mean(dim=1)


def projection_2d(R_p, t_p, R_g, t_g, pts, K) -> torch.Tensor:
    """Mean pixel distance of the model projected with both poses, utils/eval_metric.py:23-35,102-110.  K [3,3]."""
    def proj(R, t):
        x = torch.einsum("bij,nj->bni", R, pts) + t[:, None]
        x = torch.einsum("ij,bnj->bni", K.to(x), x)
        return x[..., :2] / x[..., 2:]
    return (proj(R_p, t_p) - proj(R_g, t_g)).norm(dim=-1).mean(dim=1)


def cm_degree_5(R_p, t_p, R_g, t_g):
    """utils/eval_metric.py:181-192: (translation distance * 100, angle from the trace in degrees, flag)."""
    trans = (t_p - t_g).norm(dim=1) * 100
    trace = torch.clamp((R_p * R_g).sum(dim=(1, 2)), max=3.0)
    ang = torch.rad2deg(torch.acos((trace - 1.0) / 2.0))
    return trans, ang, (trans < 5) & (ang < 5)


def rotation_angle_deg(R_p, R_g) -> torch.Tensor:
    """utils/geometric.py:36-40: 2*asin(||R_g - R_p||_F / sqrt(8)) in degrees."""
    n = (R_g - R_p).reshape(R_p.shape[0], -1).norm(dim=1)
    return 2 * torch.asin(torch.clamp(n / math.sqrt(8.0), max=1.0)) * 180.0 / math.pi


# ----------------------------------------------------------------------------- timing variant
# The functions above spell every resampling / pooling / solve out explicitly so that the restatement is
# independent of the ATen calls the reference makes.  For the CPU *timing* baseline that would be unfair
# (explicit gathers are ~3x slower than grid_sample), so `refine_inner_loop_aten` below issues the SAME ATen
# op sequence as the reference (grid_sample, interpolate, avg_pool2d, unfold, einsum in fp64,
# linalg.cholesky / cholesky_solve) -- it is what bench.py times as the CPU baseline / reference arm, and
# tests/test_oracle_golden.py checks that it agrees with the explicit restatement and the golden vectors.
def _bilinear_sampler_aten(img, coords):
    """thirdparty/raft/utils/utils.py:57-65."""
    Hs, Ws = img.shape[-2:]
    xg, yg = coords.split([1, 1], dim=-1)
    grid = torch.cat([2 * xg / (Ws - 1) - 1, 2 * yg / (Hs - 1) - 1], dim=-1)
    return F.grid_sample(img, grid, align_corners=True)


def refine_inner_loop_aten(wts, fmap1, fmap2, context_fea, geofea1, geofea2, syn_depth, K, G0, sigma=1.0,
                           n_iters=4, n_lm=3, lm_lmbda=LM_LMBDA, ep_lmbda=EP_LMBDA):
    """Same contract as refine_inner_loop; op sequence of model/PoseRefiner.py:313-365, model/CFNet.py:109-173,
    thirdparty/raft/corr.py:13-57, geometry/transformation.py:265-316 for batch size 1 (as the reference)."""
    assert fmap1.shape[0] == 1, "the reference path is batch-size-1 only (SURVEY finding 1)"
    B, D, h, w = fmap1.shape
    H, W = syn_depth.shape[-2:]
    # CorrBlock.__init__
    corr = torch.matmul(fmap1.view(B, D, h * w).transpose(1, 2), fmap2.view(B, D, h * w)).view(B * h * w, 1, h, w)
    corr = corr / torch.sqrt(torch.tensor(D).float())
    pyr = [corr]
    for _ in range(CORR_LEVELS - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    cnet = F.interpolate(context_fea, scale_factor=1 / 8, mode="bilinear", align_corners=True)
    net, inp = torch.split(cnet, [128, 128], dim=1)
    net = torch.tanh(net); inp = torch.relu(inp)
    depths = syn_depth + EPS_DEPTH                                    # [B,1,H,W]
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    grid_full = torch.stack([xx.float(), yy.float()], dim=-1)[None, None]     # [1,1,H,W,2]
    yl, xl = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    coords0 = torch.stack([xl, yl], dim=0).float()[None]
    fx, fy, cx, cy = (K[:, 0, 0].view(B, 1, 1, 1), K[:, 1, 1].view(B, 1, 1, 1), K[:, 0, 2].view(B, 1, 1, 1), K[:, 1, 2].view(B, 1, 1, 1))
    G = G0.clone()[:, None]                                            # [B,1,4,4]
    eye6 = torch.eye(6, dtype=torch.float64)
    flows = []
    r = CORR_RADIUS
    for it in range(n_iters):
        X0 = torch.stack([depths * (grid_full[..., 0] - cx) / fx, depths * (grid_full[..., 1] - cy) / fy, depths], dim=-1)
        Xh = torch.cat([X0, torch.ones_like(X0[..., :1])], dim=-1)
        X1 = torch.einsum("aijk,ai...k->ai...j", G[..., :3, :], Xh)
        Zc = torch.clamp(X1[..., 2], min=MIN_DEPTH_PROJ)
        reproj = torch.stack([fx * (X1[..., 0] / Zc) + cx, fy * (X1[..., 1] / Zc) + cy], dim=-1)
        flow_init = torch.einsum("...ijk->...kij", reproj - grid_full) * (depths > EPS_DEPTH)
        flow_init = flow_init.squeeze(1) / 8
        flow_init = F.interpolate(flow_init, scale_factor=1 / 8, mode="bilinear", align_corners=True)
        coords1 = coords0 + flow_init
        # CorrBlock.__call__
        cperm = coords1.permute(0, 2, 3, 1)
        outp = []
        for i in range(CORR_LEVELS):
            dx = torch.linspace(-r, r, 2 * r + 1); dy = torch.linspace(-r, r, 2 * r + 1)
            delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), axis=-1)
            cl = cperm.reshape(B * h * w, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
            outp.append(_bilinear_sampler_aten(pyr[i], cl).view(B, h, w, -1))
        corr_f = torch.cat(outp, dim=-1).permute(0, 3, 1, 2).contiguous().float()
        net, up_mask, dflow = update_block(wts, net, inp, corr_f, coords1 - coords0)
        coords1 = coords1 + dflow
        # upsample_flow
        mask = torch.softmax(up_mask.view(B, 1, 9, 8, 8, h, w), dim=2)
        up = F.unfold(8 * (coords1 - coords0), [3, 3], padding=1).view(B, 2, 9, 1, 1, h, w)
        flow_up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(B, 2, H, W)
        flows.append(flow_up)
        target = torch.einsum("...ijk->...jki", flow_up[:, None]) + grid_full
        tn = target.clone()
        tn[..., 0] = 2 * tn[..., 0] / (W - 1) - 1; tn[..., 1] = 2 * tn[..., 1] / (H - 1) - 1
        warp = F.grid_sample(geofea2, tn.squeeze(1))
        cw = torch.sum(geofea1 * warp, dim=1, keepdim=True).permute(0, 2, 3, 1)[:, None]
        cw = torch.exp(-torch.abs(1 - cw) / sigma) * (syn_depth > 0)[..., None].float()
        # reprojction_optim
        tgt64 = target.double(); w64 = cw.double()[..., None]
        for _ in range(n_lm):
            X1 = torch.einsum("aijk,ai...k->ai...j", G[..., :3, :], Xh)
            Xc, Yc, Zr = X1[..., 0], X1[..., 1], X1[..., 2]
            Zc = torch.clamp(Zr, min=MIN_DEPTH_PROJ)
            x1 = torch.stack([fx * (Xc / Zc) + cx, fy * (Yc / Zc) + cy], dim=-1)
            o = torch.zeros_like(Zc); one = torch.ones_like(Zc)
            zi1 = torch.where(Zc <= MIN_DEPTH_PROJ + .01, o, 1.0 / Zc); zi2 = torch.where(Zc <= MIN_DEPTH_PROJ + .01, o, 1.0 / Zc ** 2)
            jproj = torch.stack([torch.stack([fx * zi1, o, -fx * Xc * zi2], -1), torch.stack([o, fy * zi1, -fy * Yc * zi2], -1)], -2)
            jtran = torch.stack([torch.cat([one[..., None], o[..., None], o[..., None]], -1), torch.cat([o[..., None], one[..., None], o[..., None]], -1),
                                 torch.cat([o[..., None], o[..., None], one[..., None]], -1), torch.stack([o, -Zr, Yc], -1),
                                 torch.stack([Zr, o, -Xc], -1), torch.stack([-Yc, Xc, o], -1)], dim=-1)
            v = ((X0[..., -1] > MIN_DEPTH_VALID) & (Zr > MIN_DEPTH_VALID)).double()[..., None, None]
            J = torch.einsum("...ij,...jk->...ik", jproj.double(), jtran.double())
            Hm = torch.einsum("aixyrj,aixyrk->aijk", v * w64 * J, J)
            bv = torch.einsum("aixyrj,aixyr->aij", v * w64 * J, tgt64 - x1)
            Hm = Hm + ep_lmbda * eye6 + lm_lmbda * Hm * eye6
            try:
                x = torch.cholesky_solve(bv.unsqueeze(-1), torch.linalg.cholesky(Hm))
            except Exception:
                x = torch.full_like(bv.unsqueeze(-1), float("nan"))
            x = torch.where(torch.isnan(x), torch.zeros_like(x), x)
            delta = torch.clamp(x, -1.0, 1.0).squeeze(-1).float()
            G = torch.matmul(se3_exp(delta.reshape(-1, 6)).view(B, 1, 4, 4), G)
    return dict(G=G[:, 0], flows=flows, weight=cw[:, 0, :, :, 0])

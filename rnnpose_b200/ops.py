"""Thin torch <-> C-ABI glue: each function hands raw device pointers, sizes and the current CUDA stream
to libb200pose.so.  torch is used for device memory and streams only; all arithmetic happens in the
library's kernels.  Layouts: "PXC" = [B*h*w, C] pixel-major (see include/b200pose.h)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

CORR_PITCH = 328
WEIGHT_KEYS: List[str] = [
    "encoder.convc1", "encoder.convc2", "encoder.convf1", "encoder.convf2", "encoder.conv",
    "gru.convz1", "gru.convr1", "gru.convq1", "gru.convz2", "gru.convr2", "gru.convq2",
    "flow_head.conv1", "flow_head.conv2", "mask.0", "mask.2",
]
FLAG_EXACT_FP32 = 0      # CUDA-core fp32 convolutions
FLAG_TENSOR_CORES = 1    # tcgen05 convolutions on fp16 hi/lo split operands (fp32-level accuracy)
FLAG_GEO2_CHANNELS_LAST = 2   # refine_iters: geofea2 is [B, H*W, 32] (what zoom_crop(channels_last=True) writes)
FLAG_CONTEXT_TEXELS = 4       # refine_iters: context is [B,256,(H/8)*(W/8),4], the texels of the 1/8 resample (context_gather_texels)
DEFAULT_FLAGS = FLAG_TENSOR_CORES
LM_LMBDA = 1e-4   # reference config/default.py:54
EP_LMBDA = 100.0  # reference config/default.py:55


def _need_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("rnnpose_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.device} {t.dtype} "
                         f"contiguous={t.is_contiguous()}")
    return t


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ws(nbytes: int, device) -> torch.Tensor:
    """Scratch buffer whose data_ptr is 1024-byte aligned (TMA-staged planes live inside it)."""
    t = torch.empty(max(int(nbytes), 256) + 1024, dtype=torch.uint8, device=device)
    off = (-t.data_ptr()) % 1024
    return t[off:off + max(int(nbytes), 256)]


def set_option(name: str, value: int) -> None:
    """Kernel-selection option of the library (include/b200pose.h "options")."""
    _lib.check(_lib.lib().b200pose_set_option(name.encode(), int(value)), f"b200pose_set_option({name})")


def get_option(name: str) -> int:
    v = C.c_int()
    _lib.check(_lib.lib().b200pose_get_option(name.encode(), C.byref(v)), f"b200pose_get_option({name})")
    return v.value


def option_names() -> List[str]:
    L = _lib.lib()
    return [L.b200pose_option_name(i).decode() for i in range(L.b200pose_option_count())]


class options:
    """Context manager: ``with ops.options(conv_mode=1): ...`` sets options and restores the previous values."""

    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = get_option(k)
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)
        return False


def pack_weights(state: Dict[str, torch.Tensor], device="cuda") -> torch.Tensor:
    """state: ``cf_net.update_block`` state dict (keys 'encoder.convc1.weight', ...).  Returns the packed
    device blob consumed by update_block / refine_iters."""
    _need_cuda()
    L = _lib.lib()
    tens = []
    for k in WEIGHT_KEYS:
        for sfx in (".weight", ".bias"):
            tens.append(state[k + sfx].detach().to(device=device, dtype=torch.float32).contiguous())
    arr = (C.c_void_p * len(tens))(*[t.data_ptr() for t in tens])
    blob = _ws(L.b200pose_packed_weights_bytes(), device)
    _lib.check(L.b200pose_pack_weights(arr, blob.data_ptr(), _stream()), "b200pose_pack_weights")
    torch.cuda.current_stream().synchronize()      # `tens` must outlive the packing kernels
    return blob


def pyramid_level_views(pyr: torch.Tensor, B: int, h: int, w: int) -> List[torch.Tensor]:
    out, off, hl, wl = [], 0, h, w
    for _ in range(4):
        n = B * h * w * hl * wl
        out.append(pyr[off:off + n].view(B, h * w, hl, wl))
        off += n; hl //= 2; wl //= 2
    return out


def corr_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    _need_cuda()
    L = _lib.lib()
    _chk(fmap1, "fmap1"); _chk(fmap2, "fmap2")
    B, D, h, w = fmap1.shape
    pyr = torch.empty(L.b200pose_pyramid_floats(B, h, w), dtype=torch.float32, device=fmap1.device)
    _lib.check(L.b200pose_corr_pyramid(fmap1.data_ptr(), fmap2.data_ptr(), B, D, h, w, pyr.data_ptr(), _stream()),
               "b200pose_corr_pyramid")
    return pyr


def corr_pyramid_tc(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    """Tensor-core correlation volume + pyramid (D must be 256)."""
    _need_cuda()
    L = _lib.lib()
    _chk(fmap1, "fmap1"); _chk(fmap2, "fmap2")
    B, D, h, w = fmap1.shape
    assert D == 256
    pyr = torch.empty(L.b200pose_pyramid_floats(B, h, w), dtype=torch.float32, device=fmap1.device)
    nb = L.b200pose_corr_pyramid_tc_workspace_bytes(B, h, w)
    ws = _ws(nb, fmap1.device)
    _lib.check(L.b200pose_corr_pyramid_tc(fmap1.data_ptr(), fmap2.data_ptr(), B, h, w, pyr.data_ptr(), ws.data_ptr(), nb,
                                          _stream()), "b200pose_corr_pyramid_tc")
    return pyr


def corr_lookup(pyr: torch.Tensor, coords: torch.Tensor, B: int, h: int, w: int) -> torch.Tensor:
    L = _lib.lib()
    _chk(pyr, "pyramid"); _chk(coords, "coords")
    out = torch.empty(B * h * w, CORR_PITCH, dtype=torch.float32, device=pyr.device)
    _lib.check(L.b200pose_corr_lookup(pyr.data_ptr(), coords.data_ptr(), B, h, w, out.data_ptr(), _stream()),
               "b200pose_corr_lookup")
    return out


def context_init(context: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    L = _lib.lib()
    _chk(context, "context")
    B, Cc, H, W = context.shape
    assert Cc == 256
    P = B * (H // 8) * (W // 8)
    net = torch.empty(P, 128, dtype=torch.float32, device=context.device)
    xbuf = torch.zeros(P, 256, dtype=torch.float32, device=context.device)
    _lib.check(L.b200pose_context_init(context.data_ptr(), B, H, W, net.data_ptr(), xbuf.data_ptr(), _stream()),
               "b200pose_context_init")
    return net, xbuf


def flow_init(depth: torch.Tensor, K: torch.Tensor, G: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    L = _lib.lib()
    _chk(depth, "depth"); _chk(K, "K"); _chk(G, "G")
    B, H, W = depth.shape
    P = B * (H // 8) * (W // 8)
    coords1 = torch.empty(P, 2, dtype=torch.float32, device=depth.device)
    flow = torch.empty(P, 2, dtype=torch.float32, device=depth.device)
    _lib.check(L.b200pose_flow_init(depth.data_ptr(), K.data_ptr(), G.data_ptr(), B, H, W, coords1.data_ptr(),
                                    flow.data_ptr(), _stream()), "b200pose_flow_init")
    return coords1, flow


def update_block(packed: torch.Tensor, net: torch.Tensor, xbuf: torch.Tensor, corr: torch.Tensor,
                 coords1: torch.Tensor, flow: torch.Tensor, B: int, h: int, w: int,
                 flags: int = DEFAULT_FLAGS) -> Tuple[torch.Tensor, torch.Tensor]:
    """In place on net / coords1 / flow; returns (mask [P,576], dflow [P,2])."""
    L = _lib.lib()
    for t, n in ((net, "net"), (xbuf, "xbuf"), (corr, "corr"), (coords1, "coords1"), (flow, "flow")):
        _chk(t, n)
    P = B * h * w
    mask = torch.empty(P, 576, dtype=torch.float32, device=net.device)
    dflow = torch.empty(P, 2, dtype=torch.float32, device=net.device)
    nb = L.b200pose_update_workspace_bytes(B, h, w)
    ws = _ws(nb, net.device)
    _lib.check(L.b200pose_update_block(packed.data_ptr(), net.data_ptr(), xbuf.data_ptr(), corr.data_ptr(),
                                       coords1.data_ptr(), flow.data_ptr(), mask.data_ptr(), dflow.data_ptr(),
                                       B, h, w, int(flags), ws.data_ptr(), nb, _stream()), "b200pose_update_block")
    return mask, dflow


def conv_layer_info(layer: int):
    L = _lib.lib()
    v = [C.c_int() for _ in range(5)]
    _lib.check(L.b200pose_conv_layer_info(layer, *[C.byref(x) for x in v]), "b200pose_conv_layer_info")
    return tuple(x.value for x in v)      # cin0, cin1, cout, kh, kw


def conv_layer(packed: torch.Tensor, layer: int, in0: torch.Tensor, in1: Optional[torch.Tensor], B: int, h: int, w: int,
               flags: int = DEFAULT_FLAGS, workspace: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """One update-block convolution (+bias, no activation) on PXC inputs; returns [P, roundup(cout,4)]."""
    L = _lib.lib()
    _chk(in0, "in0")
    cin0, cin1, cout, kh, kw = conv_layer_info(layer)
    P = B * h * w
    if out is None:
        out = torch.empty(P, (cout + 3) // 4 * 4, dtype=torch.float32, device=in0.device)
    nb = L.b200pose_conv_layer_workspace_bytes(B, h, w)
    if workspace is None:
        workspace = _ws(nb, in0.device)
    _lib.check(L.b200pose_conv_layer(packed.data_ptr(), layer, in0.data_ptr(), in0.shape[1], _p(in1),
                                     in1.shape[1] if in1 is not None else 0, out.data_ptr(), B, h, w, int(flags),
                                     workspace.data_ptr(), nb, _stream()), "b200pose_conv_layer")
    return out


def upsample_weight(flow: torch.Tensor, mask: torch.Tensor, geofea1: Optional[torch.Tensor],
                    geofea2: Optional[torch.Tensor], depth: Optional[torch.Tensor], sigma: float,
                    B: int, H: int, W: int, want_flow_up: bool = True):
    L = _lib.lib()
    _chk(flow, "flow"); _chk(mask, "mask")
    dev = flow.device
    flow_up = torch.empty(B, 2, H, W, dtype=torch.float32, device=dev) if want_flow_up else None
    target = torch.empty(B, H, W, 2, dtype=torch.float32, device=dev)
    weight = None
    Cg = 0
    if geofea1 is not None:
        _chk(geofea1, "geofea1"); _chk(geofea2, "geofea2"); _chk(depth, "depth")
        Cg = geofea1.shape[1]
        weight = torch.empty(B, H, W, dtype=torch.float32, device=dev)
    _lib.check(L.b200pose_upsample_weight(flow.data_ptr(), mask.data_ptr(), _p(geofea1), _p(geofea2), _p(depth),
                                          float(sigma), B, Cg, H, W, _p(flow_up), target.data_ptr(), _p(weight),
                                          _stream()), "b200pose_upsample_weight")
    return flow_up, target, weight


def lm_solve(depth: torch.Tensor, target: torch.Tensor, weight: torch.Tensor, K: torch.Tensor, G: torch.Tensor,
             n_steps: int, ep_lmbda: float = EP_LMBDA, lm_lmbda: float = LM_LMBDA, taps: bool = False,
             depth_offset: float = 0.0):
    """Mirrors SE3Sequence.reprojction_optim(target, weight, depth, intrinsics, num_iters): ``depth`` is what the
    reference passes there (syn_depth + 1e-5) unless depth_offset is given.  G [B,4,4] is updated in place.  With taps=True also returns (H [n,B,6,6] f64, b [n,B,6] f64, delta [n,B,6])."""
    L = _lib.lib()
    for t, n in ((depth, "depth"), (target, "target"), (weight, "weight"), (K, "K"), (G, "G")):
        _chk(t, n)
    B, H, W = depth.shape
    dev = depth.device
    Ho = bo = do = None
    if taps:
        Ho = torch.zeros(n_steps, B, 6, 6, dtype=torch.float64, device=dev)
        bo = torch.zeros(n_steps, B, 6, dtype=torch.float64, device=dev)
        do = torch.zeros(n_steps, B, 6, dtype=torch.float32, device=dev)
    nb = L.b200pose_lm_workspace_bytes(B, H, W)
    ws = _ws(nb, dev)
    _lib.check(L.b200pose_lm_solve(depth.data_ptr(), target.data_ptr(), weight.data_ptr(), K.data_ptr(), G.data_ptr(),
                                   B, H, W, float(depth_offset), n_steps, float(ep_lmbda), float(lm_lmbda), _p(Ho), _p(bo), _p(do),
                                   ws.data_ptr(), nb, _stream()), "b200pose_lm_solve")
    return (G, Ho, bo, do) if taps else G


ENCODER_KEYS: List[str] = (["fnet.conv1"] + [f"fnet.layer1.{b}.conv{c}" for b in (0, 1) for c in (1, 2)] +
                           ["fnet.layer2.0.conv1", "fnet.layer2.0.conv2", "fnet.layer2.0.downsample.0", "fnet.layer2.1.conv1", "fnet.layer2.1.conv2",
                            "fnet.layer3.0.conv1", "fnet.layer3.0.conv2", "fnet.layer3.0.downsample.0", "fnet.layer3.1.conv1", "fnet.layer3.1.conv2",
                            "fnet.conv2"])


def encoder_pack_weights(state: Dict[str, torch.Tensor], device="cuda") -> torch.Tensor:
    """state: ImageFeaEncoder state dict (keys 'fnet.conv1.weight', ...; weights/img_fea_enc.pth).  Returns the packed blob."""
    _need_cuda()
    L = _lib.lib()
    tens = []
    for k in ENCODER_KEYS:
        for sfx in (".weight", ".bias"):
            tens.append(state[k + sfx].detach().to(device=device, dtype=torch.float32).contiguous())
    arr = (C.c_void_p * len(tens))(*[t.data_ptr() for t in tens])
    blob = _ws(L.b200pose_encoder_packed_weights_bytes(), device)
    _lib.check(L.b200pose_encoder_pack_weights(arr, blob.data_ptr(), _stream()), "b200pose_encoder_pack_weights")
    torch.cuda.current_stream().synchronize()      # `tens` must outlive the packing kernels
    return blob


def image_encoder(packed: torch.Tensor, image1: torch.Tensor, image2: torch.Tensor, workspace: Optional[torch.Tensor] = None):
    """ImageFeaEncoder.forward(image1, image2) (reference model/CFNet.py:39-49) in the library's kernels: [B,3,H,W] x 2 ->
    (fmap1, fmap2) [B,256,H/8,W/8]."""
    _need_cuda()
    L = _lib.lib()
    _chk(image1, "image1"); _chk(image2, "image2")
    B, Ci, H, W = image1.shape
    if Ci != 3 or image2.shape != image1.shape:
        raise ValueError("image_encoder: two [B,3,H,W] images")
    dev = image1.device
    f1 = torch.empty(B, 256, H // 8, W // 8, dtype=torch.float32, device=dev)
    f2 = torch.empty(B, 256, H // 8, W // 8, dtype=torch.float32, device=dev)
    nb = L.b200pose_encoder_workspace_bytes(B, H, W)
    if workspace is None or workspace.numel() < nb:
        workspace = _ws(nb, dev)
    _lib.check(L.b200pose_image_encoder(packed.data_ptr(), image1.data_ptr(), image2.data_ptr(), B, H, W, f1.data_ptr(), f2.data_ptr(),
                                        workspace.data_ptr(), workspace.numel(), _stream()), "b200pose_image_encoder")
    return f1, f2


def zoom_crop(pc_depth: torch.Tensor, K: torch.Tensor, T: torch.Tensor, image: Optional[torch.Tensor],
              geofea: Optional[torch.Tensor], out_hw, margin_ratio: float = 0.4, channels_last: bool = False,
              want_theta: bool = False):
    """Device-side PoseRefiner.gen_zoom_crop_grids + the two grid_sample crops (b200pose_zoom_crop).  pc_depth [B,H,W]
    (foreground = > 0), K [B,3,3], T [B,4,4], image [B,Ci,H,W], geofea [B,Cg,H,W].  Returns dict(image_crop, geofea_crop
    ([B,Cg,Hc,Wc] or, channels_last, [B,Hc*Wc,32]), K_crop, theta)."""
    _need_cuda()
    L = _lib.lib()
    _chk(pc_depth, "pc_depth"); _chk(K, "K"); _chk(T, "T")
    B, H, W = pc_depth.shape
    Hc, Wc = int(out_hw[0]), int(out_hw[1])
    dev = pc_depth.device
    Ci = Cg = 0
    ic = gc = None
    if image is not None:
        _chk(image, "image"); Ci = image.shape[1]
        ic = torch.empty(B, Ci, Hc, Wc, dtype=torch.float32, device=dev)
    if geofea is not None:
        _chk(geofea, "geofea"); Cg = geofea.shape[1]
        gc = torch.empty((B, Hc * Wc, Cg) if channels_last else (B, Cg, Hc, Wc), dtype=torch.float32, device=dev)
    Kc = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
    th = torch.empty(B, 2, 3, dtype=torch.float32, device=dev) if want_theta else None
    nb = L.b200pose_zoom_crop_workspace_bytes(B)
    ws = _ws(nb, dev)
    _lib.check(L.b200pose_zoom_crop(pc_depth.data_ptr(), K.data_ptr(), T.data_ptr(), _p(image), _p(geofea), B, Ci, Cg, H, W, Hc, Wc,
                                    float(margin_ratio), 1 if channels_last else 0, _p(ic), _p(gc), Kc.data_ptr(), _p(th),
                                    ws.data_ptr(), nb, _stream()), "b200pose_zoom_crop")
    return dict(image_crop=ic, geofea_crop=gc, K_crop=Kc, theta=th)


def lm_backward(depth: torch.Tensor, target: torch.Tensor, weight: torch.Tensor, K: torch.Tensor, G: torch.Tensor,
                grad_delta: torch.Tensor, ep_lmbda: float = EP_LMBDA, lm_lmbda: float = LM_LMBDA, depth_offset: float = 0.0):
    """Gradients of ONE LM step with respect to target [B,H,W,2] and weight [B,H,W] given dL/d(delta) [B,6]
    (b200pose_lm_backward; reference autograd through geometry/cholesky.py:19-28).  G is the pose entering the step."""
    L = _lib.lib()
    for t, n in ((depth, "depth"), (target, "target"), (weight, "weight"), (K, "K"), (G, "G"), (grad_delta, "grad_delta")):
        _chk(t, n)
    B, H, W = depth.shape
    gt = torch.empty(B, H, W, 2, dtype=torch.float32, device=depth.device)
    gw = torch.empty(B, H, W, dtype=torch.float32, device=depth.device)
    nb = L.b200pose_lm_backward_workspace_bytes(B, H, W)
    ws = _ws(nb, depth.device)
    _lib.check(L.b200pose_lm_backward(depth.data_ptr(), target.data_ptr(), weight.data_ptr(), K.data_ptr(), G.data_ptr(),
                                      grad_delta.data_ptr(), B, H, W, float(depth_offset), float(ep_lmbda), float(lm_lmbda),
                                      gt.data_ptr(), gw.data_ptr(), ws.data_ptr(), nb, _stream()), "b200pose_lm_backward")
    return gt, gw


class LMStep(torch.autograd.Function):
    """One differentiable LM step on the library's kernels: forward = b200pose_lm_solve(n_steps = 1) returning the clamped
    update delta [B,6] (and updating a copy of G); backward = b200pose_lm_backward.  Mirrors the autograd path of the
    reference's reprojction_optim(num_iters = 1) for target and weight (depth, K and the entering pose are constants)."""

    @staticmethod
    def forward(ctx, depth, target, weight, K, G, ep_lmbda, lm_lmbda):
        Gc = G.clone()
        _, _, _, delta = lm_solve(depth, target.contiguous(), weight.contiguous(), K, Gc, 1, ep_lmbda, lm_lmbda, taps=True)
        ctx.save_for_backward(depth, target, weight, K, G)
        ctx.lmb = (ep_lmbda, lm_lmbda)
        ctx.mark_non_differentiable(Gc)
        return delta[0], Gc

    @staticmethod
    def backward(ctx, grad_delta, _grad_G):
        depth, target, weight, K, G = ctx.saved_tensors
        gt, gw = lm_backward(depth, target.contiguous(), weight.contiguous(), K, G, grad_delta.contiguous().float(), *ctx.lmb)
        return None, gt, gw, None, None, None, None


def lm_step_autograd(depth, target, weight, K, G, ep_lmbda: float = EP_LMBDA, lm_lmbda: float = LM_LMBDA):
    """Differentiable single LM step: returns (delta [B,6], G_new [B,4,4]); gradients flow to target and weight."""
    return LMStep.apply(depth, target, weight, K, G, float(ep_lmbda), float(lm_lmbda))


def cholesky_solve(H: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """geometry/cholesky.py `solve` for 6x6 fp64 systems: H [B,6,6], b [B,6] -> x [B,6] fp32 (NaN -> 0, clamp +-1)."""
    L = _lib.lib()
    _chk(H, "H", torch.float64); _chk(b, "b", torch.float64)
    B = H.shape[0]
    if H.shape != (B, 6, 6) or b.shape != (B, 6):
        raise ValueError("cholesky_solve: H [B,6,6], b [B,6]")
    x = torch.empty(B, 6, dtype=torch.float32, device=H.device)
    _lib.check(L.b200pose_cholesky_solve(H.data_ptr(), b.data_ptr(), x.data_ptr(), B, _stream()), "b200pose_cholesky_solve")
    return x


def se3_retract(delta: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    """G <- exp(delta) G in place (SE3.increment); delta [B,6], G [B,4,4]."""
    L = _lib.lib()
    _chk(delta, "delta"); _chk(G, "G")
    _lib.check(L.b200pose_se3_retract(delta.data_ptr(), G.data_ptr(), G.shape[0], _stream()), "b200pose_se3_retract")
    return G


METRIC_COLS = 16
LINEMOD_K = ((572.4114, 0.0, 325.2611), (0.0, 573.57043, 242.04899), (0.0, 0.0, 1.0))   # data/linemod/linemod_config.py:23-25


def pose_metrics(T_pred: torch.Tensor, T_gt: torch.Tensor, pts: torch.Tensor, diameter: torch.Tensor,
                 K: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,16] per-object metric row of include/b200pose.h (ADD, ADD-S, angles, translation, 2-D projection, flags).
    Replaces utils/eval_metric.py:102-192 + thirdparty/nn (brute-force nearest neighbour) + geometric.py:36-40.
    K [B,3,3] or [3,3]; default linemod_K, which is what the reference's evaluator projects with (eval_metric.py:338)."""
    L = _lib.lib()
    T_pred, T_gt, pts, diameter = (t.contiguous().float() for t in (T_pred, T_gt, pts, diameter))
    B, n_pts = pts.shape[0], pts.shape[1]
    if K is None:
        K = torch.tensor(LINEMOD_K, dtype=torch.float32, device=pts.device)
    K = K.to(device=pts.device, dtype=torch.float32)
    if K.dim() == 2:
        K = K[None].expand(B, 3, 3)
    K = K.contiguous()
    for t, n in ((T_pred, "T_pred"), (T_gt, "T_gt"), (pts, "pts"), (diameter, "diameter"), (K, "K")):
        _chk(t, n)
    if T_pred.shape != (B, 4, 4) or T_gt.shape != (B, 4, 4) or pts.shape[2] != 3 or diameter.shape != (B,) or K.shape != (B, 3, 3):
        raise ValueError("pose_metrics: T [B,4,4], pts [B,n,3], diameter [B], K [B,3,3]")
    out = torch.empty(B, METRIC_COLS, dtype=torch.float32, device=pts.device)
    nb = L.b200pose_pose_metrics_workspace_bytes(B, n_pts)
    ws = _ws(nb, pts.device)
    _lib.check(L.b200pose_pose_metrics(T_pred.data_ptr(), T_gt.data_ptr(), pts.data_ptr(), diameter.data_ptr(), K.data_ptr(),
                                       B, n_pts, out.data_ptr(), ws.data_ptr(), nb, _stream()), "b200pose_pose_metrics")
    return out


class RefineWorkspace:
    """Caller-owned scratch for refine_iters (re-used across calls of the same shape)."""

    def __init__(self, B: int, H: int, W: int, device="cuda"):
        self.key = (B, H, W)
        self.nbytes = _lib.lib().b200pose_refine_workspace_bytes(B, H, W)
        self.buf = _ws(self.nbytes, device)


def refine_iters(packed: torch.Tensor, fmap1, fmap2, context, geofea1, geofea2, depth, K, G, sigma: float,
                 n_iters: int, n_lm: int, ep_lmbda: float = EP_LMBDA, lm_lmbda: float = LM_LMBDA,
                 workspace: Optional[RefineWorkspace] = None, want_flows: bool = False, want_weight: bool = False,
                 flags: int = DEFAULT_FLAGS):
    """The fused inner loop (b200pose_refine_iters).  depth [B,H,W]; G [B,4,4] updated in place.
    Returns dict(G=..., flow_first=..., flow_last=..., weight=...)."""
    _need_cuda()
    L = _lib.lib()
    for t, n in ((fmap1, "fmap1"), (fmap2, "fmap2"), (context, "context"), (geofea1, "geofea1"), (geofea2, "geofea2"),
                 (depth, "depth"), (K, "K"), (G, "G")):
        _chk(t, n)
    B, H, W = depth.shape
    Cg = geofea1.shape[1]
    dev = depth.device
    if workspace is None or workspace.key != (B, H, W):
        workspace = RefineWorkspace(B, H, W, dev)
    ff = torch.empty(B, 2, H, W, dtype=torch.float32, device=dev) if want_flows else None
    fl = torch.empty(B, 2, H, W, dtype=torch.float32, device=dev) if want_flows else None
    wl = torch.empty(B, H, W, dtype=torch.float32, device=dev) if want_weight else None
    _lib.check(L.b200pose_refine_iters(packed.data_ptr(), fmap1.data_ptr(), fmap2.data_ptr(), context.data_ptr(),
                                       geofea1.data_ptr(), geofea2.data_ptr(), depth.data_ptr(), K.data_ptr(),
                                       G.data_ptr(), float(sigma), B, Cg, H, W, n_iters, n_lm, float(ep_lmbda),
                                       float(lm_lmbda), int(flags), _p(ff), _p(fl), _p(wl), workspace.buf.data_ptr(),
                                       workspace.nbytes, _stream()), "b200pose_refine_iters")
    return dict(G=G, flow_first=ff, flow_last=fl, weight=wl, workspace=workspace)


def host_staging(B: int, H: int, W: int) -> torch.Tensor:
    """Pinned staging buffer for refine_iters_host(..., staging=...): b200pose_refine_host_staging_bytes(B, H, W)."""
    n = _lib.lib().b200pose_refine_host_staging_bytes(B, H, W)
    return torch.empty(n // 4, dtype=torch.float32).pin_memory()


def context_gather_texels(context: torch.Tensor, threads: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Host-only (no CUDA): context [B,256,H,W] CPU float32 -> the texels the 1/8 resample reads, [B,256,(H/8)*(W/8),4]
    (b200pose_context_gather_texels; the layout FLAG_CONTEXT_TEXELS names)."""
    if context.is_cuda or context.dtype != torch.float32 or not context.is_contiguous() or context.dim() != 4 or context.shape[1] != 256:
        raise ValueError("context_gather_texels expects a contiguous CPU float32 [B,256,H,W] tensor")
    B, _, H, W = context.shape
    if out is None:
        out = torch.empty(B, 256, (H // 8) * (W // 8), 4, dtype=torch.float32)
    _lib.check(_lib.lib().b200pose_context_gather_texels(context.data_ptr(), B, H, W, out.data_ptr(), int(threads)),
               "b200pose_context_gather_texels")
    return out


def refine_iters_host(packed: torch.Tensor, fmap1, fmap2, context, geofea1, geofea2, depth, K, G, sigma: float,
                      n_iters: int, n_lm: int, ep_lmbda: float = EP_LMBDA, lm_lmbda: float = LM_LMBDA,
                      scratch: Optional[torch.Tensor] = None, flags: int = DEFAULT_FLAGS,
                      staging: Optional[torch.Tensor] = None, threads: int = 0):
    """Host-buffer entry point (b200pose_refine_iters_host[2]): all tensors are CPU float32 (pinned for full
    copy speed); G [B,4,4] is overwritten on the host.  Synchronises the stream before returning.
    staging (host_staging(B,H,W), pinned) switches on the host-side gather of the context texels by `threads` workers."""
    _need_cuda()
    L = _lib.lib()
    for t in (fmap1, fmap2, context, geofea1, geofea2, depth, K, G):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("refine_iters_host expects contiguous CPU float32 tensors")
    B, H, W = depth.shape
    Cg = geofea1.shape[1]
    nb = L.b200pose_refine_host_scratch_bytes(B, Cg, H, W)
    if scratch is None or scratch.numel() < nb:
        scratch = _ws(nb, packed.device)
    _lib.check(L.b200pose_refine_iters_host2(packed.data_ptr(), fmap1.data_ptr(), fmap2.data_ptr(), context.data_ptr(),
                                             geofea1.data_ptr(), geofea2.data_ptr(), depth.data_ptr(), K.data_ptr(),
                                             G.data_ptr(), float(sigma), B, Cg, H, W, n_iters, n_lm, float(ep_lmbda),
                                             float(lm_lmbda), int(flags), scratch.data_ptr(), scratch.numel(),
                                             staging.data_ptr() if staging is not None else None,
                                             staging.numel() * staging.element_size() if staging is not None else 0,
                                             int(threads), _stream()),
               "b200pose_refine_iters_host2")
    return G, scratch


def launch_count(B: int, H: int, W: int, n_iters: int, n_lm: int) -> int:
    return _lib.lib().b200pose_refine_launch_count(B, H, W, n_iters, n_lm)

"""Drop-in replacement of the reference's ``model.PoseRefiner.PoseRefiner``.

Same constructor signature, same parameter names/shapes (``sigma.0``, ``cf_net.update_block.*``; plus
``image_fea_enc.fnet.*``: the encoder mirror of ``rnnpose_b200.encoder``) so that the reference's shape-matched checkpoint
loader (reference tools/eval.py:386-408) fills them, same ``forward(image, Ts, intrinsics, fea_3d, Tj_gt, obj_cls,
geofea_3d, geofea_2d)`` and the same return dict (reference model/PoseRefiner.py:366-376).

What runs where:
  * the OUTER render loop (reference PoseRefiner.py:239-313) stays PyTorch: it calls the injected renderer (out of scope,
    SURVEY section 2) and the feature encoder (by default the library's own RAFT encoder kernels behind
    ``rnnpose_b200.encoder.ImageFeaEncoder``).  The crop geometry of
    ``gen_zoom_crop_grids`` / ``get_affine_transformation`` (reference :145-213) and the two grid_sample crops run in the
    library's zoom-crop kernels (``ops.zoom_crop``: bounding box by atomics, closed-form axis-aligned affine, fused
    resample), removing the reference's numpy/cv2 host round-trip; ``zoom_crop_params`` below is
    the same geometry in torch ops, kept as a checker;
  * the INNER loop (reference :315-362) is one call into libb200pose.so (``ops.refine_iters``), natively batched
    (the reference only supports batch size 1, SURVEY finding 1).
Inference only (the reference's training loss is out of scope).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .se3 import SE3Sequence

EPS = 1e-5   # reference model/PoseRefiner.py:21


class _UpdateBlockParams(nn.Module):
    """Parameter holder with the state-dict layout of ``BasicUpdateBlock`` (reference thirdparty/raft/update.py:164-177).
    The nn.Conv2d modules are never called: their tensors are packed for the CUDA kernels."""

    def __init__(self):
        super().__init__()
        enc = nn.Module()
        enc.convc1 = nn.Conv2d(324, 256, 1)
        enc.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        enc.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        enc.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        enc.conv = nn.Conv2d(256, 126, 3, padding=1)
        self.encoder = enc
        gru = nn.Module()
        for n in ("z", "r", "q"):
            setattr(gru, f"conv{n}1", nn.Conv2d(384, 128, (1, 5), padding=(0, 2)))
        for n in ("z", "r", "q"):
            setattr(gru, f"conv{n}2", nn.Conv2d(384, 128, (5, 1), padding=(2, 0)))
        self.gru = gru
        fh = nn.Module()
        fh.conv1 = nn.Conv2d(128, 256, 3, padding=1)
        fh.conv2 = nn.Conv2d(256, 2, 3, padding=1)
        self.flow_head = fh
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 576, 1))


class _CFNetParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.update_block = _UpdateBlockParams()


def zoom_crop_params(fg_mask: torch.Tensor, K: torch.Tensor, T: torch.Tensor, out_hw, margin_ratio: float = 0.4):
    """Device-side restatement of reference PoseRefiner.gen_zoom_crop_grids/get_affine_transformation (:145-213).
    fg_mask [B,1,H,W] bool, K [B,3,3], T [B,4,4].  Returns (theta [B,2,3] for F.affine_grid, K_crop [B,3,3])."""
    B, _, H, W = fg_mask.shape
    oh, ow = out_hw
    ratio = float(H) / float(W)
    c = torch.matmul(K, T[:, :3, 3:])                     # projected model centre (:204-205)
    cxy = c[:, :2, 0] / c[:, 2:3, 0]
    zx, zy = cxy[:, 0], cxy[:, 1]
    m = fg_mask[:, 0]
    cols = m.any(dim=1); rows = m.any(dim=2)              # [B,W], [B,H]
    has = cols.any(dim=1)
    xs = torch.arange(W, device=m.device); ys = torch.arange(H, device=m.device)
    big = 10 ** 6
    x_min = torch.where(cols, xs, big).min(dim=1).values; x_max = torch.where(cols, xs, -big).max(dim=1).values
    y_min = torch.where(rows, ys, big).min(dim=1).values; y_max = torch.where(rows, ys, -big).max(dim=1).values
    zero = torch.zeros_like(x_min)
    x_min = torch.where(has, x_min, zero).float(); x_max = torch.where(has, x_max, zero).float()
    y_min = torch.where(has, y_min, zero).float(); y_max = torch.where(has, y_max, zero).float()
    left, right, up, down = zx - x_min, x_max - zx, zy - y_min, y_max - zy
    crop_h = torch.stack([ratio * right, ratio * left, up, down], dim=1).max(dim=1).values * 2 * (1 + margin_ratio)
    crop_w = crop_h / ratio
    x1, x2 = zx - crop_w / 2, zx + crop_w / 2
    y1, y2 = zy - crop_h / 2, zy + crop_h / 2
    # affine_grid theta: [-1,1]^2 of the output -> normalised input coordinates (cv2.getAffineTransform of an
    # axis-aligned box is this diagonal map)
    nx1, nx2 = x1 * 2 / W - 1, x2 * 2 / W - 1
    ny1, ny2 = y1 * 2 / H - 1, y2 * 2 / H - 1
    theta = torch.zeros(B, 2, 3, device=K.device, dtype=K.dtype)
    theta[:, 0, 0] = (nx2 - nx1) / 2; theta[:, 0, 2] = (nx2 + nx1) / 2
    theta[:, 1, 1] = (ny2 - ny1) / 2; theta[:, 1, 2] = (ny2 + ny1) / 2
    # crop pixel (0..ow-1, 0..oh-1) -> input pixel; K_crop = inv(A) K   (:188-198, :211)
    A = torch.zeros(B, 3, 3, device=K.device, dtype=K.dtype)
    A[:, 0, 0] = (x2 - x1) / (ow - 1); A[:, 0, 2] = x1
    A[:, 1, 1] = (y2 - y1) / (oh - 1); A[:, 1, 2] = y1
    A[:, 2, 2] = 1
    return theta, torch.matmul(torch.linalg.inv(A), K)


class PoseRefiner(nn.Module):
    def __init__(self, cfg, reuse=False, schedule=None, use_regressor=True, is_calibrated=True, bn_is_training=False,
                 is_training=True, renderer=None, image_fea_enc: Optional[nn.Module] = None,
                 render_image_size=None, zoom_crop_size=None):
        super().__init__()
        self.cfg = cfg
        self.legacy = True
        self.sigma = nn.ParameterList([nn.Parameter(torch.ones(1))])
        self.is_calibrated = bool(self._cfg("IS_CALIBRATED", True)) and is_calibrated
        self.is_training = is_training
        self.use_regressor = use_regressor
        if self._cfg("FLOW_NET", "raft") != "raft":
            raise NotImplementedError
        if image_fea_enc is None:
            # the library's own encoder kernels behind the reference's module layout (state-dict keys image_fea_enc.fnet.*,
            # shipped img_fea_enc weights loaded as the reference's constructor does, model/CFNet.py:33-37); any nn.Module with
            # forward(image1, image2) -> (fmap1, fmap2) -- e.g. the reference's ImageFeaEncoder -- can be passed instead
            from .encoder import ImageFeaEncoder
            image_fea_enc = ImageFeaEncoder()
        self.image_fea_enc = image_fea_enc
        self.cf_net = _CFNetParams()
        self.renderer = renderer
        self._render_image_size = render_image_size
        self._zoom_crop_size = zoom_crop_size
        self._packed = None
        self._packed_key = None
        self._ws = None
        self.flags = ops.DEFAULT_FLAGS

    def _cfg(self, key, default=None):
        c = self.cfg
        if isinstance(c, dict):
            return c.get(key, default)
        return getattr(c, key, default)

    def _sizes(self):
        if self._render_image_size is not None and self._zoom_crop_size is not None:
            return tuple(self._render_image_size), tuple(self._zoom_crop_size)
        from config.default import get_cfg                           # reference global config (PoseRefiner.py:226-227)
        b = get_cfg("BASIC")
        return tuple(b.render_image_size), tuple(b.zoom_crop_size)

    def _sigma_value(self) -> float:
        """sigma as a host scalar, read back from the device only when the parameter changes (no sync per render iteration)."""
        p = self.sigma[0]
        key = (p.data_ptr(), p._version)
        if getattr(self, "_sigma_key", None) != key:
            self._sigma_cache, self._sigma_key = float(p.detach()), key
        return self._sigma_cache

    def packed_weights(self) -> torch.Tensor:
        sd = self.cf_net.update_block.state_dict()
        key = tuple((k, v.data_ptr(), v._version, str(v.device)) for k, v in sd.items())
        if self._packed is None or key != self._packed_key:
            dev = self.sigma[0].device
            self._packed = ops.pack_weights({k: v for k, v in sd.items()}, dev)
            self._packed_key = key
        return self._packed

    @torch.no_grad()
    def forward(self, image, Ts, intrinsics, fea_3d=None, Tj_gt=None, obj_cls=None, geofea_3d=None, geofea_2d=None):
        if self.image_fea_enc is None:
            raise RuntimeError("no image feature encoder attached (pass image_fea_enc= or run inside the reference tree)")
        if geofea_3d is None or geofea_2d is None:
            raise NotImplementedError("the reference's with_corr_weight=False branch reads an undefined variable "
                                      "(PoseRefiner.py:346-347); descriptors are required")
        render_size, zoom_size = self._sizes()
        n_render = int(self._cfg("RENDER_ITER_COUNT", 1)); n_iters = int(self._cfg("ITER_COUNT", 4))
        n_lm = int(self._cfg("OPTIM_ITER_COUNT", 1))
        Hc, Wc = zoom_size
        Ti = Ts
        Tij = Ti.copy().identity()
        syn_imgs, syn_depths, Tij_gt = [], [], []
        first_flow = weight = None
        packed = self.packed_weights()
        for _ in range(n_render):
            Ti = Tij * Ti
            Tij = Ti * Ti.inv()                                           # legacy "identity" (:243-244)
            T_mat = Ti.matrix().detach().squeeze(1)
            pc_depth = self.renderer.render_pointcloud(obj_cls, T=T_mat, K=intrinsics.detach(), render_image_size=render_size)
            B = pc_depth.shape[0]
            # zoom-crop on device (reference :145-218,:287,:292): foreground box, crop intrinsics and both crops in one
            # entry; the descriptors leave channels-last, which is what the loop's foreground pipeline reads
            cl = geofea_2d.shape[1] == 32 and ops.get_option("fg_list") != 0 and ops.get_option("fg_pipeline") != 0   # (0 = dense kernels: NCHW)
            zc = ops.zoom_crop(pc_depth[:, 0].contiguous().float(), intrinsics.detach().float().contiguous(), T_mat.float().contiguous(),
                               image.float().contiguous(), geofea_2d.float().contiguous(), (Hc, Wc), channels_last=cl)
            K_crop, image_crop, geofea2_crop = zc["K_crop"], zc["image_crop"], zc["geofea_crop"]
            attr = torch.cat([fea_3d, geofea_3d], dim=-1)
            color, depth = self.renderer(obj_cls, attr, T=T_mat, K=K_crop.detach(), render_image_size=(Hc, Wc), near=0.1,
                                         far=6, render_tex=True)
            depth = depth.clone(); depth[depth == -1] = 0                  # (:138)
            syn_img, cfea, geofea1 = torch.split(color, [3, fea_3d.shape[-1], geofea_3d.shape[-1]], dim=1)
            cfea = cfea * 0.1                                              # (:283)
            syn_depth = self.renderer.render_depth(obj_cls, T=T_mat, K=K_crop.detach(), render_image_size=(Hc, Wc),
                                                   near=0.1, far=6)         # legacy second render (:296-304)
            syn_imgs += [syn_img, image_crop]
            feats1, feats2 = self.image_fea_enc(syn_img, image_crop)
            G = Tij.G[:, 0].contiguous().float().clone()
            if self._ws is None or self._ws.key != (B, Hc, Wc):
                self._ws = ops.RefineWorkspace(B, Hc, Wc, G.device)
            res = ops.refine_iters(packed, feats1.float().contiguous(), feats2.float().contiguous(), cfea.contiguous(),
                                   geofea1.contiguous(), geofea2_crop, syn_depth[:, 0].contiguous(),
                                   K_crop, G, self._sigma_value(), n_iters, n_lm, workspace=self._ws,
                                   want_flows=True, want_weight=True,
                                   flags=self.flags | (ops.FLAG_GEO2_CHANNELS_LAST if cl else 0))
            Tij = SE3Sequence(matrix=G[:, None])
            if first_flow is None:
                first_flow = [res["flow_first"]]
            weight = res["weight"]
            for _i in range(n_iters):
                syn_depths.append(syn_depth)
                if Tj_gt is not None:
                    Tij_gt.append((Tj_gt * Ti.inv()).copy(stop_gradients=True))
        Ti = Tij * Ti
        return {
            "Tij": Tij, "Ti_pred": Ti, "intrinsics": intrinsics, "flow": first_flow, "vmask": syn_depth > 0,
            "weight": weight[:, None, None], "syn_depth": syn_depths,
            "syn_img": syn_imgs + [image_crop, cfea[:, :3] * 10, geofea1[:, :3],
                        (geofea2_crop.view(B, Hc, Wc, -1).permute(0, 3, 1, 2) if cl else geofea2_crop)[:, :3]],
            "Tij_gt": Tij_gt,
        }

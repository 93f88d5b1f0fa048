"""Seeded synthetic (rendered-reference, target) crop pairs for the refinement inner loop.

SURVEY.md §8(d) "Synthetic inputs": an analytic ellipsoid with a procedural surface texture, a
ground-truth pose, a perturbed initial pose (noise scales of ``sample_poses``,
reference ``model/RNNPose.py:96-124``), ``linemod_K`` intrinsics
(reference ``data/linemod/linemod_config.py:23-25``) zoomed onto the object the way
``PoseRefiner.gen_zoom_crop_grids`` does (reference ``model/PoseRefiner.py:202-213``).

Everything here is evaluated with exactly-rounded IEEE-754 element-wise operations only
(+, -, *, /, sqrt, floor, abs in float64; no BLAS, no libm transcendental), so the same seed gives
bit-identical inputs in the build container and on the GPU box.  That is what lets
``tests/golden`` store only the (small) reference outputs.

The module doubles as the injectable *renderer* of the reference contract:
``AnalyticRenderer`` implements ``render_pointcloud`` / ``__call__`` / ``render_depth``.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

LINEMOD_K = np.array([[572.4114, 0.0, 325.2611],
                      [0.0, 573.57043, 242.04899],
                      [0.0, 0.0, 1.0]], dtype=np.float64)

CTX_DIM = 256     # context feature channels rendered from fea_3d
GEO_DIM = 32      # descriptor channels (geofea_3d / geofea_2d)
TAN_7P5_DEG = 0.13165249758739583   # tan(7.5 deg): Cayley parameter of a 15 deg rotation


def _tri(x: np.ndarray) -> np.ndarray:
    """Triangle wave with period 1 and range [-1, 1]; exact IEEE ops only."""
    return 4.0 * np.abs(x - np.floor(x + 0.5)) - 1.0


def _unit_normal(rng: np.random.Generator, n: int) -> np.ndarray:
    """Irwin-Hall(4) approximation of N(0,1) (adds only, reproducible everywhere)."""
    u = rng.random((n, 4))
    return ((u[:, 0] + u[:, 1]) + (u[:, 2] + u[:, 3]) - 2.0) * 1.7320508075688772


def _quat_to_R(q: np.ndarray) -> np.ndarray:
    q = q / np.sqrt((q * q).sum())
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


def _cayley(c: np.ndarray) -> np.ndarray:
    """Rotation from a Cayley vector c = tan(theta/2) * axis (rational; no sin/cos)."""
    x, y, z = c
    n = 1.0 + x * x + y * y + z * z
    return np.array([
        [1 + x * x - y * y - z * z, 2 * (x * y - z), 2 * (x * z + y)],
        [2 * (x * y + z), 1 - x * x + y * y - z * z, 2 * (y * z - x)],
        [2 * (x * z - y), 2 * (y * z + x), 1 - x * x - y * y + z * z]], dtype=np.float64) / n


def _mat3mul(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """3x3 product as explicit (a*b + a*b) + a*b sums: no BLAS/FMA, bit-reproducible."""
    C = np.zeros((3, 3), dtype=np.float64)
    for i in range(3):
        for j in range(3):
            C[i, j] = (A[i, 0] * B[0, j] + A[i, 1] * B[1, j]) + A[i, 2] * B[2, j]
    return C


@dataclasses.dataclass
class Scene:
    """One object: shape, texture parameters, poses, intrinsics."""
    idx: int
    axes: np.ndarray            # (3,) ellipsoid semi-axes [m]
    T_gt: np.ndarray            # (4,4) object->camera, observed frame
    T_init: np.ndarray          # (4,4) object->camera, initial estimate (rendered frame)
    K_full: np.ndarray          # (3,3) full-image intrinsics (linemod_K)
    K_crop: np.ndarray          # (3,3) zoomed crop intrinsics
    H: int
    W: int
    tex: Dict[str, np.ndarray]  # per-channel texture parameters
    occluder: Optional[Sequence[int]] = None   # (y0,y1,x0,x1) rectangle zeroed in the target

    @property
    def diameter(self) -> float:
        return float(2.0 * self.axes.max())


def _texture_params(rng: np.random.Generator, channels: int, fmax: float) -> Dict[str, np.ndarray]:
    # frequencies on a half-integer lattice so channels decorrelate; phases uniform
    k = np.floor(rng.random((channels, 3)) * (2.0 * fmax + 1.0)) * 0.5 - fmax * 0.5
    k[np.abs(k).sum(axis=1) == 0.0, 0] = 0.5
    phase = rng.random(channels)
    return {"k": k, "phase": phase}


def make_scene(idx: int, H: int = 240, W: int = 320, seed: int = 1234,
               occlude: bool = False, rot_sigma: float = 1.0, fill: float = 0.85) -> Scene:
    """Scene ``idx`` of the synthetic benchmark set (generator seed ``seed + idx``)."""
    rng = np.random.Generator(np.random.PCG64(seed + idx))
    axes = 0.03 + 0.05 * rng.random(3)
    R_gt = _quat_to_R(rng.random(4) * 2.0 - 1.0 + np.array([1e-3, 0, 0, 0]))
    t_gt = np.array([rng.random() * 0.1 - 0.05, rng.random() * 0.1 - 0.05, 0.7 + 0.5 * rng.random()])
    # perturbation: ~15 deg/axis rotation noise, (1,1,5) cm translation noise
    n = _unit_normal(rng, 6)
    c = np.clip(n[:3] * rot_sigma, -2.5, 2.5) * TAN_7P5_DEG
    dR = _cayley(c)
    dt = n[3:] * np.array([0.01, 0.01, 0.05]) * rot_sigma
    R_init = _mat3mul(dR, R_gt)
    t_init = t_gt + dt
    T_gt = np.eye(4); T_gt[:3, :3] = R_gt; T_gt[:3, 3] = t_gt
    T_init = np.eye(4); T_init[:3, :3] = R_init; T_init[:3, 3] = t_init

    # zoom-crop intrinsics: centre on the projected object origin of the initial pose and scale so
    # that the bounding sphere fills `fill` of the crop height (cf. margin_ratio=0.4 in the
    # reference crop, PoseRefiner.py:145-200)
    K = LINEMOD_K.copy()
    u0 = K[0, 0] * t_init[0] / t_init[2] + K[0, 2]
    v0 = K[1, 1] * t_init[1] / t_init[2] + K[1, 2]
    r_px = K[1, 1] * axes.max() / t_init[2]
    s = (fill * H * 0.5) / r_px
    K_crop = np.array([[s * K[0, 0], 0.0, s * (K[0, 2] - u0) + 0.5 * (W - 1)],
                       [0.0, s * K[1, 1], s * (K[1, 2] - v0) + 0.5 * (H - 1)],
                       [0.0, 0.0, 1.0]])
    # poses / intrinsics are float32 quantities in the pipeline: make them exactly representable so
    # that rendering from the float32 tensors (AnalyticRenderer) and from the Scene agree bit-for-bit
    T_gt = T_gt.astype(np.float32).astype(np.float64)
    T_init = T_init.astype(np.float32).astype(np.float64)
    K_crop = K_crop.astype(np.float32).astype(np.float64)
    tex = {
        "img": _texture_params(rng, 3, 7.0),
        "ctx": _texture_params(rng, CTX_DIM, 3.0),
        "geo": _texture_params(rng, GEO_DIM, 2.0),
        "bg": _texture_params(rng, 3, 5.0),
    }
    occ = None
    if occlude:
        oh = int(H * (0.2 + 0.2 * rng.random()))
        ow = int(W * (0.2 + 0.2 * rng.random()))
        y0 = int((H - oh) * rng.random()); x0 = int((W - ow) * rng.random())
        occ = (y0, y0 + oh, x0, x0 + ow)
    return Scene(idx, axes, T_gt, T_init, K, K_crop, H, W, tex, occ)


def _raycast(scene: Scene, T: np.ndarray, K: np.ndarray, H: int, W: int):
    """Analytic ray/ellipsoid intersection. Returns depth (H,W) (0 = background) and the unit-sphere
    surface coordinates s = Xo / axes, shape (H,W,3)."""
    R = T[:3, :3]; t = T[:3, 3]
    u = np.arange(W, dtype=np.float64)[None, :]
    v = np.arange(H, dtype=np.float64)[:, None]
    dx = (u - K[0, 2]) / K[0, 0] + 0.0 * v
    dy = (v - K[1, 2]) / K[1, 1] + 0.0 * u
    # p = R^T d, q = R^T t (explicit sums; no BLAS)
    p = [R[0, i] * dx + R[1, i] * dy + R[2, i] for i in range(3)]
    q = [R[0, i] * t[0] + R[1, i] * t[1] + R[2, i] * t[2] for i in range(3)]
    ia = 1.0 / (scene.axes * scene.axes)
    pAp = p[0] * p[0] * ia[0] + p[1] * p[1] * ia[1] + p[2] * p[2] * ia[2]
    pAq = p[0] * q[0] * ia[0] + p[1] * q[1] * ia[1] + p[2] * q[2] * ia[2]
    qAq = q[0] * q[0] * ia[0] + q[1] * q[1] * ia[1] + q[2] * q[2] * ia[2]
    disc = pAq * pAq - pAp * (qAq - 1.0)
    hit = disc > 0.0
    Z = (pAq - np.sqrt(np.where(hit, disc, 0.0))) / pAp
    hit &= Z > 0.05
    Z = np.where(hit, Z, 0.0)
    s = np.stack([(Z * p[i] - q[i]) / scene.axes[i] for i in range(3)], axis=-1)
    s = np.where(hit[..., None], s, 0.0)
    return Z, s, hit


def _field(s: np.ndarray, par: Dict[str, np.ndarray]) -> np.ndarray:
    """(C,H,W) triangle-wave field of the surface coordinate s (H,W,3)."""
    k, ph = par["k"], par["phase"]
    arg = (k[:, 0, None, None] * s[None, ..., 0] + k[:, 1, None, None] * s[None, ..., 1]) \
        + (k[:, 2, None, None] * s[None, ..., 2] + ph[:, None, None])
    return _tri(arg)


def _geo(s: np.ndarray, par: Dict[str, np.ndarray]) -> np.ndarray:
    g = _field(s, par)
    g[0] = 1.0                                   # keeps the norm away from zero
    nrm = np.sqrt((g * g).sum(axis=0, keepdims=True))
    return g / nrm


def render_reference(scene: Scene, T: Optional[np.ndarray] = None, K: Optional[np.ndarray] = None,
                     H: Optional[int] = None, W: Optional[int] = None) -> Dict[str, np.ndarray]:
    """What the reference's mesh renderer produces for the *rendered* frame: colour, context
    features, descriptors and depth (reference ``PoseRefiner.render``, ``PoseRefiner.py:117-142``)."""
    T = scene.T_init if T is None else T
    K = scene.K_crop if K is None else K
    H = scene.H if H is None else H; W = scene.W if W is None else W
    Z, s, hit = _raycast(scene, T, K, H, W)
    m = hit[None].astype(np.float64)
    img = (0.5 + 0.5 * _field(s, scene.tex["img"])) * m
    ctx = _field(s, scene.tex["ctx"]) * m
    geo = _geo(s, scene.tex["geo"]) * m
    return {"depth": Z.astype(np.float32), "img": img.astype(np.float32),
            "ctx": ctx.astype(np.float32), "geo": geo.astype(np.float32), "mask": hit}


def render_observed(scene: Scene, K: Optional[np.ndarray] = None,
                    H: Optional[int] = None, W: Optional[int] = None) -> Dict[str, np.ndarray]:
    """The *observed* (target) frame at the ground-truth pose: colour image and dense 2-D
    descriptors, with a procedural background and an optional occluder rectangle."""
    K = scene.K_crop if K is None else K
    H = scene.H if H is None else H; W = scene.W if W is None else W
    Z, s, hit = _raycast(scene, scene.T_gt, K, H, W)
    u = np.arange(W, dtype=np.float64)[None, :] / 64.0 + 0.0 * np.arange(H)[:, None]
    v = np.arange(H, dtype=np.float64)[:, None] / 64.0 + 0.0 * np.arange(W)[None, :]
    sb = np.stack([u, v, u * v], axis=-1)
    m = hit[None]
    img = np.where(m, 0.5 + 0.5 * _field(s, scene.tex["img"]), 0.35 + 0.15 * _field(sb, scene.tex["bg"]))
    geo = np.where(m, _geo(s, scene.tex["geo"]), _geo(sb, scene.tex["geo"]))
    if scene.occluder is not None and H == scene.H and W == scene.W:
        y0, y1, x0, x1 = scene.occluder
        img[:, y0:y1, x0:x1] = 0.0
        geo[:, y0:y1, x0:x1] = 0.0
    return {"img": img.astype(np.float32), "geo": geo.astype(np.float32), "mask": hit}


def model_points(scene: Scene, n_side: int = 6) -> np.ndarray:
    """Cube-sphere sample of the ellipsoid surface (object frame), for ADD / ADD-S."""
    g = (np.arange(n_side, dtype=np.float64) + 0.5) / n_side * 2.0 - 1.0
    a, b = np.meshgrid(g, g, indexing="ij")
    one = np.ones_like(a)
    faces = [np.stack(f, -1) for f in ((one, a, b), (-one, a, b), (a, one, b), (a, -one, b), (a, b, one), (a, b, -one))]
    p = np.concatenate([f.reshape(-1, 3) for f in faces], axis=0)
    p = p / np.sqrt((p * p).sum(axis=1, keepdims=True))
    return (p * scene.axes[None, :]).astype(np.float32)


def make_batch(indices: Sequence[int], H: int = 240, W: int = 320, seed: int = 1234,
               occlude: bool = False, rot_sigma: float = 1.0,
               with_images: bool = True) -> Dict[str, torch.Tensor]:
    """Inner-loop inputs for scenes ``indices`` as CPU float32 tensors.

    Keys: depth [B,1,H,W], context [B,256,H,W] (already multiplied by 0.1 as at
    ``PoseRefiner.py:283``), geofea1/geofea2 [B,32,H,W], K [B,3,3], T_init/T_gt [B,4,4],
    and (``with_images``) syn_img/obs_img [B,3,H,W] for the (out-of-scope) feature encoder.
    """
    out: Dict[str, List[np.ndarray]] = {k: [] for k in
                                        ("depth", "context", "geofea1", "geofea2", "K", "T_init", "T_gt",
                                         "syn_img", "obs_img", "diameter")}
    for i in indices:
        sc = make_scene(i, H, W, seed, occlude, rot_sigma)
        ref = render_reference(sc)
        obs = render_observed(sc)
        out["depth"].append(ref["depth"][None])
        out["context"].append((ref["ctx"] * np.float32(0.1)))
        out["geofea1"].append(ref["geo"])
        out["geofea2"].append(obs["geo"])
        out["K"].append(sc.K_crop.astype(np.float32))
        out["T_init"].append(sc.T_init.astype(np.float32))
        out["T_gt"].append(sc.T_gt.astype(np.float32))
        out["diameter"].append(np.float32(sc.diameter))
        if with_images:
            out["syn_img"].append(ref["img"]); out["obs_img"].append(obs["img"])
    res = {k: torch.from_numpy(np.stack(v)) for k, v in out.items() if len(v)}
    return res


def hash_features(shape: Sequence[int], seed: int, scale: float = 1.0) -> torch.Tensor:
    """Reproducible N(0,1)-like float32 tensor (Irwin-Hall) for kernel micro-tests / bench fmaps."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(np.prod(shape))
    return torch.from_numpy((_unit_normal(rng, n) * scale).astype(np.float32).reshape(tuple(shape)))


class AnalyticRenderer:
    """Stand-in for the reference's pytorch3d ``DiffRendererWrapper`` (out of scope)
    with the call contract of ``PoseRefiner.py:134-137,253-254,303-304``.  ``obj_cls`` selects the
    scene: ``scenes[b]`` renders batch element ``b``."""

    def __init__(self, scenes: Sequence[Scene]):
        self.scenes = list(scenes)

    def _TK(self, T, K, b):
        return T[b].detach().cpu().double().numpy(), K[b].detach().cpu().double().numpy()

    def render_pointcloud(self, obj_cls, T=None, K=None, render_image_size=None, **kw):
        return self.render_depth(obj_cls, T=T, K=K, render_image_size=render_image_size)

    def render_depth(self, obj_cls, T=None, K=None, render_image_size=None, near=0.1, far=6, **kw):
        h, w = render_image_size
        out = []
        for b, sc in enumerate(self.scenes):
            Tb, Kb = self._TK(T, K, b)
            Z, _, _ = _raycast(sc, Tb, Kb, h, w)
            out.append(Z.astype(np.float32)[None])
        return torch.from_numpy(np.stack(out)).to(T.device)

    def __call__(self, obj_cls, vert_attribute=None, T=None, K=None, render_image_size=None,
                 near=0.1, far=6, render_tex=False, **kw):
        h, w = render_image_size
        cols, deps = [], []
        for b, sc in enumerate(self.scenes):
            Tb, Kb = self._TK(T, K, b)
            r = render_reference(sc, Tb, Kb, h, w)
            cols.append(np.concatenate([r["img"], r["ctx"], r["geo"]], axis=0))
            d = r["depth"].copy(); d[~r["mask"]] = -1.0      # renderer marks background with -1
            deps.append(d[None])
        return (torch.from_numpy(np.stack(cols)).to(T.device), torch.from_numpy(np.stack(deps)).to(T.device))

"""Generates tests/golden/metrics.npz by EXECUTING the reference's evaluator (utils/eval_metric.py LineMODEvaluator:
add_metric / add2_metric / add5_metric incl. syn=True through thirdparty/nn/nn_utils.py, projection_2d,
cm_degree_5_metric) and utils/geometric.py rotation_angle, unmodified, on CPU.  Run in the build container only:

    python tests/golden/make_golden_metrics.py

What is stubbed (tests/golden/ref_harness.py install_eval_stubs): uninstalled imports the metric methods never touch,
the evaluator's constructor (reads a .ply that is not in the tree), and the compiled CUDA extension thirdparty.nn._ext,
whose single kernel is restated in numpy from nearest_neighborhood.cu:48-80.  The evaluator only keeps booleans
(mean_dist < threshold); the mean distances themselves are tapped by handing the module a numpy proxy whose
mean / rad2deg record what the reference computed (same idea as the cholesky tap of G6).

Cases: 5 point sets (anisotropic ellipsoid surfaces, a box with mirror symmetry, a 7-point degenerate set) x 12 pose
pairs: refined-like small errors, the synthetic scenes' T_init vs T_gt, pairs tuned to sit 0.5 % either side of the
0.1 d / 0.05 d / 0.02 d thresholds, a 180 deg flip (trace angle at the arccos edge) and an exact match.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

import ref_harness as RH  # noqa: E402
from rnnpose_b200 import synthetic as S  # noqa: E402


class NumpyTap:
    """numpy with mean / rad2deg recording their results (everything else is numpy itself)."""

    def __init__(self):
        self.means, self.degs = [], []

    def __getattr__(self, k):
        return getattr(np, k)

    def mean(self, *a, **kw):
        v = np.mean(*a, **kw)
        self.means.append(float(v))
        return v

    def rad2deg(self, x):
        v = np.rad2deg(x)
        self.degs.append(float(v))
        return v


def rot(axis, deg):
    a = np.asarray(axis, np.float64); a = a / np.linalg.norm(a)
    t = np.deg2rad(deg)
    Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(t) * Kx + (1 - np.cos(t)) * (Kx @ Kx)


def point_sets():
    rng = np.random.Generator(np.random.PCG64(77))
    sets = []
    for axes in ((0.035, 0.05, 0.078), (0.08, 0.031, 0.04)):          # anisotropic ellipsoid surfaces
        v = rng.standard_normal((400, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
        sets.append((v * np.asarray(axes), 2 * max(axes)))
    g = np.stack(np.meshgrid(np.linspace(-1, 1, 7), np.linspace(-1, 1, 5), np.linspace(-1, 1, 4), indexing="ij"), -1).reshape(-1, 3)
    sets.append((g * np.array([0.06, 0.04, 0.02]), 0.1497))          # box lattice: mirror symmetric, many equidistant neighbours
    sets.append((rng.standard_normal((300, 3)) * np.array([0.02, 0.05, 0.01]), 0.21))
    sets.append((rng.standard_normal((7, 3)) * 0.03, 0.12))          # fewer points than a warp
    return [(p.astype(np.float32), float(d)) for p, d in sets]


def pose_pairs(pts, diam):
    rng = np.random.Generator(np.random.PCG64(5))
    pairs = []
    Rg = rot(rng.standard_normal(3), 63.0); tg = np.array([0.02, -0.03, 0.9])

    def P(R, t):
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
        return T[:3].astype(np.float32)

    gt = P(Rg, tg)
    pairs.append((gt.copy(), gt))                                                        # exact
    pairs.append((P(rot([1, 2, 3], 0.4) @ Rg, tg + [0.0004, -0.0002, 0.001]), gt))     # refined-like
    pairs.append((P(rot([0, 1, 0], 4.0) @ Rg, tg + [0.002, 0.001, 0.006]), gt))
    pairs.append((P(rot([1, 0, 0], 180.0) @ Rg, tg), gt))                               # flip: trace = -1 edge
    pairs.append((P(rot([3, -1, 2], 25.0) @ Rg, tg + [0.01, -0.01, 0.05]), gt))        # init-like
    for sc in (0, 3):                                                                    # the synthetic scenes' own poses
        s = S.make_scene(sc, 128, 160)
        pairs.append((s.T_init[:3].astype(np.float32), s.T_gt[:3].astype(np.float32)))
    # pure translations: ADD = |dt| exactly (up to rounding), placed 0.5 % either side of each threshold
    for frac in (0.1, 0.05, 0.02):
        for side in (0.995, 1.005):
            dt = np.array([0.6, 0.0, 0.8]) * (frac * diam * side)
            pairs.append((P(Rg, tg + dt), gt))
    # 5 cm / 5 deg edges
    pairs.append((P(rot([0, 0, 1], 4.97) @ Rg, tg + [0.0, 0.0, 0.0497]), gt))
    pairs.append((P(rot([0, 0, 1], 5.03) @ Rg, tg + [0.0, 0.0, 0.0497]), gt))
    return pairs


def main():
    RH.install_eval_stubs()
    from data.linemod import linemod_config
    from utils.geometric import rotation_angle
    K = linemod_config.linemod_K
    out = {"K": K.astype(np.float64)}
    rows = []
    sets = point_sets()
    for si, (pts, diam) in enumerate(sets):
        ev, EM = RH.reference_evaluator(pts, diam)
        out[f"pts{si}"] = pts
        for pi, (pp, pg) in enumerate(pose_pairs(pts, diam)):
            tap = NumpyTap()
            EM.np = tap
            try:
                for lst in (ev.add, ev.add2, ev.add5, ev.proj2d, ev.cmd5):
                    del lst[:]
                ev.add_metric(pp, pg); ev.add2_metric(pp, pg); ev.add5_metric(pp, pg)
                ev.add_metric(pp, pg, syn=True); ev.add2_metric(pp, pg, syn=True); ev.add5_metric(pp, pg, syn=True)
                ev.projection_2d(pp, pg, K=K)
                ev.cm_degree_5_metric(pp, pg)
            finally:
                EM.np = np
            m = tap.means
            assert len(m) == 7 and m[0] == m[1] == m[2] and m[3] == m[4] == m[5] and len(tap.degs) == 1
            ang = float(rotation_angle(pg[:3, :3], pp[:3, :3]))                       # eval_metric.py:326
            trans = float(np.linalg.norm(pp[:3, 3:] - pg[:3, -1:]))                   # eval_metric.py:327
            rows.append([si, pi, diam, m[0], m[3], m[6], tap.degs[0], np.rad2deg(ang), trans,
                         float(ev.add[0]), float(ev.add[1]), float(ev.add2[0]), float(ev.add2[1]), float(ev.add5[0]),
                         float(ev.add5[1]), float(ev.proj2d[0]), float(ev.cmd5[0])])
            out[f"pose_pred_{si}_{pi}"] = pp
            out[f"pose_gt_{si}_{pi}"] = pg
    # ---- the refine goldens' own final poses through the evaluator: ADD / ADD-S of the executed reference's Ti_pred against
    # the scene's ground truth (what "ADD(-S) matching to 4 decimals" is measured against on the GPU)
    for name in ("refine_cfg0_240x320_1x1", "refine_128x160_4x3", "refine_240x320_4x3", "refine_occl_128x160_8x3"):
        g = np.load(os.path.join(HERE, name + ".npz"))
        H, W, n_iters, n_lm, seed, occl = [int(v) for v in g["meta"]]
        vals = []
        for k, idx in enumerate(int(i) for i in g["idxs"]):
            sc = S.make_scene(idx, H, W, seed, bool(occl))
            ev, EM = RH.reference_evaluator(S.model_points(sc), sc.diameter)
            pp = g["Ti_pred"][k][:3].astype(np.float32); pg = sc.T_gt[:3].astype(np.float32)
            tap = NumpyTap(); EM.np = tap
            try:
                ev.add_metric(pp, pg); ev.add_metric(pp, pg, syn=True); ev.projection_2d(pp, pg, K=K)
            finally:
                EM.np = np
            vals.append([sc.diameter, tap.means[0], tap.means[1], tap.means[2], float(ev.add[0]), float(ev.add[1]), float(ev.proj2d[0])])
        out["refine__" + name] = np.asarray(vals, np.float64)      # diameter, ADD, ADD-S, proj2d, flags
    # columns: set, pair, diameter, ADD, ADD-S, proj2d px, trace angle deg, chordal angle deg, trans,
    #          add<.1d, adds<.1d, add<.02d, adds<.02d, add<.05d, adds<.05d, proj2d<5, cm5
    out["rows"] = np.asarray(rows, np.float64)
    path = os.path.join(HERE, "metrics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(rows), "cases")
    r = out["rows"]
    print("ADD/d range", (r[:, 3] / r[:, 2]).min(), (r[:, 3] / r[:, 2]).max(), " flags:", r[:, 9:].sum(0))


if __name__ == "__main__":
    main()

"""G10: drop-in fixture.  Executes the UNMODIFIED reference PoseRefiner.forward INCLUDING its own zoom-crop
(gen_zoom_crop_grids / get_affine_transformation with cv2, reference model/PoseRefiner.py:145-213) and several
render iterations, with the analytic stub renderer.  Stores what the out-of-scope encoder produced (replayed on the
GPU box), the crop parameters the reference computed, and the final pose.

Run in the build container only:  python tests/golden/make_golden_dropin.py
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_harness as RH  # noqa: E402

RH.install_stubs()
from rnnpose_b200 import synthetic as S  # noqa: E402

IMG_HW = (240, 320)       # "full" image / render size
CROP_HW = (128, 160)      # zoom crop fed to the inner loop


def dropin_scene(idx):
    """Scene whose K_crop (object filling half of the image height) plays the role of the full-image intrinsics."""
    return S.make_scene(idx, IMG_HW[0], IMG_HW[1], seed=4321, fill=0.5)


def dropin_inputs(sc):
    obs = S.render_observed(sc, K=sc.K_crop, H=IMG_HW[0], W=IMG_HW[1])
    image = torch.from_numpy(obs["img"])[None]
    geo2 = torch.from_numpy(obs["geo"])[None]
    K = torch.from_numpy(sc.K_crop.astype(np.float32))[None]
    T0 = torch.from_numpy(sc.T_init.astype(np.float32))[None, None]
    Tgt = torch.from_numpy(sc.T_gt.astype(np.float32))[None, None]
    return image, geo2, K, T0, Tgt


def main():
    import model.PoseRefiner as PR
    from config.default import get_cfg
    from geometry.transformation import SE3Sequence
    n_render, n_iters, n_lm = 2, 2, 1
    rec = {k: [] for k in ("fmaps", "theta", "K_crop", "Ti_pred", "Tij")}
    idxs = [0, 1]
    for idx in idxs:
        sc = dropin_scene(idx)
        ren = S.AnalyticRenderer([sc])
        get_cfg().merge({"zoom_crop_size": list(CROP_HW), "render_image_size": list(IMG_HW)}, "BASIC")
        net = PR.PoseRefiner(RH.motion_cfg(iter_count=n_iters, optim_iter_count=n_lm, render_iter_count=n_render), renderer=ren)
        net.eval()
        fm = []
        orig_enc = net.image_fea_enc.forward

        def enc(a, b):
            f1, f2 = orig_enc(a, b)
            fm.append(torch.stack([f1[0].float(), f2[0].float()]))
            return f1, f2
        net.image_fea_enc.forward = enc
        crops = []
        orig_crop = net.gen_zoom_crop_grids

        def crop(fg_mask, K, T, output_size, model_center=None, margin_ratio=0.4):
            ys, xs = np.nonzero(fg_mask[0, 0].numpy())
            grids, Kc = orig_crop(fg_mask, K, T, output_size=output_size, model_center=model_center, margin_ratio=margin_ratio)
            # recover theta from the grid corners (affine_grid of an axis-aligned box)
            crops.append((grids.clone(), Kc.clone()))
            return grids, Kc
        net.gen_zoom_crop_grids = crop
        image, geo2, K, T0, Tgt = dropin_inputs(sc)
        with torch.no_grad():
            out = net(image, SE3Sequence(matrix=T0), K, fea_3d=torch.zeros(1, 4, 256), Tj_gt=SE3Sequence(matrix=Tgt),
                      obj_cls=None, geofea_3d=torch.zeros(1, 4, 32), geofea_2d=geo2)
        rec["fmaps"].append(torch.stack(fm))                                  # [n_render, 2, 256, h, w]
        rec["theta"].append(torch.stack([g[0][0, [0, 0, -1], [0, -1, 0]] for g in crops]))   # 3 corner samples of the grid
        rec["K_crop"].append(torch.stack([g[1][0] for g in crops]))
        rec["Ti_pred"].append(out["Ti_pred"].G[0, 0]); rec["Tij"].append(out["Tij"].G[0, 0])
        print("dropin", idx, "Ti_pred moved by", (out["Ti_pred"].G[0, 0] - T0[0, 0]).abs().max().item())
    path = os.path.join(HERE, "dropin_2x2x1.npz")
    np.savez_compressed(path, meta=np.array([n_render, n_iters, n_lm]), idxs=np.array(idxs),
                        **{k: torch.stack(v).numpy() for k, v in rec.items()})
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()

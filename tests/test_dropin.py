"""Drop-in boundary: the PoseRefiner mirror against the executed reference INCLUDING the
reference's own zoom-crop and two render iterations (fixture tests/golden/dropin_2x2x1.npz)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rnnpose_b200 import synthetic as S
from rnnpose_b200.refiner import PoseRefiner, zoom_crop_params
from rnnpose_b200.se3 import SE3Sequence
from tests.util import golden, load_update_weights

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
IMG_HW, CROP_HW = (240, 320), (128, 160)


def scene(idx):
    return S.make_scene(idx, IMG_HW[0], IMG_HW[1], seed=4321, fill=0.5)


def inputs(sc):
    obs = S.render_observed(sc, K=sc.K_crop, H=IMG_HW[0], W=IMG_HW[1])
    return (torch.from_numpy(obs["img"])[None], torch.from_numpy(obs["geo"])[None],
            torch.from_numpy(sc.K_crop.astype(np.float32))[None], torch.from_numpy(sc.T_init.astype(np.float32))[None, None],
            torch.from_numpy(sc.T_gt.astype(np.float32))[None, None])


def test_zoom_crop_matches_reference_cv2_path():
    """Device-side crop geometry == reference get_affine_transformation (numpy + cv2.getAffineTransform)."""
    g = golden("dropin_2x2x1.npz")
    for k, idx in enumerate(int(i) for i in g["idxs"]):
        sc = scene(idx)
        image, geo2, K, T0, _ = inputs(sc)
        ren = S.AnalyticRenderer([sc])
        pc = ren.render_pointcloud(None, T=T0[:, 0], K=K, render_image_size=IMG_HW)
        theta, K_crop = zoom_crop_params(pc > 0, K, T0[:, 0], CROP_HW)
        torch.testing.assert_close(K_crop[0], torch.from_numpy(g["K_crop"][k, 0]), rtol=1e-5, atol=1e-3)
        grid = F.affine_grid(theta, torch.Size([1, 1, *CROP_HW]))
        corners = grid[0, [0, 0, -1], [0, -1, 0]]
        torch.testing.assert_close(corners, torch.from_numpy(g["theta"][k, 0]), rtol=1e-5, atol=1e-5)


def test_state_dict_layout_matches_reference_checkpoint_keys():
    net = PoseRefiner({"FLOW_NET": "raft"}, renderer=None, image_fea_enc=torch.nn.Identity(),
                      render_image_size=IMG_HW, zoom_crop_size=CROP_HW)
    sd = net.state_dict()
    assert sd["sigma.0"].shape == (1,)
    ref = load_update_weights()
    for k, v in ref.items():
        assert sd["cf_net.update_block." + k].shape == v.shape, k
    assert len([k for k in sd if k.startswith("cf_net.update_block.")]) == 30
    missing, unexpected = net.cf_net.load_state_dict({"update_block." + k: v for k, v in ref.items()}, strict=True), None


class ReplayEncoder(torch.nn.Module):
    """Stands in for the (out-of-scope) RAFT BasicEncoder: replays the feature maps the reference computed."""

    def __init__(self, fmaps):
        super().__init__()
        self.fmaps, self.i = fmaps, 0

    def forward(self, a, b):
        f = self.fmaps[self.i]; self.i += 1
        return f[0:1].to(a.device), f[1:2].to(a.device)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 1], ids=["fp32", "tcgen05"])
def test_pose_refiner_dropin_matches_reference(flags):
    g = golden("dropin_2x2x1.npz")
    n_render, n_iters, n_lm = [int(v) for v in g["meta"]]
    dev = torch.device("cuda:0")
    for k, idx in enumerate(int(i) for i in g["idxs"]):
        sc = scene(idx)
        image, geo2, K, T0, Tgt = inputs(sc)
        cfg = {"FLOW_NET": "raft", "IS_CALIBRATED": True, "ITER_COUNT": n_iters, "RENDER_ITER_COUNT": n_render,
               "OPTIM_ITER_COUNT": n_lm}
        net = PoseRefiner(cfg, renderer=S.AnalyticRenderer([sc]), image_fea_enc=ReplayEncoder(torch.from_numpy(g["fmaps"][k])),
                          render_image_size=IMG_HW, zoom_crop_size=CROP_HW)
        net.cf_net.load_state_dict({"update_block." + kk: v for kk, v in load_update_weights().items()}, strict=True)
        net = net.to(dev)
        net.flags = flags
        out = net(image.to(dev), SE3Sequence(matrix=T0.to(dev)), K.to(dev), fea_3d=torch.zeros(1, 4, 256, device=dev),
                  Tj_gt=SE3Sequence(matrix=Tgt.to(dev)), obj_cls=None, geofea_3d=torch.zeros(1, 4, 32, device=dev),
                  geofea_2d=geo2.to(dev))
        err = (out["Ti_pred"].G[0, 0].cpu() - torch.from_numpy(g["Ti_pred"][k])).abs().max().item()
        print(f"[dropin] scene {idx} flags={flags}: max |dSE3| vs executed reference = {err:.3e}")
        assert err < 1e-4
        assert out["weight"].shape == (1, 1, 1, *CROP_HW) and out["flow"][0].shape == (1, 2, *CROP_HW)
        assert len(out["syn_depth"]) == n_render * n_iters and len(out["Tij_gt"]) == n_render * n_iters


@pytest.mark.gpu
def test_zoom_crop_kernel_matches_reference_cv2_path():
    """b200pose_zoom_crop (bounding box + crop intrinsics + both resampled crops on device) against the crop parameters the
    executed reference computed with numpy + cv2 (dropin_2x2x1.npz) and against F.affine_grid + F.grid_sample on the same
    theta; NCHW and channels-last descriptor outputs; natively batched; an empty foreground gives the reference's zero box."""
    from rnnpose_b200 import ops
    g = golden("dropin_2x2x1.npz")
    dev = torch.device("cuda:0")
    idxs = [int(i) for i in g["idxs"]]
    scs = [scene(i) for i in idxs]
    ins = [inputs(sc) for sc in scs]
    image = torch.cat([i[0] for i in ins]).to(dev); geo2 = torch.cat([i[1] for i in ins]).to(dev)
    K = torch.cat([i[2] for i in ins]).to(dev); T0 = torch.cat([i[3][:, 0] for i in ins]).to(dev)
    pcs = torch.cat([S.AnalyticRenderer([sc]).render_pointcloud(None, T=i[3][:, 0], K=i[2], render_image_size=IMG_HW) for sc, i in zip(scs, ins)])
    pc = pcs[:, 0].contiguous().to(dev)
    out = ops.zoom_crop(pc, K, T0, image, geo2, CROP_HW, want_theta=True)
    for k in range(len(idxs)):
        torch.testing.assert_close(out["K_crop"][k].cpu(), torch.from_numpy(g["K_crop"][k, 0]), rtol=1e-5, atol=1e-3)
    grid = F.affine_grid(out["theta"], torch.Size([len(idxs), 1, *CROP_HW]), align_corners=False)
    corners = grid[:, [0, 0, -1], [0, -1, 0]].cpu()
    torch.testing.assert_close(corners, torch.from_numpy(g["theta"][:, 0]), rtol=1e-5, atol=1e-5)
    # geometry also equals the torch restatement used as checker
    theta_t, Kc_t = zoom_crop_params(pc[:, None] > 0, K, T0, CROP_HW)
    torch.testing.assert_close(out["theta"], theta_t, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["K_crop"], Kc_t, rtol=1e-5, atol=1e-3)
    # the resampled crops
    ref_img = F.grid_sample(image, grid, align_corners=False); ref_geo = F.grid_sample(geo2, grid, align_corners=False)
    torch.testing.assert_close(out["image_crop"], ref_img, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(out["geofea_crop"], ref_geo, rtol=1e-4, atol=2e-5)
    cl = ops.zoom_crop(pc, K, T0, image, geo2, CROP_HW, channels_last=True)
    B = len(idxs)
    assert cl["geofea_crop"].shape == (B, CROP_HW[0] * CROP_HW[1], 32)
    assert torch.equal(cl["geofea_crop"].view(B, *CROP_HW, 32).permute(0, 3, 1, 2), out["geofea_crop"])
    assert torch.equal(cl["image_crop"], out["image_crop"]) and torch.equal(cl["K_crop"], out["K_crop"])
    # an output size that is not a multiple of the 64-pixel block, and an empty foreground in sample 1
    pc2 = pc.clone(); pc2[1] = 0
    odd = ops.zoom_crop(pc2, K, T0, image, geo2, (50, 70), want_theta=True, channels_last=True)
    th2, Kc2 = zoom_crop_params(pc2[:, None] > 0, K, T0, (50, 70))
    torch.testing.assert_close(odd["theta"], th2, rtol=1e-5, atol=1e-6)
    assert torch.isfinite(odd["K_crop"][0]).all()
    grid2 = F.affine_grid(odd["theta"][:1], torch.Size([1, 1, 50, 70]), align_corners=False)
    torch.testing.assert_close(odd["geofea_crop"][0].view(50, 70, 32).permute(2, 0, 1), F.grid_sample(geo2[:1], grid2, align_corners=False)[0],
                               rtol=1e-4, atol=2e-5)


@pytest.mark.gpu
def test_pose_refiner_standalone_with_library_encoder():
    """The drop-in with its default encoder (the library's RAFT encoder kernels; nothing from the reference tree): two render
    iterations end to end on the GPU, checked against the CPU oracles chained the same way (encoder oracle -> refine oracle on
    the crops the zoom-crop kernel produced)."""
    from oracle import encoder_oracle as E, refine_oracle as O
    from rnnpose_b200 import ops
    from rnnpose_b200.assets import load_encoder_weights
    dev = torch.device("cuda:0")
    sc = scene(0)
    image, geo2, K, T0, Tgt = inputs(sc)
    cfg = {"FLOW_NET": "raft", "IS_CALIBRATED": True, "ITER_COUNT": 2, "RENDER_ITER_COUNT": 1, "OPTIM_ITER_COUNT": 2}
    net = PoseRefiner(cfg, renderer=S.AnalyticRenderer([sc]), render_image_size=IMG_HW, zoom_crop_size=CROP_HW)
    sd = net.state_dict()
    assert any(k.startswith("image_fea_enc.fnet.conv1") for k in sd) and len([k for k in sd if k.startswith("image_fea_enc.")]) == 32
    net.cf_net.load_state_dict({"update_block." + kk: v for kk, v in load_update_weights().items()}, strict=True)
    net = net.to(dev)
    # images in [0,255] as a well-conditioned encoder input (the reference pipeline feeds [0,1])
    img255 = (image * 255.0).to(dev)
    out = net(img255, SE3Sequence(matrix=T0.to(dev)), K.to(dev), fea_3d=torch.zeros(1, 4, 256, device=dev),
              Tj_gt=SE3Sequence(matrix=Tgt.to(dev)), obj_cls=None, geofea_3d=torch.zeros(1, 4, 32, device=dev), geofea_2d=geo2.to(dev))
    G = out["Ti_pred"].G[0, 0].cpu()
    assert torch.isfinite(G).all() and out["flow"][0].shape == (1, 2, *CROP_HW)
    # CPU chain on the same crops
    ren = S.AnalyticRenderer([sc])
    pc = ren.render_pointcloud(None, T=T0[:, 0], K=K, render_image_size=IMG_HW)
    zc = ops.zoom_crop(pc[:, 0].contiguous().to(dev), K.to(dev), T0[:, 0].contiguous().to(dev), img255, geo2.to(dev), CROP_HW)
    Kc = zc["K_crop"].cpu()
    color, depth = ren(None, torch.zeros(1, 4, 288), T=T0[:, 0], K=Kc, render_image_size=CROP_HW, render_tex=True)
    syn_img, cfea, geo1 = torch.split(color, [3, 256, 32], dim=1)
    syn_depth = ren.render_depth(None, T=T0[:, 0], K=Kc, render_image_size=CROP_HW)
    with torch.no_grad():
        f1, f2 = E.image_encoder(load_encoder_weights(), syn_img, zc["image_crop"].cpu())
        ref = O.refine_inner_loop(load_update_weights(), f1, f2, cfea * 0.1, geo1, zc["geofea_crop"].cpu(), syn_depth, Kc,
                                  torch.eye(4)[None], n_iters=2, n_lm=2)
    Ti_ref = torch.matmul(ref["G"][0], T0[0, 0])
    err = (G - Ti_ref).abs().max().item()
    print(f"[dropin standalone] max |dSE3| vs CPU oracle chain = {err:.3e}")
    assert err < 1e-3        # syn_img is in [0,1] here (renderer output): the encoder is ill-conditioned on it, see test_gpu_encoder.py

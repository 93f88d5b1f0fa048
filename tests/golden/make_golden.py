"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (/root/reference) on CPU fp32.

Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin ``oracle/refine_oracle.py`` (tests/test_oracle_golden.py) and, on the GPU box, the
CUDA path (tests/test_gpu_*.py).  Large inputs are not stored: they are regenerated bit-exactly from
seeds by ``rnnpose_b200.synthetic`` (exactly-rounded ops only); only small inputs, the encoder
feature maps (the encoder is out of scope and stays PyTorch) and the reference OUTPUTS are stored.

Vectors:
  G1 corr_lookup.npz   CorrBlock pyramid + lookup                 thirdparty/raft/corr.py:12-67
  G2 update_block.npz  BasicUpdateBlock.forward, shipped weights   thirdparty/raft/update.py:164-188
  G3 cfnet_seq.npz     GRU_CFUpdator.forward x3 (state carry)      model/CFNet.py:109-173
  G4 upsample.npz      GRU_CFUpdator.upsample_flow                 model/CFNet.py:95-106
  G5 weight.npz        correspondence weight                       model/PoseRefiner.py:342-345
  G6 lm.npz            SE3Sequence.reprojction_optim (H,b,delta,G) geometry/transformation.py:265-316
  G7 expm.npz          _se3_matrix_expm both branches              geometry/se3.py:228-281
  G8 refine_*.npz      full PoseRefiner.forward, stub renderer     model/PoseRefiner.py:221-376
  G9 cholesky.npz      geometry/cholesky.py __test__ system + random 6x6 systems
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_harness as RH  # noqa: E402

RH.install_stubs()
from rnnpose_b200 import synthetic as S  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f"wrote {name}: {os.path.getsize(path) / 1e6:.2f} MB  keys={list(out)}")


def shipped_update_weights():
    sd = torch.load(os.path.join(RH.REF, "weights", "gru_update.pth"), map_location="cpu")
    return {k[len("update_block."):]: v.float() for k, v in sd.items()}


# ------------------------------------------------------------------------------------------- G1
def g1_corr_lookup():
    from thirdparty.raft.corr import CorrBlock
    B, D, h, w = 2, 32, 17, 22
    f1 = S.hash_features((B, D, h, w), 101)
    f2 = S.hash_features((B, D, h, w), 102)
    coords = torch.stack(torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")[::-1], 0)
    coords = coords[None].repeat(B, 1, 1, 1) + S.hash_features((B, 2, h, w), 103, 3.0)
    coords[0, :, 0, 0] = torch.tensor([-7.3, 2.5])          # far out of range
    coords[0, :, 0, 1] = torch.tensor([w + 5.2, h + 9.9])
    coords[0, :, 0, 2] = torch.tensor([3.0, 4.0])            # integer coordinates
    coords[1, :, 1, 1] = torch.tensor([w - 1.0, h - 1.0])    # exactly on the last pixel
    cb = CorrBlock(f1, f2, num_levels=4, radius=4)
    out = cb(coords)
    save("corr_lookup.npz", coords=coords, out=out,
         pyr0=cb.corr_pyramid[0].reshape(B, h * w, h, w), pyr1=cb.corr_pyramid[1].reshape(B, h * w, 8, 11),
         pyr3=cb.corr_pyramid[3].reshape(B, h * w, 2, 2), meta=np.array([B, D, h, w, 101, 102]))


# ------------------------------------------------------------------------------------------- G2
def g2_update_block():
    from easydict import EasyDict
    from thirdparty.raft.update import BasicUpdateBlock
    args = EasyDict(corr_levels=4, corr_radius=4)
    ub = BasicUpdateBlock(args, hidden_dim=128)
    ub.load_state_dict(shipped_update_weights(), strict=True)
    ub.eval()
    B, h, w = 2, 9, 12
    net = torch.tanh(S.hash_features((B, 128, h, w), 201))
    inp = torch.relu(S.hash_features((B, 128, h, w), 202))
    corr = S.hash_features((B, 324, h, w), 203, 2.0)
    flow = S.hash_features((B, 2, h, w), 204, 4.0)
    with torch.no_grad():
        net2, mask, dflow = ub(net, inp, corr, flow)
    save("update_block.npz", net_out=net2, mask=mask, dflow=dflow, meta=np.array([B, h, w, 201, 202, 203, 204]))


# ------------------------------------------------------------------------------------------- G3/G4
def g3_cfnet_seq():
    from model.CFNet import GRU_CFUpdator
    cf = GRU_CFUpdator(RH.motion_cfg().raft)
    cf.eval()
    B, H, W = 1, 128, 160
    h, w = H // 8, W // 8
    f1 = S.hash_features((B, 256, h, w), 301); f2 = S.hash_features((B, 256, h, w), 302)
    ctx = S.hash_features((B, 256, H, W), 303, 0.1)
    flows = []
    lows = []
    with torch.no_grad():
        for it in range(3):
            fi = S.hash_features((B, 2, H, W), 310 + it, 6.0)
            out = cf(f1, f2, flow_init=fi.clone(), context_fea=ctx, update_corr_fn=(it == 0))
            flows.append(out[-1]); lows.append(cf.net.clone())
    save("cfnet_seq.npz", flow_up=torch.stack(flows), net=torch.stack(lows), meta=np.array([B, H, W, 301, 302, 303, 310]))
    # G4: upsample_flow in isolation
    fl = S.hash_features((2, 2, 7, 9), 401, 3.0); mk = S.hash_features((2, 576, 7, 9), 402, 2.0)
    save("upsample.npz", out=cf.upsample_flow(fl, mk), meta=np.array([2, 7, 9, 401, 402]))


# ------------------------------------------------------------------------------------------- G5
def g5_weight():
    from geometry.projective_ops import normalize_coords_grid
    B, C, H, W = 2, 32, 40, 56
    g1 = S.hash_features((B, C, H, W), 501); g1 = g1 / g1.norm(dim=1, keepdim=True)
    g2 = S.hash_features((B, C, H, W), 502); g2 = g2 / g2.norm(dim=1, keepdim=True)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    tgt = torch.stack([xx, yy], -1)[None, None].repeat(B, 1, 1, 1, 1) + S.hash_features((B, 1, H, W, 2), 503, 5.0)
    tgt[0, 0, 0, 0] = torch.tensor([-3.0, -2.0]); tgt[0, 0, 0, 1] = torch.tensor([W + 1.5, 3.0])
    depth = (S.hash_features((B, 1, H, W), 504) > -0.3).float() * 0.9
    sigma = 0.7
    warp = F.grid_sample(g2, normalize_coords_grid(tgt).squeeze(1))                 # PoseRefiner.py:343
    cw = torch.sum(g1 * warp, dim=1, keepdim=True).permute(0, 2, 3, 1)[:, None]     # :344
    cw = torch.exp(-torch.abs(1 - cw) / sigma) * (depth > 0)[..., None].float()      # :345
    save("weight.npz", weight=cw[:, 0, :, :, 0], meta=np.array([B, C, H, W, 501, 502, 503, 504]), sigma=np.float32(sigma))


# ------------------------------------------------------------------------------------------- G6
def lm_inputs(B, H, W, seed):
    """Shared by make_golden and the tests: depth/target/weight/K/G for the LM step."""
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    depth = 0.8 + 0.2 * torch.sigmoid(S.hash_features((B, 1, H, W), seed))
    depth = depth * (S.hash_features((B, 1, H, W), seed + 1) > -0.8).float()        # background zeros
    depth[:, :, 0, :4] = 0.05                                                        # X0z <= 0.1 -> invalid
    depth = depth + 1e-5
    K = torch.tensor([[300.0, 0, W / 2 - 0.3], [0, 305.0, H / 2 + 0.2], [0, 0, 1]])[None].repeat(B, 1, 1)
    K[1:, 0, 0] *= 1.1
    target = torch.stack([xx, yy], -1)[None, None].repeat(B, 1, 1, 1, 1) + S.hash_features((B, 1, H, W, 2), seed + 2, 2.0)
    weight = torch.rand(B, 1, H, W, 1, generator=torch.Generator().manual_seed(seed + 3))
    weight[:, :, 1, :5] = 0.0
    from oracle.refine_oracle import se3_exp
    xi = torch.tensor([[0.01, -0.02, 0.015, 0.02, -0.01, 0.03], [-0.03, 0.01, 0.0, -0.015, 0.025, -0.02]])[:B]
    G = se3_exp(xi)[:, None]
    return depth, target, weight, K, G


def g6_lm():
    import geometry.transformation as GT
    from geometry.transformation import SE3Sequence
    B, H, W = 2, 48, 64
    res = {}
    taps = []
    orig = GT.cholesky_solve

    def tap(Hm, b):
        taps.append((Hm.clone(), b.clone()))
        return orig(Hm, b)
    GT.cholesky_solve = tap
    for n in (1, 3):
        taps.clear()
        depth, target, weight, K, G = lm_inputs(B, H, W, 600)
        T = SE3Sequence(matrix=G.clone())
        out = T.reprojction_optim(target, weight, depth, K, num_iters=n)
        res[f"G_out_{n}"] = out.G[:, 0]
        res[f"Hd_{n}"] = torch.stack([t[0][:, 0] for t in taps])       # DAMPED H as passed to the solver
        res[f"b_{n}"] = torch.stack([t[1][:, 0] for t in taps])
    # a point behind the camera / at the Jacobian cut-off: pose that pushes some Z below 0.02
    depth, target, weight, K, G = lm_inputs(B, H, W, 600)
    G2 = G.clone(); G2[0, 0, 2, 3] -= 0.9
    T = SE3Sequence(matrix=G2.clone())
    taps.clear()
    out = T.reprojction_optim(target, weight, depth, K, num_iters=2)
    res["G_in_near"] = G2[:, 0]; res["G_out_near"] = out.G[:, 0]
    res["Hd_near"] = torch.stack([t[0][:, 0] for t in taps]); res["b_near"] = torch.stack([t[1][:, 0] for t in taps])
    # NaN injection: a NaN weight poisons H, b -> solver output NaN -> zero update (cholesky.py:42-45)
    depth, target, weight, K, G = lm_inputs(B, H, W, 600)
    weight[0, 0, 5, 5, 0] = float("nan")
    T = SE3Sequence(matrix=G.clone())
    try:
        out = T.reprojction_optim(target, weight, depth, K, num_iters=1)
        res["G_out_nan"] = out.G[:, 0]; res["nan_raises"] = np.array(0)
    except Exception as e:  # torch.cholesky raises on a NaN matrix in recent torch
        print("NaN case raised:", type(e).__name__, str(e)[:80])
        res["nan_raises"] = np.array(1)
    GT.cholesky_solve = orig
    depth, target, weight, K, G = lm_inputs(B, H, W, 600)
    res["G_in"] = G[:, 0]
    # inputs are stored (sigmoid / rand / sin / cos are not bit-reproducible across machines)
    res.update(depth=depth[:, 0], target=target[:, 0], weight=weight[:, 0, :, :, 0], K=K)
    save("lm.npz", meta=np.array([B, H, W, 600]), **res)


# ------------------------------------------------------------------------------------------- G7 / G9
def g7_expm_cholesky():
    from geometry.se3 import _se3_matrix_expm
    from geometry import cholesky
    xi = torch.tensor([[0.1, -0.2, 0.3, 0.4, -0.5, 0.6],
                       [1.0, 1.0, -1.0, 1.0, -1.0, 1.0],
                       [0.3, 0.2, 0.1, 3e-5, -2e-5, 5e-5],       # theta < 1e-4 -> Taylor branch
                       [0.3, 0.2, 0.1, 0.0, 0.0, 0.0],
                       [-0.7, 0.0, 0.2, 6e-5, 6e-5, 5.2e-5],     # theta just above/below threshold
                       [0.0, 0.0, 0.0, 1e-3, 0.0, 0.0]], dtype=torch.float32)
    save("expm.npz", xi=xi, G=_se3_matrix_expm(xi))
    np.random.seed(0)
    M = np.random.uniform(size=(3, 3))
    H3 = torch.tensor(M @ M.T); b3 = torch.tensor(np.random.uniform(size=(3,)))
    x3 = cholesky.solve(H3, b3)                                    # cholesky.py:54-67 (__test__ system)
    g = torch.Generator().manual_seed(7)
    A = torch.randn(8, 6, 6, generator=g, dtype=torch.float64)
    H6 = A @ A.transpose(1, 2) + 0.5 * torch.eye(6, dtype=torch.float64)
    b6 = torch.randn(8, 6, generator=g, dtype=torch.float64) * torch.tensor([1, 10, 0.1, 5, 1, 30.0], dtype=torch.float64)
    x6 = cholesky.solve(H6, b6)
    save("cholesky.npz", H3=H3, b3=b3, x3=x3, H6=H6, b6=b6, x6=x6)


# ------------------------------------------------------------------------------------------- G8
def g8_refine(name, idxs, H, W, n_iters, n_lm, seed=1234, occlude=False):
    """Full reference PoseRefiner.forward (one render iteration, B=1 per call -- the reference
    cannot batch) with the analytic stub renderer.  The zoom-crop is the identity because the
    synthetic K_crop already is the zoomed intrinsics (the crop is SURVEY 8(f)-1, not hot path)."""
    import model.PoseRefiner as PR
    from geometry.transformation import SE3Sequence

    def ident_grids(self, fg_mask, K, T, output_size, model_center=None, margin_ratio=0.4):
        theta = torch.eye(2, 3)[None].repeat(fg_mask.shape[0], 1, 1)
        return F.affine_grid(theta, torch.Size(output_size)), K
    PR.PoseRefiner.gen_zoom_crop_grids = ident_grids

    rec = {k: [] for k in ("fmap1", "fmap2", "Ti_pred", "Tij", "G0", "flow_last", "flow_first", "weight")}
    for idx in idxs:
        sc = S.make_scene(idx, H, W, seed, occlude)
        ren = S.AnalyticRenderer([sc])
        net = RH.build_reference_refiner(ren, H, W, iter_count=n_iters, optim_iter_count=n_lm)
        cap = {}
        orig_fwd = net.cf_net.forward

        def fwd(fmap1, fmap2, *a, **kw):
            if kw.get("update_corr_fn", True):
                cap["fmap1"] = fmap1.clone(); cap["fmap2"] = fmap2.clone(); cap["ctx"] = kw["context_fea"].clone()
            return orig_fwd(fmap1, fmap2, *a, **kw)
        net.cf_net.forward = fwd
        obs = S.render_observed(sc)
        image = torch.from_numpy(obs["img"])[None]; geo2 = torch.from_numpy(obs["geo"])[None]
        K = torch.from_numpy(sc.K_crop.astype(np.float32))[None]
        Ts = SE3Sequence(matrix=torch.from_numpy(sc.T_init.astype(np.float32))[None, None])
        Tgt = SE3Sequence(matrix=torch.from_numpy(sc.T_gt.astype(np.float32))[None, None])
        with torch.no_grad():
            out = net(image, Ts, K, fea_3d=torch.zeros(1, 4, 256), Tj_gt=Tgt, obj_cls=None,
                      geofea_3d=torch.zeros(1, 4, 32), geofea_2d=geo2)
        # input equivalence: what the reference fed its inner loop == rnnpose_b200.synthetic.make_batch
        mb = S.make_batch([idx], H, W, seed, occlude)
        assert torch.allclose(cap["ctx"], mb["context"], atol=1e-7), "context mismatch"
        assert torch.equal(out["syn_depth"][0], mb["depth"]), "depth mismatch"
        G0 = (Ts * Ts.inv()).G                             # legacy "identity", PoseRefiner.py:243-244
        rec["fmap1"].append(cap["fmap1"][0]); rec["fmap2"].append(cap["fmap2"][0])
        rec["Ti_pred"].append(out["Ti_pred"].G[0, 0]); rec["Tij"].append(out["Tij"].G[0, 0]); rec["G0"].append(G0[0, 0])
        fl = net.flow_history
        rec["flow_first"].append(out["flow"][-1][0, :, ::4, ::4])
        rec["flow_last"].append(fl[-1][-1][0, :, ::4, ::4] if len(fl) else out["flow"][-1][0, :, ::4, ::4])
        rec["weight"].append(out["weight"][0, 0, 0, ::4, ::4])
        print(name, idx, "done; |Ti_pred - T_init| max", (out["Ti_pred"].G[0, 0] - Ts.G[0, 0]).abs().max().item())
    save(name, meta=np.array([H, W, n_iters, n_lm, seed, int(occlude)]), idxs=np.array(idxs),
         **{k: torch.stack(v) for k, v in rec.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["g1", "g2", "g3", "g5", "g6", "g7", "g8"]
    with torch.no_grad():
        if "g1" in which: g1_corr_lookup()
        if "g2" in which: g2_update_block()
        if "g3" in which: g3_cfnet_seq()
        if "g5" in which: g5_weight()
        if "g6" in which: g6_lm()
        if "g7" in which: g7_expm_cholesky()
        if "g8" in which:
            g8_refine("refine_cfg0_240x320_1x1.npz", [0], 240, 320, 1, 1)          # BASELINE configs[0]
            g8_refine("refine_128x160_4x3.npz", [0, 1, 2], 128, 160, 4, 3)
            g8_refine("refine_240x320_4x3.npz", [0], 240, 320, 4, 3)
            g8_refine("refine_occl_128x160_8x3.npz", [5], 128, 160, 8, 3, occlude=True)

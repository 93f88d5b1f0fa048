"""End-to-end parity of the fused inner loop (b200pose_refine_iters through the C-ABI) with
(1) the golden outputs of the executed reference, (2) the oracle on batched inputs, and size-independent
properties at the BASELINE.json sizes.  Tolerance (north_star): final SE3 within 1e-4 abs."""
import os

import numpy as np
import pytest
import torch

from oracle import refine_oracle as O
from rnnpose_b200 import synthetic as S
from tests.util import golden, load_update_weights

pytestmark = pytest.mark.gpu
SE3_TOL = 1e-4


def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def ops():
    from rnnpose_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def packed(ops):
    return ops.pack_weights(load_update_weights(), dev())


def run_gpu(ops, packed, fmap1, fmap2, mb, G0, n_iters, n_lm, **kw):
    d = dev()
    G = G0.clone().to(d).contiguous()
    res = ops.refine_iters(packed, fmap1.to(d).contiguous(), fmap2.to(d).contiguous(), mb["context"].to(d),
                           mb["geofea1"].to(d), mb["geofea2"].to(d), mb["depth"][:, 0].contiguous().to(d),
                           mb["K"].to(d), G, 1.0, n_iters, n_lm, **kw)
    torch.cuda.synchronize()
    return res


@pytest.mark.parametrize("flags", [0, 1], ids=["fp32", "tcgen05"])
@pytest.mark.parametrize("name", ["refine_cfg0_240x320_1x1.npz", "refine_128x160_4x3.npz",
                                  "refine_240x320_4x3.npz", "refine_occl_128x160_8x3.npz"])
def test_refine_matches_reference_golden(ops, packed, name, flags):
    g = golden(name)
    H, W, n_iters, n_lm, seed, occl = [int(v) for v in g["meta"]]
    idxs = [int(i) for i in g["idxs"]]
    mb = S.make_batch(idxs, H, W, seed, bool(occl), with_images=False)
    res = run_gpu(ops, packed, T(g["fmap1"]), T(g["fmap2"]), mb, T(g["G0"]), n_iters, n_lm, want_flows=True, want_weight=True,
                  flags=flags)
    Tij = res["G"].cpu()
    Ti_pred = torch.matmul(Tij, mb["T_init"])                       # PoseRefiner.py:365
    err = (Ti_pred - T(g["Ti_pred"])).abs().max().item()
    print(f"[parity] {name} flags={flags}: max |dSE3| vs executed reference = {err:.3e}")
    assert err < SE3_TOL, f"final SE3 differs from the reference by {err}"
    assert (Tij - T(g["Tij"])).abs().max().item() < SE3_TOL
    torch.testing.assert_close(res["flow_last"].cpu()[:, :, ::4, ::4], T(g["flow_last"]), rtol=1e-3, atol=2e-2)
    torch.testing.assert_close(res["flow_first"].cpu()[:, :, ::4, ::4], T(g["flow_first"]), rtol=1e-3, atol=2e-2)
    torch.testing.assert_close(res["weight"].cpu()[:, ::4, ::4], T(g["weight"]), rtol=1e-3, atol=1e-3)
    # ADD(-S) of prediction vs ground truth must agree with the reference to 4 decimals (in units of the diameter): the
    # device metric kernel on the GPU pose against the reference's own evaluator (utils/eval_metric.py, executed by
    # tests/golden/make_golden_metrics.py) on the reference's pose
    from rnnpose_b200 import metrics as M
    rm = T(golden("metrics.npz")["refine__" + name[:-4]])                 # diameter, ADD, ADD-S, proj2d, flags
    pts = torch.stack([torch.from_numpy(S.model_points(S.make_scene(idx, H, W, seed, bool(occl)))) for idx in idxs])
    met = M.pose_metrics(Ti_pred.cuda(), mb["T_gt"].cuda(), pts.cuda(), rm[:, 0].float().cuda(), torch.arange(len(idxs)).cuda()).cpu().double()
    for col, gcol in ((0, 1), (1, 2)):
        dev_ = ((met[:, col] - rm[:, gcol]).abs() / rm[:, 0]).max().item()
        print(f"[parity] {name} flags={flags}: max |d{M.METRIC_NAMES[col]}| / diameter vs reference evaluator = {dev_:.2e}")
        assert dev_ < 5e-5
    assert torch.equal(met[:, 6], rm[:, 4]) and torch.equal(met[:, 7], rm[:, 5])


@pytest.mark.parametrize("conv_mode", [1, 2, 3], ids=["pair", "vreuse", "pair+vreuse"])
def test_refine_golden_240x320_second_generation_kernels(ops, packed, conv_mode, libopt):
    """The executed reference's 240x320 4x3 result through the second-generation convolution kernel variants."""
    libopt("conv_mode", conv_mode)
    g = golden("refine_240x320_4x3.npz")
    H, W, n_iters, n_lm, seed, occl = [int(v) for v in g["meta"]]
    mb = S.make_batch([int(i) for i in g["idxs"]], H, W, seed, bool(occl), with_images=False)
    res = run_gpu(ops, packed, T(g["fmap1"]), T(g["fmap2"]), mb, T(g["G0"]), n_iters, n_lm, flags=1)
    err = (torch.matmul(res["G"].cpu(), mb["T_init"]) - T(g["Ti_pred"])).abs().max().item()
    print(f"[parity] 240x320 4x3 conv_mode={conv_mode}: max |dSE3| vs executed reference = {err:.3e}")
    assert err < SE3_TOL


def test_refine_batched_vs_oracle(ops, packed):
    """A native batch of 3 different scenes equals three oracle (= reference B=1) runs."""
    H, W, idxs = 128, 160, [7, 8, 9]
    mb = S.make_batch(idxs, H, W, with_images=False)
    f1 = S.hash_features((3, 256, H // 8, W // 8), 61); f2 = S.hash_features((3, 256, H // 8, W // 8), 62)
    G0 = torch.eye(4)[None].repeat(3, 1, 1)
    ref = O.refine_inner_loop(load_update_weights(), f1, f2, mb["context"], mb["geofea1"], mb["geofea2"], mb["depth"],
                              mb["K"], G0, n_iters=2, n_lm=2)
    res = run_gpu(ops, packed, f1, f2, mb, G0, 2, 2, want_flows=True)
    assert (res["G"].cpu() - ref["G"]).abs().max().item() < SE3_TOL
    torch.testing.assert_close(res["flow_last"].cpu(), ref["flows"][-1], rtol=1e-3, atol=2e-2)


@pytest.mark.parametrize("conv_mode", [None, 0, 1, 2, 3], ids=["default", "gen1", "pair", "vreuse", "pair+vreuse"])
def test_refine_full_size_batch_properties(ops, packed, conv_mode, libopt):
    """BASELINE configs[1] shape (B=32, 240x320, 4x3): per-sample results do not depend on batch position or
    batch size (bit-exact), outputs are finite rigid transforms, and a subset agrees with the oracle.  Run for every
    tensor-core convolution kernel variant (B200POSE_CONV_MODE, conv_umma.cu)."""
    if conv_mode is not None:
        libopt("conv_mode", conv_mode)
    H, W, B = 240, 320, 32
    uniq = S.make_batch([0, 1, 2, 3], H, W, with_images=False)
    rep = {k: v.repeat(8, *([1] * (v.dim() - 1))) for k, v in uniq.items() if k != "diameter"}
    f1u = S.hash_features((4, 256, H // 8, W // 8), 71); f2u = S.hash_features((4, 256, H // 8, W // 8), 72)
    f1 = f1u.repeat(8, 1, 1, 1); f2 = f2u.repeat(8, 1, 1, 1)
    G0 = torch.eye(4)[None].repeat(B, 1, 1)
    G = run_gpu(ops, packed, f1, f2, rep, G0, 4, 3)["G"].cpu()
    assert torch.isfinite(G).all()
    for k in range(4, B):
        assert torch.equal(G[k], G[k % 4]), f"sample {k} differs from its copy {k % 4}"
    R = G[:, :3, :3]
    torch.testing.assert_close(torch.matmul(R, R.transpose(1, 2)), torch.eye(3)[None].repeat(B, 1, 1), rtol=0, atol=1e-5)
    assert torch.all(G[:, 3] == torch.tensor([0.0, 0.0, 0.0, 1.0]))
    # batch of 1 == slot 0 of the batch of 32
    one = {k: v[:1].contiguous() for k, v in uniq.items() if k != "diameter"}
    G1 = run_gpu(ops, packed, f1u[:1], f2u[:1], one, G0[:1], 4, 3)["G"].cpu()
    # (the CTA-pair kernel is only taken for machine-filling problems, so B=1 and B=32 may run M=128 and M=256 MMAs)
    assert torch.equal(G1[0], G[0]) or (conv_mode in (None, 1, 3) and (G1[0] - G[0]).abs().max().item() < 1e-6)
    # oracle on two of the samples
    sub = {k: v[:2].contiguous() for k, v in uniq.items() if k != "diameter"}
    ref = O.refine_inner_loop(load_update_weights(), f1u[:2], f2u[:2], sub["context"], sub["geofea1"], sub["geofea2"],
                              sub["depth"], sub["K"], G0[:2], n_iters=4, n_lm=3)
    assert (G[:2] - ref["G"]).abs().max().item() < SE3_TOL


def test_refine_host_entry_matches_device_entry(ops, packed):
    H, W, idxs = 128, 160, [11, 12]
    mb = S.make_batch(idxs, H, W, with_images=False)
    f1 = S.hash_features((2, 256, H // 8, W // 8), 81); f2 = S.hash_features((2, 256, H // 8, W // 8), 82)
    G0 = torch.eye(4)[None].repeat(2, 1, 1)
    Gd = run_gpu(ops, packed, f1, f2, mb, G0, 2, 1)["G"].cpu()
    Gh = G0.clone()
    ops.refine_iters_host(packed, f1, f2, mb["context"], mb["geofea1"], mb["geofea2"], mb["depth"][:, 0].contiguous(),
                          mb["K"], Gh, 1.0, 2, 1)
    assert torch.equal(Gh, Gd)


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("B", [2, 9])
def test_refine_host_entry_with_gathered_context_texels(ops, packed, B, threads):
    """b200pose_refine_iters_host2: host threads gather the texels of the 1/8 resample, only those cross PCIe; and the device
    entry with B200POSE_FLAG_CONTEXT_TEXELS fed by b200pose_context_gather_texels.  Both bit-identical to the plain call
    (B = 9: three sub-batches, the last one ragged; more planes than threads and vice versa)."""
    H, W = 128, 160
    idxs = list(range(30, 30 + B))
    mb = S.make_batch(idxs, H, W, with_images=False)
    f1 = S.hash_features((B, 256, H // 8, W // 8), 85); f2 = S.hash_features((B, 256, H // 8, W // 8), 86)
    G0 = torch.eye(4)[None].repeat(B, 1, 1)
    Gd = run_gpu(ops, packed, f1, f2, mb, G0, 2, 2)["G"].cpu()
    Gh = G0.clone()
    staging = ops.host_staging(B, H, W)
    staging.fill_(float("nan"))
    ops.refine_iters_host(packed, f1, f2, mb["context"], mb["geofea1"], mb["geofea2"], mb["depth"][:, 0].contiguous(),
                          mb["K"], Gh, 1.0, 2, 2, staging=staging, threads=threads)
    assert torch.equal(Gh, Gd)
    tex = ops.context_gather_texels(mb["context"], threads=threads)
    assert torch.equal(tex.flatten(), staging[:tex.numel()])
    mt = dict(mb); mt["context"] = tex
    Gt = run_gpu(ops, packed, f1, f2, mt, G0, 2, 2, flags=ops.DEFAULT_FLAGS | ops.FLAG_CONTEXT_TEXELS)["G"].cpu()
    assert torch.equal(Gt, Gd)


@pytest.mark.parametrize("margin", [0, 8, 24])
def test_refine_host_entry_windowed_second_descriptor_map(ops, packed, libopt, margin):
    """Pinned host buffers: the second descriptor map is copied only inside the foreground box + margin; samples outside it come
    from the mapped host buffer (upsample_weight_pixel<WINDOW>).  Margin 0 with a perturbed initial pose sends many samples
    outside the copied box.  Bit-identical to the device entry, and to the plain copy (sparse_g2 = 0)."""
    H, W, B = 128, 160, 5
    idxs = list(range(40, 40 + B))
    mb = S.make_batch(idxs, H, W, with_images=False)
    f1 = S.hash_features((B, 256, H // 8, W // 8), 87); f2 = S.hash_features((B, 256, H // 8, W // 8), 88)
    G0 = torch.eye(4)[None].repeat(B, 1, 1)
    G0[:, 0, 3] = 0.03; G0[:, 1, 3] = -0.02                               # a few centimetres: flows of tens of pixels
    Gd = run_gpu(ops, packed, f1, f2, mb, G0, 3, 2)["G"].cpu()
    pin = {k: mb[k].contiguous().pin_memory() for k in ("context", "geofea1", "geofea2", "K")}
    depth = mb["depth"][:, 0].contiguous().pin_memory()
    f1p, f2p = f1.pin_memory(), f2.pin_memory()
    staging = ops.host_staging(B, H, W)
    libopt("g2_margin", margin)
    for sparse in (1, 0):
        libopt("sparse_g2", sparse)
        Gh = G0.clone().pin_memory()
        _, scratch = ops.refine_iters_host(packed, f1p, f2p, pin["context"], pin["geofea1"], pin["geofea2"], depth, pin["K"], Gh, 1.0, 3, 2,
                                           staging=staging, threads=2)
        assert torch.equal(Gh, Gd), f"sparse_g2={sparse}"
        # poison the device staging area: whatever the next call does not copy must not be read
        scratch.fill_(255)
        Gh2 = G0.clone().pin_memory()
        ops.refine_iters_host(packed, f1p, f2p, pin["context"], pin["geofea1"], pin["geofea2"], depth, pin["K"], Gh2, 1.0, 3, 2,
                              scratch=scratch, staging=staging, threads=2)
        assert torch.equal(Gh2, Gd), f"sparse_g2={sparse} after poisoning the staging buffers"


@pytest.mark.parametrize("planes", [32, 96, 224])
def test_refine_host_entry_partial_gather(ops, packed, libopt, planes):
    """Option host_gather_planes: the host threads gather only the first `planes` context planes of every object, the kernel
    reads the others in place from the pinned map (ranks sharing a host's cores).  Bit-identical to the device entry."""
    H, W, B = 128, 160, 5
    mb = S.make_batch(list(range(50, 50 + B)), H, W, with_images=False)
    f1 = S.hash_features((B, 256, H // 8, W // 8), 89); f2 = S.hash_features((B, 256, H // 8, W // 8), 90)
    G0 = torch.eye(4)[None].repeat(B, 1, 1)
    Gd = run_gpu(ops, packed, f1, f2, mb, G0, 2, 2)["G"].cpu()
    pin = {k: mb[k].contiguous().pin_memory() for k in ("context", "geofea1", "geofea2", "K")}
    libopt("host_gather_planes", planes)
    staging = ops.host_staging(B, H, W); staging.fill_(float("nan"))
    Gh = G0.clone().pin_memory()
    ops.refine_iters_host(packed, f1.pin_memory(), f2.pin_memory(), pin["context"], pin["geofea1"], pin["geofea2"],
                          mb["depth"][:, 0].contiguous().pin_memory(), pin["K"], Gh, 1.0, 2, 2, staging=staging, threads=3)
    assert torch.equal(Gh, Gd)
    n_used = B * planes * (H // 8) * (W // 8) * 4
    assert torch.isfinite(staging[:n_used]).all() and torch.isnan(staging[n_used:]).all()     # exactly the gathered planes were written


def test_refine_iters_cuda_graph_capture(ops, packed):
    """b200pose_refine_iters is stream-ordered and allocation-free: captured into a CUDA graph (PDL and cluster launches
    included) and replayed, it reproduces the eager result bit for bit, also after the inputs change in place."""
    H, W, idxs = 128, 160, [21, 22]
    mb = S.make_batch(idxs, H, W, with_images=False)
    d = torch.device("cuda:0")
    f1 = S.hash_features((2, 256, H // 8, W // 8), 83).to(d); f2 = S.hash_features((2, 256, H // 8, W // 8), 84).to(d)
    t = {k: mb[k].to(d).contiguous() for k in ("context", "geofea1", "geofea2", "K")}
    depth = mb["depth"][:, 0].contiguous().to(d)
    G0 = torch.eye(4, device=d)[None].repeat(2, 1, 1).contiguous()
    ws = ops.RefineWorkspace(2, H, W, d)

    def call(G):
        ops.refine_iters(packed, f1, f2, t["context"], t["geofea1"], t["geofea2"], depth, t["K"], G, 1.0, 3, 3, workspace=ws)

    Ge = G0.clone(); call(Ge); torch.cuda.synchronize()              # eager (also initialises the library's lazy statics)
    Gs = G0.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            call(Gs)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2):
        Gs.copy_(G0); graph.replay(); torch.cuda.synchronize()
        assert torch.equal(Gs, Ge)
    # new inputs in the same buffers: the graph reads them at replay time
    G1 = O.se3_exp(torch.tensor([[0.01, 0.0, -0.01, 0.02, 0.01, 0.0], [0.0, 0.02, 0.01, -0.01, 0.0, 0.02]])).to(d)
    Ge2 = G1.clone(); call(Ge2); torch.cuda.synchronize()
    Gs.copy_(G1); graph.replay(); torch.cuda.synchronize()
    assert torch.equal(Gs, Ge2) and not torch.equal(Ge2, Ge)


def test_error_codes(ops, packed):
    from rnnpose_b200 import _lib
    L = _lib.lib()
    assert L.b200pose_refine_iters(None, None, None, None, None, None, None, None, None, 1.0, 1, 32, 128, 160, 1, 1,
                                   100.0, 1e-4, 0, None, None, None, None, 0, None) == -1
    mb = S.make_batch([0], 64, 64, with_images=False)       # h/8 = 8 -> level 3 is 1x1: rejected like the reference's NaN
    f = S.hash_features((1, 256, 8, 8), 1)
    with pytest.raises(RuntimeError, match="unsupported shape"):
        run_gpu(ops, packed, f, f, mb, torch.eye(4)[None], 1, 1)


@pytest.mark.parametrize("flags", [0, 1], ids=["fp32", "tcgen05"])
def test_refine_high_res_480x640_vs_oracle(ops, packed, flags):
    """BASELINE configs[4] shape (480x640 crops, 4 pyramid levels: 60x80, 30x40, 15x20, 7x10) against the oracle."""
    H, W = 480, 640
    mb = S.make_batch([21], H, W, with_images=False)
    f1 = S.hash_features((1, 256, H // 8, W // 8), 91); f2 = S.hash_features((1, 256, H // 8, W // 8), 92)
    G0 = torch.eye(4)[None]
    ref = O.refine_inner_loop(load_update_weights(), f1, f2, mb["context"], mb["geofea1"], mb["geofea2"], mb["depth"],
                              mb["K"], G0, n_iters=2, n_lm=2)
    res = run_gpu(ops, packed, f1, f2, mb, G0, 2, 2, flags=flags)
    err = (res["G"].cpu() - ref["G"]).abs().max().item()
    print(f"[parity] 480x640 flags={flags}: max |dSE3| vs oracle = {err:.3e}")
    assert err < SE3_TOL


@pytest.mark.parametrize("flags", [0, 1], ids=["fp32", "tcgen05"])
def test_refine_ragged_size_136x168(ops, packed, flags):
    """h x w = 17 x 21: pixel tiles hang over the image on both axes, the pyramid pools with floor (17x21 -> 8x10 -> 4x5 ->
    2x2) and P = 357 has no MMA-sized divisor, so the tensor-core path falls back to the fp32 volume kernel."""
    H, W = 136, 168
    mb = S.make_batch([31, 32], H, W, with_images=False)
    f1 = S.hash_features((2, 256, H // 8, W // 8), 93); f2 = S.hash_features((2, 256, H // 8, W // 8), 94)
    G0 = torch.eye(4)[None].repeat(2, 1, 1)
    ref = O.refine_inner_loop(load_update_weights(), f1, f2, mb["context"], mb["geofea1"], mb["geofea2"], mb["depth"],
                              mb["K"], G0, n_iters=3, n_lm=2)
    res = run_gpu(ops, packed, f1, f2, mb, G0, 3, 2, flags=flags, want_flows=True)
    err = (res["G"].cpu() - ref["G"]).abs().max().item()
    print(f"[parity] 136x168 flags={flags}: max |dSE3| vs oracle = {err:.3e}")
    assert err < SE3_TOL
    torch.testing.assert_close(res["flow_last"].cpu(), ref["flows"][-1], rtol=1e-3, atol=2e-2)


def test_refine_zero_iterations_and_single_object(ops, packed):
    """n_iters = 0 leaves the pose untouched; B = 1, one iteration, zero LM steps runs the flow net only."""
    H, W = 128, 160
    mb = S.make_batch([33], H, W, with_images=False)
    f = S.hash_features((1, 256, H // 8, W // 8), 95)
    G0 = O.se3_exp(torch.tensor([[0.01, 0.0, 0.02, 0.0, 0.03, 0.0]]))
    assert torch.equal(run_gpu(ops, packed, f, f, mb, G0, 0, 3)["G"].cpu(), G0)
    res = run_gpu(ops, packed, f, f, mb, G0, 1, 0, want_flows=True)
    assert torch.equal(res["G"].cpu(), G0) and torch.isfinite(res["flow_last"]).all()


def test_refine_foreground_list_matches_dense_kernels(ops, packed, libopt):
    """The per-call foreground list (depth > 0 compacted once; upsample + weight and LM run over the list) against the dense
    kernels: identical weights, poses equal up to the fp64 summation order.  Includes an object-free crop (empty list)."""
    H, W = 128, 160
    mb = S.make_batch([41, 42, 43], H, W, with_images=False)
    mb["depth"][1] = 0.0                                    # sample 1: no foreground at all
    f1 = S.hash_features((3, 256, H // 8, W // 8), 96); f2 = S.hash_features((3, 256, H // 8, W // 8), 97)
    G0 = torch.eye(4)[None].repeat(3, 1, 1)
    libopt("fg_list", 0)
    dense = run_gpu(ops, packed, f1, f2, mb, G0, 3, 3, want_weight=True, want_flows=True)
    libopt("fg_list", 1); libopt("fg_pipeline", 0); libopt("lm_cluster", 0)     # list + the spin-barrier LM kernel
    fg0 = run_gpu(ops, packed, f1, f2, mb, G0, 3, 3, want_weight=True)
    assert (fg0["G"].cpu() - dense["G"].cpu()).abs().max().item() < 1e-6
    libopt("lm_cluster", 1)                                                      # list + the cluster LM kernel (default)
    fg = run_gpu(ops, packed, f1, f2, mb, G0, 3, 3, want_weight=True)
    # default: the foreground pipeline (channels-last descriptors, float4 records, cluster LM kernel), with and without the
    # dense outputs requested, tensor-core and exact-fp32 convolutions
    libopt("fg_pipeline", 1)
    for kw in (dict(want_weight=True, want_flows=True), dict()):
        pipe = run_gpu(ops, packed, f1, f2, mb, G0, 3, 3, **kw)
        assert torch.equal(pipe["G"][1].cpu(), G0[1])
        assert (pipe["G"].cpu() - dense["G"].cpu()).abs().max().item() < 1e-6
        if kw:
            # (the weights of the LAST iteration sit behind two LM rounds whose poses differ by ~1e-7 between the variants)
            torch.testing.assert_close(pipe["weight"].cpu(), dense["weight"].cpu(), rtol=0, atol=2e-5)
            assert torch.equal(pipe["flow_first"].cpu(), dense["flow_first"].cpu())
            torch.testing.assert_close(pipe["flow_last"].cpu(), dense["flow_last"].cpu(), rtol=0, atol=1e-4)   # poses differ ~1e-7
    libopt("fg_pipeline", 0); d32 = run_gpu(ops, packed, f1, f2, mb, G0, 3, 3, flags=0)
    libopt("fg_pipeline", 1); p32 = run_gpu(ops, packed, f1, f2, mb, G0, 3, 3, flags=0)
    assert (p32["G"].cpu() - d32["G"].cpu()).abs().max().item() < 1e-6
    libopt("fg_pipeline", 0)
    assert torch.equal(fg["G"][1].cpu(), G0[1]) and torch.equal(dense["G"][1].cpu(), G0[1])
    assert (fg["G"].cpu() - dense["G"].cpu()).abs().max().item() < 1e-6
    torch.testing.assert_close(fg["weight"].cpu(), dense["weight"].cpu(), rtol=0, atol=2e-6)
    # opt-in: the persistent list-driven upsample + weight kernel (same per-pixel function, background set once per call)
    libopt("fg_upsample", 1)
    fgu = run_gpu(ops, packed, f1, f2, mb, G0, 3, 3, want_weight=True)
    assert (fgu["G"].cpu() - dense["G"].cpu()).abs().max().item() < 1e-6
    torch.testing.assert_close(fgu["weight"].cpu(), dense["weight"].cpu(), rtol=0, atol=2e-6)
    ref = O.refine_inner_loop(load_update_weights(), f1, f2, mb["context"], mb["geofea1"], mb["geofea2"], mb["depth"],
                              mb["K"], G0, n_iters=3, n_lm=3)
    assert (fg["G"].cpu() - ref["G"]).abs().max().item() < SE3_TOL


@pytest.mark.parametrize("rings,dynamic,xmajor,merge", [(24, 0, 1, 0), (24, 1, 1, 1), (33, 0, 0, 0), (24, 1, 0, 1), (24, 0, 1, 1)],
                         ids=["default", "unit-queue-interleaved", "rings3+3-ymajor", "unit-queue-ymajor-interleaved", "interleaved"])
def test_refine_chained_convolutions(ops, packed, libopt, rings, dynamic, xmajor, merge):
    """The eleven convolutions of a pass in one persistent launch with tile-level dependencies (conv_mode 19, the default)
    against the layer-by-layer launches (conv_mode 3) at the bench shape (the chain needs a machine-filling batch), for both
    shared-memory ring geometries."""
    H, W, B = 240, 320, 32
    uniq = S.make_batch([0, 1, 2, 3], H, W, with_images=False)
    rep = {k: v.repeat(8, *([1] * (v.dim() - 1))) for k, v in uniq.items() if k != "diameter"}
    f1 = S.hash_features((4, 256, H // 8, W // 8), 71).repeat(8, 1, 1, 1); f2 = S.hash_features((4, 256, H // 8, W // 8), 72).repeat(8, 1, 1, 1)
    G0 = torch.eye(4)[None].repeat(B, 1, 1)
    libopt("conv_mode", 3)
    ref = run_gpu(ops, packed, f1, f2, rep, G0, 4, 3)["G"].cpu()
    libopt("conv_mode", 19); libopt("chain_rings", rings); libopt("chain_dynamic", dynamic); libopt("chain_xmajor", xmajor); libopt("chain_merge", merge)
    got = run_gpu(ops, packed, f1, f2, rep, G0, 4, 3)["G"].cpu()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 1e-6
    for k in range(4, B):
        assert torch.equal(got[k], got[k % 4])
    # an odd number of pixel tiles (99 x 3 = 297: the last CTA pair has a dummy half) and a 480x640 batch
    for (Bo, Ho, Wo) in ((99, 128, 160), (10, 480, 640)):
        u2 = S.make_batch([0, 1], Ho, Wo, with_images=False)
        r2 = {k: v.repeat((Bo + 1) // 2, *([1] * (v.dim() - 1)))[:Bo].contiguous() for k, v in u2.items() if k != "diameter"}
        g1 = S.hash_features((2, 256, Ho // 8, Wo // 8), 73).repeat((Bo + 1) // 2, 1, 1, 1)[:Bo].contiguous()
        g2 = S.hash_features((2, 256, Ho // 8, Wo // 8), 74).repeat((Bo + 1) // 2, 1, 1, 1)[:Bo].contiguous()
        Go = torch.eye(4)[None].repeat(Bo, 1, 1)
        libopt("conv_mode", 3)
        a = run_gpu(ops, packed, g1, g2, r2, Go, 2, 2)["G"].cpu()
        libopt("conv_mode", 19)
        b = run_gpu(ops, packed, g1, g2, r2, Go, 2, 2)["G"].cpu()
        assert (a - b).abs().max().item() < 1e-6 and torch.equal(b[0], b[2])

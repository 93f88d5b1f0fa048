#!/usr/bin/env python
"""bench.py -- refined poses/sec of the RNNPose recurrent pose-refinement inner loop on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
(N>1: launched under torch.distributed.run, one rank per GPU).  One JSON line on rank 0.

A "step" = one pass of the hot path (b200pose_refine_iters: ITER_COUNT recurrent iterations x OPTIM_ITER_COUNT LM steps)
over one batch of synthetic crop pairs per GPU.  Workloads (BASELINE.json configs):
  --config cfg1  (default)  32 objects/GPU, 240x320, 4 x 3            configs[1]; at --gpus 8 this is configs[2] (256 objects)
  --config cfg3             32 objects/GPU, 240x320, 8 x 3, occluded scenes, sample_poses-style perturbed initial poses
                            (configs[3]: 128 objects on 4 GPUs)
  --config cfg4             480x640, 4 x 3; --global-batch B (8..512) or --sweep for the whole batch-size sweep (configs[4])
  --global-batch B          total objects over all ranks (e.g. --config cfg1 --global-batch 256 at N=1: the strong-scaling point)
Batches larger than --chunk objects per GPU are processed in chunks (every chunk is real work on resident inputs).
  value : poses/s with the loop's inputs resident in HBM (CUDA events, max over ranks)
  e2e   : poses/s through the host-buffer C-ABI entry (pinned host inputs -> H2D -> loop -> D2H of the poses)
  roofline / cpu_baseline : see DESIGN.md section "Measurement"
--impl reference : the CPU oracle port (restatement of the reference's PyTorch CPU path; /root/reference
does not exist on the GPU box) timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU legs (reference arm, cpu_baseline) must be
# free to use the host cores, so lift that before torch / OpenMP initialise (the pool size is then probed).
if os.environ.get("OMP_NUM_THREADS") == "1":
    try:
        _n = len(os.sched_getaffinity(0))
    except Exception:
        _n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(_n)
    os.environ["MKL_NUM_THREADS"] = str(_n)

import torch  # noqa: E402

UNIQUE_SCENES = 8            # distinct synthetic scenes per rank, tiled to the batch
FLOP_PER_LOWRES_PX = 6236672  # update-block convolutions, SURVEY.md Appendix A.2
# dram__bytes_read.sum + dram__bytes_write.sum of the convolution launch(es) of one update-block pass at B=32, 240x320, from
# the committed `ncu --set full` capture named in TRAFFIC_SOURCE (ncu flushes caches per launch: an upper bound)
NCU_TRAFFIC_BYTES_PER_PASS = {"chain": 939_701_248, "layers": 775_000_000}
TRAFFIC_SOURCE = {"chain": "profiles/r2h/chain_ncu_raw.csv (conv_chain_kernel incl. the flow head, one launch, B=32, 240x320; ncu flushes caches per launch)",
                  "layers": "profiles/r1c_conv_umma2_ncu_raw.csv (11 conv launches of one pass, B=32, 240x320)"}

CONFIGS = {
    #        H    W   iters lm  objects/GPU occluded  chunk
    "cfg1": (240, 320, 4, 3, 32, False, 64),
    "cfg3": (240, 320, 8, 3, 32, True, 64),
    "cfg4": (480, 640, 4, 3, 8, False, 16),
}
SWEEP_BATCHES = (8, 16, 32, 64, 128, 256, 512)


def workload_name(cfg, per_gpu):
    H, W, it, lm, _, occ, _ = CONFIGS[cfg]
    return (f"synthetic {H}x{W} crops, batch {per_gpu}/GPU, {it} recurrent iters x {lm} LM steps" +
            (", occluded targets, perturbed initial poses" if occ else ""))


def load_weights():
    from rnnpose_b200.assets import load_update_weights
    return load_update_weights()


def make_inputs(rank: int, batch: int, unique: int, H: int, W: int, occlude: bool = False, fmap_fn=None):
    """CPU float32 inputs of the inner loop for `batch` objects (unique scenes tiled).  fmap_fn(syn_img, obs_img) -> (fmap1,
    fmap2) supplies the feature maps (the library's encoder on the GPU); without it they are hash noise."""
    from rnnpose_b200 import synthetic as S
    unique = min(unique, batch)
    idx = [rank * unique + i for i in range(unique)]
    mb = S.make_batch(idx, H, W, occlude=occlude, with_images=fmap_fn is not None)
    h, w = H // 8, W // 8
    if fmap_fn is not None:
        f1, f2 = fmap_fn(mb.pop("syn_img"), mb.pop("obs_img"))
    else:
        f1, f2 = S.hash_features((unique, 256, h, w), 9000 + rank), S.hash_features((unique, 256, h, w), 9500 + rank)
    rep = (batch + unique - 1) // unique
    out = {k: v.repeat(rep, *([1] * (v.dim() - 1)))[:batch].contiguous() for k, v in mb.items()}
    out["fmap1"] = f1.repeat(rep, 1, 1, 1)[:batch].contiguous()
    out["fmap2"] = f2.repeat(rep, 1, 1, 1)[:batch].contiguous()
    out["depth"] = out["depth"][:, 0].contiguous()
    out["G0"] = torch.eye(4)[None].repeat(batch, 1, 1).contiguous()
    out["scene_idx"] = torch.tensor(idx).repeat(rep)[:batch]
    return out


def encoder_flops_per_image(H, W):
    """Algorithmic FLOPs of BasicEncoder on one H x W image (thirdparty/raft/extractor.py:118-232 layer shapes)."""
    d = lambda v: (v - 1) // 2 + 1
    H1, W1 = d(H), d(W); H2, W2 = d(H1), d(W1); H3, W3 = d(H2), d(W2)
    f = H1 * W1 * (147 * 64 + 4 * 9 * 64 * 64)
    f += H2 * W2 * (9 * 64 * 96 + 3 * 9 * 96 * 96 + 64 * 96)
    f += H3 * W3 * (9 * 96 * 128 + 3 * 9 * 128 * 128 + 96 * 128 + 128 * 256)
    return 2 * f


def time_encoder(ops, dev, batch, H, W, peaks):
    """The f2 row's bench leg: ImageFeaEncoder on `batch` crop pairs (2 x batch images), CUDA events, inputs resident."""
    try:
        from rnnpose_b200.assets import load_encoder_weights
        packed = ops.encoder_pack_weights(load_encoder_weights(), dev)
        a = torch.rand(batch, 3, H, W, device=dev) * 255; b = torch.rand(batch, 3, H, W, device=dev) * 255
        nb = ops._lib.lib().b200pose_encoder_workspace_bytes(batch, H, W)
        ws = ops._ws(nb, dev)
        for _ in range(2):
            ops.image_encoder(packed, a, b, workspace=ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        reps = 5
        for _ in range(reps):
            ops.image_encoder(packed, a, b, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = 2 * batch * encoder_flops_per_image(H, W) / (ms * 1e-3) / 1e12
        return {"ms_per_batch": ms, "pairs_per_s": batch / (ms * 1e-3), "algorithmic_tflops": tf,
                "frac_of_burst_bf16_peak": tf / float(peaks.get("bf16_tflops")), "batch_pairs": batch,
                "note": "b200pose_image_encoder (f2): 16 tcgen05 convolutions (fp16x3 split; the 7x7 stem re-indexed as a 4x1 convolution of gathered channels) + InstanceNorm passes; not part of `value`"}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def encoder_fmap_fn(dev):
    """Feature maps from the library's own RAFT encoder kernels (csrc/encoder.cu, shipped img_fea_enc weights) on the synthetic
    crop pair, computed once before the timed region (the metric starts from resident feature maps).  Returns
    (fn, description); falls back to hash-noise maps if the encoder cannot run."""
    try:
        from rnnpose_b200 import ops
        from rnnpose_b200.assets import load_encoder_weights
        packed = ops.encoder_pack_weights(load_encoder_weights(), dev)

        def fn(a, b):
            f1, f2 = ops.image_encoder(packed, a.to(dev).contiguous(), b.to(dev).contiguous())
            torch.cuda.synchronize()
            if not (torch.isfinite(f1).all() and torch.isfinite(f2).all()):
                raise RuntimeError("non-finite encoder output")
            return f1.cpu(), f2.cpu()
        return fn, "feature maps from the library's RAFT encoder kernels on the synthetic crops (shipped img_fea_enc weights)"
    except Exception as e:  # noqa: BLE001
        return None, f"hash-noise feature maps (encoder unavailable: {type(e).__name__})"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled every ~2 ms from
    a thread (the timed region is tens of milliseconds: `nvidia-smi -lms 100` can miss it entirely); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu, self.nvml, self.h, self.run = [], None, gpu_index, None, None, False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(self.gpu)
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.h = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.run = True
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown, n.nvmlClocksEventReasonSwThermalSlowdown,
                n.nvmlClocksEventReasonSwPowerCap]
        while self.run:
            try:
                t = time.time()
                mhz = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1e3
                self.rows.append((t, [str(mhz), str(self.max_mhz), str(pw)] + ["Active" if r & b else "Not Active" for b in bits]))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        if self.nvml is not None:
            self.run = False
            rows = [r for (t, r) in self.rows if t0 <= t <= t1]
            src = "nvml, 2 ms period, inside the timed region"
        else:
            time.sleep(0.15)
            self.proc.terminate()
            rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
            src = "nvidia-smi -lms 100"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "source": src}
        sm = sorted(float(r[0]) for r in rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = max(float(r[2]) for r in rows)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(rows[0][1]), "power_w_max": pw,
                "samples": len(rows), "reasons": reasons, "source": src}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def bind_numa(dev_index: int, world: int = 1):
    """Bind this rank's CPU threads and its future host allocations (the pinned e2e buffers) to the NUMA node of its GPU,
    BEFORE anything is pinned: at N=8 every rank otherwise allocates on the node it happens to start on and half of the
    H2D traffic crosses the socket interconnect (round-1 SCALE run: e2e efficiency 0.49 with all ranks on node 0).
    Best effort: reports what it could do."""
    info = {"gpu_numa_node": None, "cpus_bound": None, "mempolicy": "unchanged"}
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bdf = None
        if hasattr(pr, "pci_bus_id") and hasattr(pr, "pci_device_id"):
            bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            q = subprocess.run(["nvidia-smi", "--query-gpu=uuid,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
            uuid = str(getattr(pr, "uuid", ""))
            for line in q.splitlines():
                u, b = [x.strip() for x in line.split(",")]
                if uuid and uuid in u:
                    bdf = b[-12:].lower()
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read()) if bdf else -1
        info["gpu_numa_node"] = node
        if node < 0:
            # the box does not say which node the GPU hangs off: spread the ranks over the visible nodes in device order (GPUs
            # 0..n/2-1 on socket 0 is the usual board layout) so that at least the host DRAM bandwidth of every socket is used
            nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
            if len(nodes) < 2 or world < 2:
                return info
            node = nodes[min(len(nodes) - 1, dev_index * len(nodes) // world)]
            info["guessed_node"] = node
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus_bound"] = len(allowed)
        else:
            info["cpus_bound"] = 0                      # the node's cores are outside this container's cpuset
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, NR_set_mempolicy = 1, 238
        rc = libc.syscall(NR_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        info["mempolicy"] = f"preferred node {node}" if rc == 0 else f"set_mempolicy failed (errno {ctypes.get_errno()})"
    except Exception as e:  # noqa: BLE001
        info["error"] = f"{type(e).__name__}: {e}"
    return info


_THREADS = None


def pick_cpu_threads(wts) -> int:
    """"All the host threads it can use": torch's intra-op pool gets slower, not faster, when it is
    oversubscribed on the small tensors of this path, so probe a few pool sizes on one update-block call
    (the FLOP-dominant piece) and keep the fastest.  The choice is reported as `cores`."""
    global _THREADS
    if _THREADS is not None:
        return _THREADS
    from oracle import refine_oracle as O
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (avail, 64, 32, 16, 8) if 1 <= c <= avail}, reverse=True)
    g = torch.Generator().manual_seed(0)
    net = torch.randn(1, 128, 30, 40, generator=g); inp = torch.randn(1, 128, 30, 40, generator=g)
    corr = torch.randn(1, 324, 30, 40, generator=g); flow = torch.randn(1, 2, 30, 40, generator=g)
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            O.update_block(wts, net, inp, corr, flow)
            t0 = time.time()
            for _ in range(3):
                O.update_block(wts, net, inp, corr, flow)
            dt = time.time() - t0
            if dt < best_t:
                best, best_t = c, dt
    _THREADS = best
    torch.set_num_threads(best)
    return best


def cpu_oracle_rate(inputs, n_objects: int, wts, n_iters: int, n_lm: int):
    """Oracle port on the host cores: `n_objects` objects, one reference-style B=1 call each."""
    from oracle import refine_oracle as O
    pick_cpu_threads(wts)
    t0 = time.time()
    outs = []
    with torch.no_grad():
        for i in range(n_objects):
            sl = slice(i, i + 1)
            # the variant that issues the reference's own ATen op sequence (grid_sample, interpolate, einsum f64, ...)
            r = O.refine_inner_loop_aten(wts, inputs["fmap1"][sl], inputs["fmap2"][sl], inputs["context"][sl],
                                         inputs["geofea1"][sl], inputs["geofea2"][sl], inputs["depth"][sl][:, None],
                                         inputs["K"][sl], inputs["G0"][sl], sigma=1.0, n_iters=n_iters, n_lm=n_lm)
            outs.append(r["G"])
    dt = time.time() - t0
    return n_objects / dt, dt, torch.cat(outs)


def metric_name(cfg):
    H, W, it, lm = CONFIGS[cfg][:4]
    return f"refined poses/sec ({it} recur iters x {lm} LM steps, {H}x{W})"


def run_reference(args):
    """--impl reference: CPU oracle port, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    H, W, n_iters, n_lm, per_gpu, occl, _ = CONFIGS[args.config]
    per_step = 4 if H * W <= 240 * 320 else 1
    inputs = make_inputs(0, per_step, per_step, H, W, occl)
    wts = load_weights()
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_oracle_rate(inputs, 1, wts, n_iters, n_lm)
    t0 = time.time()
    n = 0
    for _ in range(args.steps):
        cpu_oracle_rate(inputs, per_step, wts, n_iters, n_lm); n += per_step
    dt = time.time() - t0
    v = n / dt
    cores = torch.get_num_threads()
    per = args.global_batch // max(1, args.gpus) if args.global_batch else per_gpu
    line = {"impl": "reference", "metric": metric_name(args.config), "value": v,
            "unit": "poses/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (LM step f64)", "data": "synthetic",
            "config": {"workload": workload_name(args.config, per), "sample": f"{per_step} objects per step, one B=1 call each"},
            "cpu_baseline": {"value": v, "unit": "poses/s", "cores": cores, "kind": "port",
                             "sample": f"{n} objects x ({n_iters}x{n_lm}) at {H}x{W}, oracle/refine_oracle.py::refine_inner_loop_aten, torch CPU fp32"},
            "e2e": {"value": v, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def conv_pass_roofline(ops, packed, Bc, H, W, FLAGS, peaks, peak_src, exact, ms_step, n_iters, passes_per_step):
    from rnnpose_b200 import _lib
    """Roofline of the dominant kernel -- the convolutions of one update-block pass (one chained launch by default) -- timed
    live at the chunk batch size with CUDA events around the launch.  The passes run ALONE (not inside the long step), so the
    denominator is the BURST dense-bf16 peak of MEASURED_PEAKS.json; the sustained-peak fraction is reported next to it."""
    dev = packed.device
    h, w = H // 8, W // 8
    P = Bc * h * w
    net = torch.tanh(torch.randn(P, 128, device=dev)); xbuf = torch.relu(torch.randn(P, 256, device=dev))
    corr = torch.randn(P, 328, device=dev); c1 = torch.randn(P, 2, device=dev); fl = torch.randn(P, 2, device=dev)
    for _ in range(3):
        ops.update_block(packed, net, xbuf, corr, c1, fl, Bc, h, w, flags=FLAGS)
    reps = 10
    # (1) the convolution launch(es) alone: CUDA events recorded by the library on its stream immediately around them
    #     (b200pose_debug_set_conv_events); (2) the whole pass incl. its helper launches, for the share of the step
    L = _lib.lib()
    pairs = []
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    r0.record()
    for _ in range(reps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(); k1.record()                       # creates the handles; the library re-records them
        if not exact:
            L.b200pose_debug_set_conv_events(k0.cuda_event, k1.cuda_event)
        ops.update_block(packed, net, xbuf, corr, c1, fl, Bc, h, w, flags=FLAGS)
        pairs.append((k0, k1))
    r1.record(); torch.cuda.synchronize()
    ub_ms = r0.elapsed_time(r1) / reps
    conv_ms = ub_ms if exact else sum(a.elapsed_time(b) for a, b in pairs) / reps
    flops = FLOP_PER_LOWRES_PX * P
    achieved = flops / (conv_ms * 1e-3) / 1e12
    burst = float(peaks.get("bf16_tflops"))
    sustained = float(peaks.get("bf16_tflops_sustained", burst))
    chain = (ops.get_option("conv_mode") & 16) != 0
    kern = ("conv_gemm_kernel<128|64> (FFMA)" if exact else
            ("conv_chain_kernel (the 11 convolutions of a pass in one persistent launch of CTA pairs, tcgen05.mma cta_group::2 + TMA + TMEM)"
             if chain else "conv_umma2_kernel (tcgen05.mma cta_group::2 on CTA pairs + TMA + TMEM), 11 launches"))
    same_shape = (Bc, H, W) == (32, 240, 320) and not exact
    tkey = "chain" if chain else "layers"
    return {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
            "frac_of_sustained_peak": achieved / sustained,
            "traffic": NCU_TRAFFIC_BYTES_PER_PASS[tkey] if same_shape else None,
            "traffic_source": TRAFFIC_SOURCE[tkey] if same_shape else None,
            "peak_source": f"{peak_src}: dense bf16 burst (the pass is timed alone); sustained {sustained:.1f}",
            "kernel": kern + "; CUDA events recorded by the library on its stream immediately before / after the convolution launch(es) of a pass",
            "algorithmic_flops_per_pass": flops, "ms_per_launch": conv_ms, "ms_per_pass_with_helpers": ub_ms, "batch": Bc,
            "share_of_step": (conv_ms * n_iters * passes_per_step) / ms_step}


def run_workload(args, cfg, per_gpu, rank, local_rank, world, dev, numa, want_cpu=True, want_e2e=True):
    """One workload on this rank's GPU; returns the JSON line (rank 0) or None."""
    from rnnpose_b200 import dist as D, metrics as M, ops, synthetic as S
    H, W, N_ITERS, N_LM, _, occl, chunk_cap = CONFIGS[cfg]
    chunk = min(per_gpu, args.chunk or chunk_cap)
    n_chunks = (per_gpu + chunk - 1) // chunk
    assert per_gpu % chunk == 0, f"--global-batch per GPU ({per_gpu}) must be a multiple of the chunk ({chunk})"
    FLAGS = ops.FLAG_EXACT_FP32 if args.exact_fp32 else ops.FLAG_TENSOR_CORES
    dtype = ("f32 (CUDA-core FFMA convolutions; LM step f64)" if args.exact_fp32 else
             "f32-equivalent: tcgen05 kind::f16 on fp16 hi/lo split operands (22-bit), 3 MMAs, fp32 TMEM accumulate; LM step f64")

    # resident inputs: one chunk's worth per distinct chunk (up to 4 distinct chunks; more chunks re-use them round robin --
    # every chunk is still far larger than the 126 MB L2)
    n_res = min(n_chunks, 4 if H * W <= 240 * 320 else 1)
    fmap_fn, fmap_desc = (None, "hash-noise feature maps") if args.fmaps == "hash" else encoder_fmap_fn(dev)
    try:
        inputs = [make_inputs(rank * n_res + c, chunk, UNIQUE_SCENES, H, W, occl, fmap_fn) for c in range(n_res)]
    except Exception as e:  # noqa: BLE001  (the benchmark of the loop must not depend on the encoder row)
        fmap_desc = f"hash-noise feature maps (encoder failed: {type(e).__name__})"
        inputs = [make_inputs(rank * n_res + c, chunk, UNIQUE_SCENES, H, W, occl) for c in range(n_res)]
    keys = ("fmap1", "fmap2", "context", "geofea1", "geofea2", "depth", "K", "G0")
    host = {k: inputs[0][k].pin_memory() for k in keys}                       # e2e: chunk 0's buffers, pinned after bind_numa
    d = [{k: (host[k] if c == 0 else inputs[c][k]).to(dev, non_blocking=True) for k in keys} for c in range(n_res)]
    wts = load_weights()
    packed = ops.pack_weights(wts, dev)
    ws = ops.RefineWorkspace(chunk, H, W, dev)
    Gs = [d[c % n_res]["G0"].clone() for c in range(n_chunks)]

    def step():
        for c in range(n_chunks):
            t = d[c % n_res]
            Gs[c].copy_(t["G0"])
            ops.refine_iters(packed, t["fmap1"], t["fmap2"], t["context"], t["geofea1"], t["geofea2"], t["depth"], t["K"], Gs[c],
                             1.0, N_ITERS, N_LM, workspace=ws, flags=FLAGS)

    # everything the closing metric gather needs is resident before the timed region
    T_init_d = torch.cat([inputs[c % n_res]["T_init"] for c in range(n_chunks)]).to(dev)
    T_gt_d = torch.cat([inputs[c % n_res]["T_gt"] for c in range(n_chunks)]).to(dev)
    diam_d = torch.cat([inputs[c % n_res]["diameter"] for c in range(n_chunks)]).to(dev)
    sidx = torch.cat([inputs[c % n_res]["scene_idx"] for c in range(n_chunks)])
    sidx_d = sidx.to(dev)
    scene_pts = {int(i): torch.from_numpy(S.model_points(S.make_scene(int(i), H, W, occlude=occl))) for i in sidx.unique()}
    pts = torch.stack([scene_pts[int(i)] for i in sidx]).to(dev)

    def gather_metrics():
        """per-object metrics + the single all-gather that closes the job (reference tools/train.py:724-741)"""
        met = M.pose_metrics(torch.matmul(torch.cat(Gs), T_init_d), T_gt_d, pts, diam_d, sidx_d)
        return D.all_gather_metrics(met)

    for _ in range(args.warmup):
        step()
    gather_metrics()
    torch.cuda.synchronize(); D.barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); D.barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    gm = gather_metrics()                 # inside the timed region: poses/s includes the closing NCCL all-gather
    ev1.record()
    torch.cuda.synchronize(); D.barrier()
    t_wall1 = time.time()
    ms = D.max_over_ranks(ev0.elapsed_time(ev1), dev)
    clocks = sampler.stop(t_wall0, t_wall1)
    value = world * per_gpu * args.steps / (ms * 1e-3)
    G_dev0 = Gs[0].clone()

    # ---- e2e: host buffers through the C-ABI host entry (H2D + loop + D2H inside the timed region), chunk by chunk
    e2e = None
    if want_e2e:
        ke = args.e2e_steps or max(3, min(args.steps, 10))
        Gh = host["G0"].clone().pin_memory()
        scratch = None
        del ws
        cores = len(os.sched_getaffinity(0))
        if args.host_gather_threads is not None:
            variants = [(args.host_gather_threads, 256)]
        else:                                      # the ways of not copying the whole context map; the fastest one is the line's e2e
            tg = max(1, min(8, cores // max(1, world)))
            variants = [(-1, 0), (tg, 256), (tg, 128)]          # (host threads, context planes per object they gather)
        staging = ops.host_staging(chunk, H, W) if any(t >= 0 for t, _ in variants) else None

        def e2e_step(threads):
            nonlocal scratch
            for _c in range(n_chunks):
                Gh.copy_(host["G0"])
                _, scratch = ops.refine_iters_host(packed, host["fmap1"], host["fmap2"], host["context"], host["geofea1"],
                                                   host["geofea2"], host["depth"], host["K"], Gh, 1.0, N_ITERS, N_LM,
                                                   scratch=scratch, flags=FLAGS, staging=staging if threads >= 0 else None,
                                                   threads=max(threads, 0))

        # bytes that cross PCIe per chunk: cudaMemcpy of every input except the context map, plus either the context rows the
        # context-init kernel reads directly from the pinned host buffer (rows floor(y*s) and +1 of each 1/8-res row) or the
        # texels gathered by the host threads (4 floats per channel and low-res pixel)
        sy = (H - 1) / (H // 8 - 1)
        rows = set()
        for y in range(H // 8):
            y0 = min(int(y * sy), H - 1); rows.update((y0, min(y0 + 1, H - 1)))
        sparse_g1 = ops.get_option("sparse_g1") != 0
        g1_bytes = (int((host["depth"] > 0).sum()) * host["geofea1"].shape[1] * 4) if sparse_g1 else host["geofea1"].numel() * 4
        g2_bytes = host["geofea2"].numel() * 4
        if ops.get_option("sparse_g2") != 0:       # second descriptor map: only the foreground box + margin is copied (api.cu geo2_window)
            mg = max(0, ops.get_option("g2_margin"))
            g2_bytes = 0
            fgm = host["depth"] > 0
            for b in range(fgm.shape[0]):
                ys = torch.nonzero(fgm[b].any(dim=1)).flatten(); xs = torch.nonzero(fgm[b].any(dim=0)).flatten()
                if ys.numel() == 0:
                    continue
                y0, y1 = max(0, int(ys[0]) - mg), min(H, int(ys[-1]) + 1 + mg)
                x0, x1 = max(0, int(xs[0]) - mg) & ~7, min(W, (min(W, int(xs[-1]) + 1 + mg) + 7) & ~7)
                g2_bytes += (y1 - y0) * (x1 - x0) * host["geofea2"].shape[1] * 4
        other = sum(host[k].numel() * 4 for k in ("fmap1", "fmap2", "depth", "K", "G0")) + g1_bytes + g2_bytes
        d2h = n_chunks * Gh.numel() * 4
        runs = []
        for threads, planes in variants:
            if threads >= 0:
                ops.set_option("host_gather_planes", planes)
            e2e_step(threads)
            torch.cuda.synchronize(); D.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(ke):
                e2e_step(threads)
            e1.record()
            torch.cuda.synchronize(); D.barrier()
            my_ms = e0.elapsed_time(e1)
            ms_e2e = D.max_over_ranks(my_ms, dev)
            n_tex = planes if threads >= 0 else 0                      # planes gathered by the host; the rest is read in place, rows only
            ctx_bytes = chunk * (n_tex * (H // 8) * (W // 8) * 4 + (256 - n_tex) * len(rows) * W) * 4
            h2d = n_chunks * (other + ctx_bytes)
            runs.append({"context": "mapped rows (zero-copy)" if threads < 0 else
                         (f"texels gathered by {threads} host threads" if planes == 256 else
                          f"{planes} of 256 planes as texels gathered by {threads} host threads, the rest as mapped rows"),
                         "gathered_planes": n_tex,
                         "host_gather_threads": threads, "value": world * per_gpu * ke / (ms_e2e * 1e-3),
                         "ms_per_step": ms_e2e / ke, "h2d_bytes_per_step": h2d,
                         "h2d_gbs_this_rank": h2d * ke / (my_ms * 1e-3) / 1e9,
                         "max_abs_diff_vs_device_entry": (Gh.to(dev) - G_dev0).abs().max().item()})
        ops.set_option("host_gather_planes", 256)
        best = max(runs, key=lambda r: r["value"])
        e2e = {"value": best["value"], "unit": "poses/s", "h2d_bytes_per_step": best["h2d_bytes_per_step"], "d2h_bytes_per_step": d2h,
               "steps": ke, "ms_per_step": best["ms_per_step"], "max_abs_diff_vs_device_entry": max(r["max_abs_diff_vs_device_entry"] for r in runs),
               "h2d_gbs_this_rank": best["h2d_gbs_this_rank"], "context": best["context"], "variants": runs, "host_cpus": cores,
               "host_input_bytes": n_chunks * sum(host[k].numel() * 4 for k in host), "numa": numa,
               "note": "the context map [B,256,H,W] is never copied whole: either its needed rows are read in place from pinned host memory by the kernel, or host threads gather the 4 texels per low-res sample into a pinned staging buffer (b200pose_refine_iters_host2); the first descriptor map is read only at the pixels with depth > 0, the second one is copied only inside the foreground box + margin (samples outside it are read in place; outside-window traffic is not counted)"}
        del scratch
    else:
        del ws

    peaks, peak_src = measured_peaks()
    enc_leg = time_encoder(ops, dev, chunk, H, W, peaks) if (rank == 0 and want_cpu and args.fmaps != "hash") else None
    roofline = conv_pass_roofline(ops, packed, chunk, H, W, FLAGS, peaks, peak_src, args.exact_fp32, ms / args.steps, N_ITERS, n_chunks)

    line = None
    if rank == 0:
        cpu = None
        oracle_check = None
        if world == 1 and want_cpu:
            n_cpu = max(1, min(args.cpu_objects, chunk)) if H * W <= 240 * 320 else 1
            rate, dt, G_cpu = cpu_oracle_rate(inputs[0], n_cpu, wts, N_ITERS, N_LM)
            cpu = {"value": rate, "unit": "poses/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{n_cpu} objects x ({N_ITERS}x{N_LM}) at {H}x{W} in {dt:.1f}s, oracle/refine_oracle.py::refine_inner_loop_aten (reference ATen op sequence, torch CPU fp32, LM fp64)"}
            k = min(2, n_cpu)
            oracle_check = {"objects": k, "max_abs_dSE3_vs_oracle": float((G_dev0[:k].cpu() - G_cpu[:k]).abs().max())}
        opts = {n: ops.get_option(n) for n in ops.option_names()}
        line = {
            "metric": metric_name(cfg), "value": value, "unit": "poses/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "data": f"synthetic (seeded ellipsoid scenes, {UNIQUE_SCENES} unique per chunk tiled to {chunk}; {fmap_desc}; shipped gru_update weights)",
            "config": {"workload": workload_name(cfg, per_gpu), "name": cfg, "global_batch": world * per_gpu,
                       "chunk": chunk, "chunks_per_step": n_chunks,
                       "parallelism": f"dp{world} (objects sharded, one all-gather of metrics)",
                       "l2": f"resident inputs per chunk ({sum(d[0][k].numel() * 4 for k in keys) / 1e9:.1f} GB) exceed the 126 MB L2; no explicit flush"},
            "e2e": e2e,
            "gpu_launches": args.steps * n_chunks * ops.launch_count(chunk, H, W, N_ITERS, N_LM) + 2,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "encoder": enc_leg, "options": opts,
            "accuracy_vs_gt": {"objects": int(gm.shape[0]), "mean_add_over_diameter": float((gm[:, 0] / gm[:, 15]).mean()),
                               "add_0.1d_recall": float(gm[:, 6].mean()), "adds_0.1d_recall": float(gm[:, 7].mean()),
                               "proj2d_5px_recall": float(gm[:, 12].mean()), "cm5deg5_recall": float(gm[:, 13].mean()),
                               "note": "pose vs ground truth after refinement on synthetic scenes; a sanity number, not parity"},
            "oracle_check": oracle_check,
        }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="cfg1", choices=["cfg1", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--global-batch", type=int, default=None, help="total objects over all ranks (default: 32 per GPU; cfg4: 8 per GPU)")
    ap.add_argument("--chunk", type=int, default=None, help="objects per b200pose_refine_iters call (default: up to 64 at 240x320, 16 at 480x640)")
    ap.add_argument("--sweep", action="store_true", help="cfg4: batch-size sweep 8..512, one JSON line with a `sweep` list")
    ap.add_argument("--fmaps", default="encoder", choices=["encoder", "hash"], help="feature maps: the library's encoder on the synthetic crops, or hash noise")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--host-gather-threads", type=int, default=None,
                    help="e2e leg: -1 = context rows read in place from pinned memory, T >= 0 = texels gathered by T host threads "
                         "(0 = library default); default: measure both, report the faster")
    ap.add_argument("--cpu-objects", type=int, default=8)
    ap.add_argument("--exact-fp32", action="store_true", help="CUDA-core fp32 convolutions instead of the tcgen05 path")
    args = ap.parse_args()
    if args.config == "cfg2":
        args.config = "cfg1"                      # configs[2] is configs[1]'s per-GPU workload on 8 GPUs
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(3, args.warmup)

    from rnnpose_b200 import dist as D
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    rank, local_rank, world = D.init_from_env("nccl")
    assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = bind_numa(local_rank, world)           # before any pinned allocation

    if args.sweep:
        assert args.config == "cfg4", "--sweep is the configs[4] batch-size sweep"
        rows = []
        for gb in SWEEP_BATCHES:
            if gb % world:
                continue
            ln = run_workload(args, "cfg4", gb // world, rank, local_rank, world, dev, numa, want_cpu=False, want_e2e=False)
            if ln:
                rows.append({"global_batch": gb, "value": ln["value"], "ms_per_step": ln["ms_per_step"], "chunk": ln["config"]["chunk"],
                             "roofline_frac": ln["roofline"]["frac"], "conv_tflops": ln["roofline"]["achieved"],
                             "ms_per_conv_launch": ln["roofline"]["ms_per_launch"]})
            torch.cuda.empty_cache()
        line = run_workload(args, "cfg4", max(1, 8 // world) if world <= 8 else 1, rank, local_rank, world, dev, numa)
        if line:
            line["sweep"] = rows
    else:
        per_gpu = (args.global_batch // world) if args.global_batch else CONFIGS[args.config][4]
        assert per_gpu >= 1 and (not args.global_batch or args.global_batch % world == 0)
        line = run_workload(args, args.config, per_gpu, rank, local_rank, world, dev, numa)
    if line:
        print(json.dumps(line), flush=True)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    return line


if __name__ == "__main__":
    main()

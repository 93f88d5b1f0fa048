"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: rank sharding (reference
utils/distributed_utils.py:154-169) and the single metric all-gather (reference tools/train.py:724-741)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rnnpose_b200 import dist as D
from rnnpose_b200 import metrics as M


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, out_q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    D.init_from_env("gloo")
    idx = D.shard_indices(n_total, rank, world)
    # "refine" = a deterministic function of the object index; metrics carry the index in the last column
    g = torch.Generator().manual_seed(0)
    base = torch.randn(64, 7, generator=g)
    local = torch.cat([base[idx], torch.tensor(idx, dtype=torch.float32)[:, None]], dim=1)
    full = D.all_gather_metrics(local, n_total)
    D.barrier()
    mx = D.max_over_ranks(float(rank + 1), torch.device("cpu"))
    if rank == 0:
        out_q.put((full.clone(), mx))
    dist.destroy_process_group()


def test_shard_indices_cover_and_pad():
    assert D.shard_indices(5, 0, 2) == [0, 2, 4] and D.shard_indices(5, 1, 2) == [1, 3, 0]
    assert sorted(D.shard_indices(8, 0, 4) + D.shard_indices(8, 1, 4) + D.shard_indices(8, 2, 4) + D.shard_indices(8, 3, 4)) == list(range(8))


def test_all_gather_metrics_world2_gloo():
    world, n_total = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    [p.start() for p in procs]
    full, mx = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert full.shape == (n_total, 8)
    assert full[:, 7].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]          # object order restored, padding dropped
    g = torch.Generator().manual_seed(0)
    torch.testing.assert_close(full[:, :7], torch.randn(64, 7, generator=g)[:n_total])
    assert mx == 2.0


def test_pose_metrics_against_oracle():
    from oracle import refine_oracle as O
    g = torch.Generator().manual_seed(1)
    xi = torch.randn(4, 6, generator=g) * 0.1
    Tp = O.se3_exp(xi); Tg = O.se3_exp(xi + 0.01)
    pts = torch.randn(50, 3, generator=g) * 0.05
    m = M.pose_metrics(Tp, Tg, pts[None].repeat(4, 1, 1), torch.full((4,), 0.15), torch.arange(4))
    torch.testing.assert_close(m[:, 0], O.add_metric(Tp[:, :3, :3], Tp[:, :3, 3], Tg[:, :3, :3], Tg[:, :3, 3], pts, False))
    torch.testing.assert_close(m[:, 1], O.add_metric(Tp[:, :3, :3], Tp[:, :3, 3], Tg[:, :3, :3], Tg[:, :3, 3], pts, True))
    torch.testing.assert_close(m[:, 2], O.rotation_angle_deg(Tp[:, :3, :3], Tg[:, :3, :3]))
    assert m.shape == (4, len(M.METRIC_NAMES))

"""World-size-2 gloo run on CPU of the only cross-rank logic the path has: batch sharding
(sample b -> rank b mod W), the scalar loss all-reduce and MeanIoU's three counter all-reduces
(reference: misc/metric_util.py:69-73, dataset/__init__.py:54-60)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussianformer_b200.metric import MeanIoU


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _labels(seed, n=4000):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 18, (n,), generator=g), torch.randint(0, 18, (n,), generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    samples = [b for b in range(6) if b % world == rank]          # batch sharding
    metric = MeanIoU(list(range(1, 17)), empty_label=17)
    loss = torch.zeros(1, dtype=torch.float64)
    for b in samples:
        pred, gt = _labels(b)
        metric.after_step(pred, gt, mask=gt != 0)
        loss += (pred == gt).double().mean()
    dist.all_reduce(loss)                                           # the "loss all-reduce"
    miou, iou = metric.after_epoch()
    if rank == 0:
        q.put((float(loss), miou, iou, samples))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss2, miou2, iou2, samples0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert samples0 == [0, 2, 4]
    metric = MeanIoU(list(range(1, 17)), empty_label=17)
    loss = 0.0
    for b in range(6):
        pred, gt = _labels(b)
        metric.after_step(pred, gt, mask=gt != 0)
        loss += float((pred == gt).double().mean())
    miou, iou = metric.after_epoch()
    assert abs(loss - loss2) < 1e-12 and abs(miou - miou2) < 1e-9 and abs(iou - iou2) < 1e-9

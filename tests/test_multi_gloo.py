"""CPU, world_size 2 over gloo: the N>1 path's host logic (frame sharding + sum-form reduce of the accumulator pair)
reproduces single-process stacking.  Each rank stacks its shard with the oracle accumulator (the checker); the
reduce under test is serstacker_b200.multi.reduce_sum_form, the same function bench.py / combine_pipeline call on
NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from serstacker_b200 import multi


def test_shard_frames_partition():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_frames(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multi.shard_frames(4, 2, 2)


def _frames(n, h, w, seed):
    rng = np.random.default_rng(seed)
    fr = rng.random((n, h, w)).astype(np.float32)
    wt = (rng.random((n, h, w)) - 0.15).astype(np.float32)     # some non-positive weights: skipped pixels
    return fr, wt


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import accumulation as oacc
        fr, wt = _frames(n, 24, 40, 7)
        lo, hi = multi.shard_frames(n, rank, world)
        acc = oacc.WeightedAverage()
        for i in range(lo, hi):
            acc.add(fr[i], wt[i])
        A, W = acc.accumulator, acc.weights
        a = torch.from_numpy((A * W).astype(np.float32))       # running mean -> sum form
        w_ = torch.from_numpy(W.astype(np.float32).copy())
        total = multi.reduce_sum_form(a, w_, hi - lo, dst=0)
        if rank == 0:
            Wn = w_.numpy()
            mean = np.where(Wn > 0, a.numpy() / np.where(Wn > 0, Wn, 1), 0).astype(np.float32)
            np.savez(out, mean=mean, W=Wn, total=total)
    finally:
        dist.destroy_process_group()


def test_sharded_stack_equals_single_process(tmp_path):
    from oracle import accumulation as oacc
    n, world = 9, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, port, n, out), nprocs=world, join=True)
    got = np.load(out)
    fr, wt = _frames(n, 24, 40, 7)
    acc = oacc.WeightedAverage()
    for i in range(n):
        acc.add(fr[i], wt[i])
    assert int(got["total"]) == n
    assert np.allclose(got["W"], acc.weights, rtol=1e-6, atol=1e-7)
    m = acc.weights > 0
    assert np.abs(got["mean"][m] - acc.accumulator[m]).max() <= 2e-6

"""GPU, 2 ranks over NCCL (skipped on boxes with one GPU): frames sharded over two processes, each stacking its shard on
its own GPU through ssk_stack, then ssk_stack_reduce (one ncclReduce group through the C ABI) - the combined stack on rank 0
must equal the single-GPU stack of the whole sequence (SURVEY.md section 4 tier 5 / 8e)."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pipeline(max_batch, method=1):
    from serstacker_b200 import api
    ro = api.registration_options(motion_type=3, interpolation=2, ecc=dict(ecc_method=3, ecch_max_level=-1))
    return api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=method, max_batch=max_batch))


def _sequence():
    from serstacker_b200 import synth
    frames, _, _ = synth.make_planet_sequence(480, 270, 11, seed=2, radius=100, sigma_t=4.0, sigma_rot_deg=0.2,
                                              sigma_scale=0.002, blur_range=(0.8, 2.5), dtype="f32")
    return frames


def _worker(rank, world, idfile, out):
    import torch
    from serstacker_b200 import multi
    torch.cuda.set_device(rank)
    if rank == 0:
        uid = multi.nccl_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            assert time.time() - t0 < 120, "rank 0 never published the NCCL id"
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    comm = multi.NcclComm(uid, world, rank)
    frames = _sequence()
    lo, hi = multi.shard_frames(len(frames) - 1, rank, world)
    p = _pipeline(4)
    p.set_reference(frames[0])
    res = p.add_frames(frames[1 + lo:1 + hi])
    local = p.accumulated_frames()
    total = multi.reduce_pipeline(p, comm, dst=0)
    if rank == 0:
        avg, mask = p.compute()
        np.savez(out, avg=avg, mask=mask, total=total, local=local, w=p.accumulator().get_acc_counters())
    else:
        assert total == local          # the other ranks keep their local state
    comm.destroy()


def test_two_rank_nccl_stack_equals_single_gpu_stack(gpu, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out, idfile = str(tmp_path / "rank0.npz"), str(tmp_path / "nccl_id")
    mp.spawn(_worker, args=(2, idfile, out), nprocs=2, join=True)
    got = np.load(out)
    frames = _sequence()
    p = _pipeline(4)
    p.set_reference(frames[0])
    res = p.add_frames(frames[1:])
    avg, mask = p.compute()
    w = p.accumulator().get_acc_counters()
    assert int(got["total"]) == p.accumulated_frames() == sum(r["ok"] for r in res)
    assert np.array_equal(got["mask"], mask)
    m = mask > 0
    rel = float(np.sqrt(((got["avg"][m].astype(np.float64) - avg[m]) ** 2).sum()) / np.sqrt((avg[m].astype(np.float64) ** 2).sum()))
    relw = float(np.abs(got["w"][m] - w[m]).max() / np.abs(w[m]).max())
    print("2-rank NCCL combine vs single GPU: stack rel-L2 = %.3g, weights max rel = %.3g, frames %d" % (rel, relw, int(got["total"])))
    assert rel <= 1e-6
    assert relw <= 1e-6

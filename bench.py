#!/usr/bin/env python
"""
bench.py - frames/s of the stacking hot path (register + warp + stack) on BASELINE.json config #2:
synthetic 1920x1080 mono 32F frames, ECCH pyramid registration (AFFINE, INVERSE_COMPOSITIONAL_LM, translation first),
bicubic remap, sharpness-weighted average.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--chunk C] [--impl ours|reference] [--config 1|2|3|4|5]

The job is the one BASELINE.json names: ONE sequence of K x B frames (B = 8192 by default) stacked into ONE image, the
frames of every step sharded over the N ranks (one process per GPU, multi.shard_frames), i.e. STRONG scaling.  A "step" is
one pass of the hot path over B frames (B / N per rank, in chunks of C = 1024 frames: one ECC launch, four launches of the
fused warp+accumulate kernel).  The timed region runs from the first
frame of the first step to the finished stack on rank 0: per-frame loop on every rank, ssk_stack_reduce (one ncclReduce
group over NVLink through the C ABI) and compute() of the stack into host memory are inside it.

`value` is measured with the frames resident in HBM; `e2e` runs the same job through the C ABI with pinned HOST frames
(H2D of every frame, D2H of the per-frame registration results and of the stack inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/, the same OpenCV kernels through cv2) on the host cores.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1080, 1920
NPIX = H * W
WORKLOAD = "config#2: 1920x1080 mono32F, ECCH(IC-LM, affine, translation-first, full pyramid) + bicubic remap + sharpness-weighted average"


# ------------------------------------------------------------------------------------------------------------
# synthetic frames (torch on the GPU: data generation only, not part of the measured path)
# ------------------------------------------------------------------------------------------------------------
def make_frames_gpu(n, seed, device, radius=400.0, scene_seed=2):
    """Synthetic planetary sequence: one scene (scene_seed: belts, spots) seen through per-frame jitter, defocus and
    noise (seed).  Ranks share the scene - they stack shards of the same sequence - and differ in the jitter."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    rng = np.random.default_rng(scene_seed)
    yy, xx = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32),
                            torch.arange(W, device=device, dtype=torch.float32), indexing="ij")
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    belts = [(rng.uniform(-0.8, 0.8) * radius, rng.uniform(0.03, 0.08) * radius, rng.uniform(-0.25, 0.25)) for _ in range(6)]
    spots = [(rng.uniform(-0.7, 0.7) * radius, rng.uniform(-0.7, 0.7) * radius, rng.uniform(2.0, 6.0) * radius / 150.0,
              rng.uniform(-0.3, 0.3)) for _ in range(40)]
    out = torch.empty((n, H, W), device=device, dtype=torch.float32)
    rng = np.random.default_rng(seed + 7919)
    for i in range(n):
        if i == 0:
            A = np.array([[1.0, 0, 0], [0, 1.0, 0]])
            sig = 0.8
        else:
            tx, ty = np.clip(rng.normal(0.0, 4.0, 2), -10, 10)
            a = math.radians(rng.normal(0.0, 0.2))
            s = rng.normal(1.0, 0.002)
            ca, sa = s * math.cos(a), s * math.sin(a)
            A = np.array([[ca, -sa, cx - ca * cx + sa * cy + tx], [sa, ca, cy - sa * cx - ca * cy + ty]])
            sig = rng.uniform(0.8, 2.5)
        sx = A[0, 0] * xx + A[0, 1] * yy + (A[0, 2] - cx)
        sy = A[1, 0] * xx + A[1, 1] * yy + (A[1, 2] - cy)
        r2 = (sx * sx + sy * sy) / (radius * radius)
        mu = torch.sqrt(torch.clamp(1.0 - r2, 0.0, 1.0))
        img = 0.8 * (1.0 - 0.6 * (1.0 - mu))
        tex = torch.zeros_like(img)
        for (by, bs, ba) in belts:
            tex += ba * torch.exp(-0.5 * ((sy - by) / bs) ** 2)
        for (px, py, ps, pa) in spots:
            tex += pa * torch.exp(-0.5 * (((sx - px) / ps) ** 2 + ((sy - py) / ps) ** 2))
        img = img * (1.0 + tex)
        edge = torch.clamp((1.0 - torch.sqrt(r2)) * radius + 0.5, 0.0, 1.0)
        img = 0.02 + (img - 0.02) * edge
        # defocus blur (separable Gaussian) + noise
        k = int(2 * math.ceil(3 * sig) + 1)
        t = torch.arange(k, device=device, dtype=torch.float32) - k // 2
        kern = torch.exp(-0.5 * (t / sig) ** 2)
        kern = kern / kern.sum()
        im4 = img[None, None]
        im4 = torch.nn.functional.conv2d(torch.nn.functional.pad(im4, (k // 2, k // 2, 0, 0), mode="replicate"), kern.view(1, 1, 1, k))
        im4 = torch.nn.functional.conv2d(torch.nn.functional.pad(im4, (0, 0, k // 2, k // 2), mode="replicate"), kern.view(1, 1, k, 1))
        noise = torch.randn((H, W), generator=g).to(device) * 0.01
        out[i] = torch.clamp(im4[0, 0] + noise, 0.0, 1.0)
    return out


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def oracle_options():
    import cv2
    from oracle import pipeline as opl, transforms as otf, ecc as oecc
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_AFFINE
    so.registration.interpolation = cv2.INTER_CUBIC
    so.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    so.registration.ecc.ecch_max_level = -1
    return so


class CpuStacker:
    """The reference's CPU path (oracle restatement over cv2) as the pipeline runs it: the reference frame is set up ONCE
    (setup_reference_frame), then every frame goes through process_input_sequence's body."""

    def __init__(self, reference, threads):
        import cv2
        from oracle import accumulation as oacc
        from oracle.registration import FrameRegistration
        cv2.setNumThreads(threads)
        self.so = oracle_options()
        self.reg = FrameRegistration(self.so.registration)
        self.reg.setup_reference_frame(reference, None)
        self.acc = oacc.WeightedAverage()

    def add(self, frames):
        from oracle import pipeline as opl
        t0 = time.perf_counter()
        for f in frames:
            opl.process_frame(self.reg, self.acc, self.so, f)
        return time.perf_counter() - t0


def bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs of the NUMA node its GPU hangs off, so that the pinned host frames of the e2e leg
    are allocated next to the GPU's PCIe root (matters at 4-8 ranks per box).  Best effort: silently skipped when
    sysfs does not tell."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(index), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(frames_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused kernel per launch, from the committed ncu capture
    (profiles/fused_traffic.json names the capture and the commit it was taken at); None when the capture does not match."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json")))
        if int(t["frames_per_launch"]) == int(frames_per_launch):
            return float(t["dram_bytes_per_launch"]), "%s @ %s" % (t["source"], t["commit"])
    except Exception:
        pass
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8192, help="frames per step, whole job (sharded over the ranks)")
    ap.add_argument("--chunk", type=int, default=1024, help="frames per launch (max_batch of the pipeline)")
    ap.add_argument("--pool", type=int, default=256, help="distinct synthetic frames resident per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=48, help="frames of the CPU baseline sample")
    ap.add_argument("--verify", type=int, default=8, help="frames per rank of the N-rank vs single-GPU stack check (N > 1)")
    ap.add_argument("--seed", type=int, default=2, help="seed of the synthetic frame pool (rank r uses seed + 1000 r)")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json config: 2 = the headline line (default); 1, 3, 4, 5 = the secondary rows (one GPU, bench_secondary.py)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        # The reference's own CPU implementation of the path on the host cores (the reference cannot be compiled here:
        # DESIGN.md section 2, so this is the oracle port over the same OpenCV kernels).  A step is a bounded sample of the
        # workload; the reference frame is set up once, as the pipeline does.
        if rank != 0:
            return 0
        import torch
        ncores = os.cpu_count() or 1
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        per_step = 6
        n = 1 + per_step * (args.steps + args.warmup)
        pool = make_frames_gpu(min(n, 1 + per_step * 4), 2, dev).cpu().numpy()
        fr = [pool[1 + (i % (len(pool) - 1))] for i in range(n - 1)]
        cpu = CpuStacker(pool[0], ncores)
        for w in range(args.warmup):
            cpu.add(fr[w * per_step:(w + 1) * per_step])
        dt = 0.0
        for s in range(args.steps):
            k0 = (args.warmup + s) * per_step
            dt += cpu.add(fr[k0:k0 + per_step])
        fps = args.steps * per_step / dt
        print(json.dumps({
            "impl": "reference", "metric": "frames/sec register+warp+stack 1080p", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": per_step},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ncores, "kind": "port",
                             "sample": "%d frames per step x %d steps of the same workload, reference frame set up once (oracle/ = reference restated over cv2 %s, cv2 threads = %d)" % (
                                 per_step, args.steps, __import__("cv2").__version__, ncores)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    if args.config != 2:
        if rank != 0:
            return 0
        import bench_secondary
        return bench_secondary.run(args, peaks, ClockSampler)

    import torch
    import torch.distributed as dist
    from serstacker_b200 import api, capi, multi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch                                   # frames per step, whole job
    CH = max(1, args.chunk)
    lo, hi = multi.shard_frames(B, rank, world)      # this rank's frames of every step
    mine = hi - lo
    pool_n = max(args.pool, min(mine, CH, 1024) + 1)     # a launch never sees a frame twice
    pool = make_frames_gpu(pool_n, args.seed + 1000 * rank, dev)       # frame 0 = unjittered reference scene
    ref = make_frames_gpu(1, 2, dev)[0] if rank != 0 else pool[0]      # every rank registers against the same reference frame
    if rank != 0:
        pool[0] = ref

    ro = api.registration_options(motion_type=capi.MOTION_AFFINE, interpolation=capi.INTER_CUBIC,
                                  ecc=dict(ecc_method=capi.ECC_INVERSE_COMPOSITIONAL_LM, ecch_max_level=-1))
    so = api.stack_options(registration=ro, accumulation_method=capi.STACK_WEIGHTED_AVERAGE, max_batch=CH)
    pipe = api.c_image_stacking_pipeline(so)
    pipe.set_reference(capi.device_mat(ref.data_ptr(), H, W, np.float32))
    stream = torch.cuda.ExternalStream(pipe.stream(), device=dev)
    comm = None
    if world > 1:
        comm = multi.create_comm_from_torch()        # ncclComm_t owned through the C ABI (what a C++ host would pass)
        multi.reduce_pipeline(pipe, comm, 0)         # warms the communicator; the accumulator is still empty
        pipe.reset()

    def dev_step(step):
        g0 = step * B + lo
        return [capi.device_mat(pool[1 + ((g0 + i) % (pool_n - 1))].data_ptr(), H, W, np.float32) for i in range(mine)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    avg_host = np.empty((H, W), np.float32)
    mask_host = np.empty((H, W), np.uint8)

    def finish(p):
        """End of the job: partial stacks to rank 0 (one ncclReduce group through ssk_stack_reduce), rank 0 reads the stack."""
        if world > 1:
            multi.reduce_pipeline(p, comm, 0)
        if rank == 0:
            capi.check(capi.lib.ssk_stack_compute(p._h, C.byref(capi.mat(avg_host)), C.byref(capi.mat(mask_host))))

    # ---------------- device-resident throughput --------------------------------------------------------
    for s in range(args.warmup):
        pipe.add_frames_async(dev_step(s))
    finish(pipe)
    pipe.sync()
    fused_ms = pipe.stage_times()[3]                 # last launch of the warm-up: the fused warp+accumulate kernel alone
    frames_last_launch = mine - ((mine - 1) // CH) * CH
    pipe.reset()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = capi.lib.ssk_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    keep = []
    for s in range(args.steps):
        keep.append(pipe.add_frames_async(dev_step(args.warmup + s)))
    finish(pipe)
    e1.record(stream)
    pipe.sync()
    barrier()
    launches = capi.lib.ssk_kernel_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    stage = pipe.stage_times()                      # last launch: prep, weights, ECC, warp+accumulate (ms)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    n_l = torch.tensor([launches], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_l, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    accumulated = pipe.accumulated_frames()          # rank 0: frames of all ranks (the reduce carries the count)
    value = args.steps * B / (ms * 1e-3)
    keep = None

    # ---------------- N ranks against one GPU: the combined stack equals the single-GPU stack ------------
    combine_rel_l2 = None
    if world > 1 and args.verify > 0:
        V = min(args.verify, pool_n - 1)
        pv = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=capi.STACK_WEIGHTED_AVERAGE, max_batch=V))
        pv.set_reference(capi.device_mat(ref.data_ptr(), H, W, np.float32))
        pv.add_frames_async([capi.device_mat(pool[1 + i].data_ptr(), H, W, np.float32) for i in range(V)])
        multi.reduce_pipeline(pv, comm, 0)
        if rank == 0:
            got, gmask = pv.compute()
            p1 = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=capi.STACK_WEIGHTED_AVERAGE, max_batch=V))
            p1.set_reference(capi.device_mat(ref.data_ptr(), H, W, np.float32))
            for r in range(world):       # every rank's frames come from its seed: rank 0 regenerates them
                fr = pool if r == 0 else make_frames_gpu(V + 1, args.seed + 1000 * r, dev)
                p1.add_frames_async([capi.device_mat(fr[1 + i].data_ptr(), H, W, np.float32) for i in range(V)])
                p1.sync()
            want, wmask = p1.compute()
            m = (gmask > 0) & (wmask > 0)
            combine_rel_l2 = float(np.sqrt(((got[m].astype(np.float64) - want[m]) ** 2).sum()) / np.sqrt((want[m].astype(np.float64) ** 2).sum()))
            assert pv.accumulated_frames() == p1.accumulated_frames(), (pv.accumulated_frames(), p1.accumulated_frames())
            assert np.array_equal(gmask, wmask) and combine_rel_l2 <= 1e-6, combine_rel_l2
        barrier()

    # ---------------- end to end through the C ABI with host buffers ------------------------------------
    n_host = max(1, min(mine, 128, pool_n - 1))
    host = torch.empty((n_host, H, W), dtype=torch.float32).pin_memory()
    host.copy_(pool[1:1 + n_host])
    host_np = host.numpy()
    pipe2 = api.c_image_stacking_pipeline(so)
    pipe2.set_reference(ref.cpu().numpy())
    e2e_steps = max(1, min(args.steps, 4))

    def host_chunks(step):
        g0 = step * B + lo
        fr = [host_np[(g0 + i) % n_host] for i in range(mine)]
        return [fr[i:i + CH] for i in range(0, mine, CH)]

    for ch in host_chunks(0)[:2]:
        pipe2.add_frames(ch)
    pipe2.reset()
    barrier()
    # streaming use of the public API: a chunk is submitted (H2D of its frames from pinned memory + processing enqueued),
    # then the per-frame registration results of the previous chunk are read back (D2H) while this one runs
    t0 = time.perf_counter()
    ok_frames = 0
    prev = None
    for s in range(e2e_steps):
        for ch in host_chunks(s):
            ticket = pipe2.submit(ch)
            if prev is not None:
                ok_frames += sum(1 for r in pipe2.wait(prev) if r["ok"])
            prev = ticket
    if prev is not None:
        ok_frames += sum(1 for r in pipe2.wait(prev) if r["ok"])
    finish(pipe2)
    pipe2.sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    e2e = e2e_steps * B / e2e_s
    # what limits it: the same bytes by bare cudaMemcpyAsync from the same pinned buffer, all ranks at once
    scratch = torch.empty((n_host, H, W), dtype=torch.float32, device=dev)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    reps = max(1, (e2e_steps * mine) // n_host)
    for _ in range(reps):
        scratch.copy_(host, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    bare = torch.tensor([reps * n_host * NPIX * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(bare, op=dist.ReduceOp.MIN)
    h2d_bare = float(bare.item())
    h2d_e2e = e2e_steps * mine * NPIX * 4 / e2e_s / 1e9
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # ---------------- roofline of the fused warp+accumulate kernel --------------------------------------
    peak, peak_src = peaks()
    # The fused stage of a chunk of frames_last_launch frames is ceil(frames / 256) launches of k_fused_tma (KPLAN = 256 frames per
    # launch, csrc/ssk_fused_impl.cuh) plus k_fill_jobs; the stage time of the LAST chunk of the timed region (CUDA events on the
    # pipeline's stream) divided by that count is the kernel's average launch duration
    n_kl = (frames_last_launch + 255) // 256
    FL = frames_last_launch / n_kl
    fused_ms = stage[3] / n_kl
    t_k = fused_ms * 1e-3
    bytes_kernel = NPIX * (4 + 4) * FL + NPIX * 16   # frame + weight map read per frame; mean + weight RMW once per launch
    bytes_survey = NPIX * 24 * FL                    # N*(4 + 8 + 8 + 4) per frame (un-batched RMW)
    # roofline.achieved follows the contract: SURVEY section 8(d)'s per-frame figure x the frames of one launch / launch time.
    # The kernel keeps the accumulator tile on chip across the batch, so the bytes it really has to move are fewer
    # (frame + weight map once per frame, accumulators once per batch): reported beside it as *_resident_acc.
    achieved = bytes_survey / t_k / 1e9
    achieved_resident = bytes_kernel / t_k / 1e9
    traffic, traffic_src = measured_traffic(FL)

    # ---------------- CPU baseline (rank 0, N=1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        ncores = os.cpu_count() or 1
        sample = pool[:1 + args.cpu_sample].cpu().numpy()
        st = CpuStacker(sample[0], ncores)
        st.add(list(sample[1:4]))
        dt_cpu = st.add(list(sample[1:]))
        cpu = {"value": args.cpu_sample / dt_cpu, "unit": "frames/s", "cores": ncores, "kind": "port",
               "sample": "%d frames of the same workload in %.1f s, reference frame set up once (oracle/: reference restated over cv2 %s, cv2 threads = %d)" % (
                   args.cpu_sample, dt_cpu, __import__("cv2").__version__, ncores)}

    if rank == 0:
        line = {
            "metric": "frames/sec register+warp+stack 1080p", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "job": "%d frames stacked into one image, every step's %d frames sharded over %d rank(s); timed through the reduce to rank 0 and compute() into host memory" % (
                           args.steps * B, B, world),
                       "frames_per_step": B, "frames_per_step_per_gpu": mine, "frames_per_launch": CH, "resident_pool_frames": pool_n,
                       "l2_policy": "inputs larger than L2: %d distinct frames (%.1f GB) per GPU cycled, %.0f MB touched per step per GPU" % (
                           pool_n, pool_n * NPIX * 4 / 1e9, mine * NPIX * 4 / 1e6),
                       "accumulated_frames": accumulated},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * NPIX * 4,
                    "d2h_bytes_per_step": B * (C.sizeof(capi.ssk_transform) + C.sizeof(capi.ssk_ecc_status)) + (NPIX * 5) // max(e2e_steps, 1),
                    "steps": e2e_steps, "registered_frames_rank0": ok_frames,
                    "h2d_gbs_per_gpu": h2d_e2e, "h2d_gbs_per_gpu_bare_memcpy": h2d_bare,
                    "limiter": "host-to-device copy of the frames: the job moves %.1f GB/s per GPU against %.1f GB/s of bare cudaMemcpyAsync from the same pinned buffer with all %d rank(s) copying at once" % (
                        h2d_e2e, h2d_bare, world)},
            "gpu_launches": int(n_l.item()),
            "roofline": {"kernel": "k_fused_tma (fused bicubic warp + eroded mask + weight warp + running weighted mean; one launch per 256 frames over all tiles)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launch_ms": fused_ms, "frames_per_launch": FL,
                         "algorithmic_bytes_per_launch": bytes_survey,
                         "algorithmic_bytes_per_frame": NPIX * 24,
                         "achieved_resident_acc": achieved_resident, "frac_resident_acc": achieved_resident / peak,
                         "bytes_per_launch_resident_acc": bytes_kernel},
            "stage_ms_per_launch": {"prep": stage[0], "weights": stage[1], "ecc": stage[2], "warp_accumulate": stage[3], "frames": frames_last_launch},
            "combine_rel_l2": combine_rel_l2,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

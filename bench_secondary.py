"""Secondary rows of the measurement plan: `python bench.py --config 1|3|4|5`.

One GPU, frames resident in HBM, the same JSON shape as the headline line of bench.py: frames/s of the config's per-frame
path, the HBM-roofline fraction of its accumulate stage against section 8(d)'s algorithmic bytes per frame, and the oracle
(CPU restatement) timed on a bounded sample of the same workload.  Config #2 is the headline line of bench.py itself.

  #1  640x480 mono16, translation ECC (forward-additive, single level) + LINEAR/REFLECT101 warp + average      ssk_stack
  #3  4096x3000 RGGB16, debayer_nn2 -> translation ECC -> Bayer-average accumulation of the raw samples          ssk_stack
  #4  2048x2048 mono32F Jupiter: derotation remap + lpg weights + blend accumulation (pose given)               ssk_jdr_derotate_and_add
  #5  2448x2048 RGB32F focus stack: lpg(k=6, p=2, 0, 0) -> GaussianBlur(1) -> weighted add, no registration     ssk_lpg / ssk_gaussian_blur / ssk_acc_add
"""
import ctypes as C
import json
import math
import os
import time

import numpy as np


def _planet_u16(n, w, h, seed):
    from serstacker_b200 import synth
    frames, _, bpp = synth.make_planet_sequence(w, h, n, seed=seed, radius=150, sigma_t=3.0, dtype="u16")
    return frames, bpp


def _jovian(size, center, axes, lon, lat, pa, seed):
    """Synthetic textured oblate planet at a pose (longitude, tilt, position angle) - data generation only."""
    import cv2
    w, h = size
    cl, sl, ct, st, cp, sp = math.cos(lon), math.sin(lon), math.cos(lat), math.sin(lat), math.cos(pa), math.sin(pa)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    xs, ys = xx - center[0], yy - center[1]
    xr, yr = cp * xs + sp * ys, -sp * xs + cp * ys
    A, B = axes[0], axes[1]
    r2 = (xr / A) ** 2 + (yr / B) ** 2
    z = np.sqrt(np.clip(1 - r2, 0, 1))
    latp = np.arcsin(np.clip(yr / B * ct + z * st, -1, 1))
    lonp = np.arctan2(xr / A, z) + lon
    tex = 0.5 + 0.2 * np.sin(7 * latp) + 0.15 * np.sin(9 * lonp + 3 * latp) + 0.1 * np.cos(23 * lonp) * np.cos(11 * latp)
    rng = np.random.default_rng(seed)
    img = np.where(r2 < 1, tex * (0.4 + 0.6 * z), 0.02) + rng.normal(0, 0.003, (h, w))
    return cv2.GaussianBlur(img.astype(np.float32), (0, 0), 1.0)


def _rot(lon, lat, pa):
    """XYZscreen = R XYZplanet for a pose (rotation about the polar axis, tilt towards the viewer, position angle)."""
    cl, sl, ct, st, cp, sp = math.cos(lon), math.sin(lon), math.cos(lat), math.sin(lat), math.cos(pa), math.sin(pa)
    Ry = np.array([[cl, 0, sl], [0, 1, 0], [-sl, 0, cl]])
    Rx = np.array([[1, 0, 0], [0, ct, -st], [0, st, ct]])
    Rz = np.array([[cp, -sp, 0], [sp, cp, 0], [0, 0, 1]])
    return Rz @ Rx @ Ry


def _focus_frames(n, w, h, seed):
    import cv2
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    scene = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        tex = cv2.resize(rng.random((h // 8, w // 8)).astype(np.float32), (w, h), interpolation=cv2.INTER_CUBIC)
        fine = rng.random((h, w)).astype(np.float32)
        scene[..., c] = np.clip(0.35 + 0.4 * (tex - 0.5) + (cv2.GaussianBlur(fine, (0, 0), 1.0) - 0.5), 0, 1)
    depth = (xx / w + 0.5 * yy / h) / 1.5
    b1, b2 = cv2.GaussianBlur(scene, (0, 0), 1.5), cv2.GaussianBlur(scene, (0, 0), 4.0)
    out = []
    for i in range(n):
        d = np.clip(np.abs(depth - (i + 0.5) / n) * 3.0, 0, 1)[..., None]
        f = np.where(d < 0.5, scene * (1 - 2 * d) + b1 * (2 * d), b1 * (2 - 2 * d) + b2 * (2 * d - 1)).astype(np.float32)
        out.append(np.clip(f + rng.normal(0, 0.002, f.shape).astype(np.float32), 0, 1))
    return out


def run(args, peaks, ClockSampler):
    import torch
    from serstacker_b200 import api, capi
    cfg = args.config
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peak, peak_src = peaks()
    sampler = ClockSampler(0)
    launches0 = None
    cpu = None
    ncores = os.cpu_count() or 1

    if cfg in (1, 3):
        if cfg == 1:
            W, H, B, POOL = 640, 480, 2000, 64        # 20 steps = 40 000 frames: long enough for steady clocks
            frames, bpp = _planet_u16(POOL, W, H, seed=1)
            workload = "config#1: 640x480 mono16, ECC(forward-additive, translation, single level) + LINEAR/REFLECT101 remap + average"
            ro = api.registration_options(motion_type=capi.MOTION_TRANSLATION, interpolation=capi.INTER_LINEAR,
                                          ecc=dict(ecc_method=capi.ECC_FORWARD_ADDITIVE, ecch_max_level=0))
            so = api.stack_options(registration=ro, accumulation_method=capi.STACK_AVERAGE, max_batch=min(args.chunk, 250))
            bytes_frame = W * H * (2 + 8 + 8)                    # N (s_in + 8C + 8)
            kname = "fused LINEAR warp + eroded mask + running mean of 16-bit frames (k_fused_staged / k_fused_generic)"
        else:
            from serstacker_b200 import synth
            W, H, B, POOL = 4096, 3000, 64, 8
            frames, _, bpp = synth.make_bayer_sequence(W, H, POOL, seed=3)
            workload = "config#3: 4096x3000 RGGB16, debayer_nn2 -> ECC(IC-LM, translation, full pyramid) at scale 0.5 -> Bayer-average accumulation"
            ro = api.registration_options(motion_type=capi.MOTION_TRANSLATION, interpolation=capi.INTER_LINEAR,
                                          ecc=dict(ecc_method=capi.ECC_INVERSE_COMPOSITIONAL_LM, ecch_max_level=-1))
            so = api.stack_options(registration=ro, accumulation_method=capi.STACK_BAYER_AVERAGE, bayer_colorid=capi.COLORID_BAYER_RGGB,
                                   max_batch=min(args.chunk, int(os.environ.get("SSK_C3_CHUNK", "32"))))
            bytes_frame = W * H * (2 + 24 + 24)                  # N (s_in + 2*12 + 2*12)
            kname = "Bayer gather of the raw samples through the analytic map into acc / cntr (k_bayer_warp_accumulate)"
        CH = so.max_batch
        pool = [torch.from_numpy(f.view(np.int16)).to(dev) for f in frames]     # 16-bit samples (torch has no uint16 arithmetic: raw bytes only)
        pipe = api.c_image_stacking_pipeline(so)
        pipe.set_reference(frames[0], bpp=bpp)
        stream = torch.cuda.ExternalStream(pipe.stream(), device=dev)

        def step_mats(s):
            return [capi.device_mat(pool[1 + ((s * B + i) % (POOL - 1))].data_ptr(), H, W, np.uint16) for i in range(B)]

        for s in range(args.warmup):
            pipe.add_frames_async(step_mats(s))
        pipe.sync()
        pipe.reset()
        torch.cuda.synchronize()
        sampler.start()
        launches0 = capi.lib.ssk_kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        keep = [pipe.add_frames_async(step_mats(args.warmup + s)) for s in range(args.steps)]
        out = pipe.compute()
        e1.record(stream)
        pipe.sync()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        stage = pipe.stage_times()
        FL = B - ((B - 1) // CH) * CH
        t_k = stage[3] * 1e-3
        accumulated = pipe.accumulated_frames()
        # e2e: the same job from pinned host frames through the public call
        host = torch.empty((min(B, POOL - 1), H, W), dtype=torch.int16).pin_memory()
        for i in range(host.shape[0]):
            host[i].copy_(torch.from_numpy(frames[1 + i].view(np.int16)))
        hnp = host.numpy().view(np.uint16)
        pipe2 = api.c_image_stacking_pipeline(so)
        pipe2.set_reference(frames[0], bpp=bpp)
        hf = [hnp[i % hnp.shape[0]] for i in range(B)]
        pipe2.add_frames(hf[:CH])
        pipe2.reset()
        ocn = 3 if cfg == 3 else 1
        avg_host = torch.empty((H, W) if ocn == 1 else (H, W, ocn), dtype=torch.float32).pin_memory().numpy()     # the stack is read into pinned memory
        mask_host = torch.empty((H, W), dtype=torch.uint8).pin_memory().numpy()
        # streaming use of the public API, as bench.py's headline e2e leg: a chunk is submitted (H2D from pinned memory + processing
        # enqueued), the per-frame results of the previous chunk are read back while it runs; E2E_STEPS passes over the B frames, one
        # read-out of the stack at the end
        E2E_STEPS = 4
        t0 = time.perf_counter()
        prev = None
        for s2 in range(E2E_STEPS):
            for i in range(0, B, CH):
                ticket = pipe2.submit(hf[i:i + CH])
                if prev is not None:
                    pipe2.wait(prev)
                prev = ticket
        pipe2.wait(prev)
        capi.check(capi.lib.ssk_stack_compute(pipe2._h, C.byref(capi.mat(avg_host)), C.byref(capi.mat(mask_host))))
        e2e_s = (time.perf_counter() - t0) / E2E_STEPS
        e2e = {"value": B / e2e_s, "unit": "frames/s", "steps": E2E_STEPS, "h2d_gbs": B * W * H * 2 / e2e_s / 1e9, "h2d_bytes_per_step": B * W * H * 2,
               "d2h_bytes_per_step": B * (C.sizeof(capi.ssk_transform) + C.sizeof(capi.ssk_ecc_status)) + W * H * (13 if cfg == 3 else 5)}
        stage_ms = {"prep": stage[0], "weights": stage[1], "ecc": stage[2], "warp_accumulate": stage[3], "frames": FL}
        # CPU baseline
        from oracle import pipeline as opl, transforms as otf, ecc as oecc
        import cv2
        cv2.setNumThreads(ncores)
        so_o = opl.StackingOptions(accumulation_method=opl.ACC_BAYER_AVERAGE if cfg == 3 else opl.ACC_AVERAGE)
        so_o.registration.motion_type = otf.IMAGE_MOTION_TRANSLATION
        so_o.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM if cfg == 3 else oecc.ECC_ALIGN_FORWARD_ADDITIVE
        so_o.registration.ecc.ecch_max_level = -1 if cfg == 3 else 0
        ns = 3 if cfg == 3 else min(48, POOL - 1)
        t0 = time.perf_counter()
        if cfg == 3:
            opl.run_bayer_stacking(frames[1:1 + ns], bpp, so_o, 8, reference=frames[0])
        else:
            opl.run_stacking([opl.to_float_frame(f, bpp) for f in frames[1:1 + ns]], so_o, reference=opl.to_float_frame(frames[0], bpp))
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": ns / dt_cpu, "unit": "frames/s", "cores": ncores, "kind": "port",
               "sample": "%d frames of the same workload in %.1f s incl. the reference set-up (oracle/ over cv2 %s, cv2 threads = %d)" % (ns, dt_cpu, cv2.__version__, ncores)}
        value = args.steps * B / (ms * 1e-3)
        frames_total = args.steps * B
    else:
        import cv2
        if cfg == 4:
            W = H = 2048
            POOL, B = 6, 100
            center, A = (1021.4, 1030.8), 700.0
            axes = (A, A * 0.93512560845968779724, A)
            target = (0.3, math.radians(3.0), math.radians(15.0))
            period, wts = 35740.632, 190.0
            times = [(-150.0 + 60.0 * i) for i in range(POOL)]
            poses = [(target[0] - 2 * math.pi * t / period, target[1], target[2]) for t in times]
            frames = [_jovian((W, H), center, axes, p[0], p[1], p[2], seed=i) for i, p in enumerate(poses)]
            workload = "config#4: 2048x2048 mono32F Jupiter sequence, per-frame derotation remap + lpg(k=2,p=2,dscale=2,uscale=6) weights + GaussianBlur + blend accumulation (pose given)"
            bytes_frame = W * H * (4 + 8 + 8 + 4)
            kname = "whole per-frame chain of ssk_jdr_derotate_and_add (derotation map, lpg, weight rules, GaussianBlur, TRANSPARENT remap, weighted add): time of the chain, not of one kernel"
            Rt = _rot(*target)
            half = int(A) + 24
            cbox = (max(0, int(center[0]) - half), max(0, int(center[1]) - half), min(W, 2 * half), min(H, 2 * half))
            acc = api.c_weigthed_average()
            dfr = [torch.from_numpy(f).to(dev) for f in frames]
            d = lambda v, n: (C.c_double * n)(*[float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)])

            # per-frame arguments marshalled once (the measured loop is the library, not ctypes object construction)
            jargs = []
            for k in range(POOL):
                jargs.append((C.byref(capi.device_mat(dfr[k].data_ptr(), H, W, np.float32)), d(center, 2), d(axes, 3), d(_rot(*poses[k]), 9), d(Rt, 9),
                              math.degrees(target[2]), (C.c_int * 4)(*cbox), 1.0 / (1.0 + abs(times[k]) / wts), int(times[k] == 0)))

            def one_on(i, m):
                _, c2, a3, rc, rt, ang, cb, ws, ism = jargs[i % POOL]
                capi.check(capi.lib.ssk_jdr_derotate_and_add(acc._h, m, None, c2, a3, rc, rt, ang, cb, ws, ism, 1, 2.0, 2.0, 2, 6))

            def one(i):
                one_on(i, jargs[i % POOL][0])
            h2d = W * H * 4
        else:
            W, H, POOL, B = 2448, 2048, 4, 100
            frames = _focus_frames(POOL, W, H, seed=5)
            workload = "config#5: 2448x2048 RGB32F focus stack, lpg(k=6,p=2,dscale=0,uscale=0) -> GaussianBlur(1) -> weighted average, no registration"
            bytes_frame = W * H * (12 + 24 + 8 + 4)
            kname = "whole per-frame chain ssk_lpg -> ssk_gaussian_blur -> ssk_acc_add: time of the chain, not of one kernel"
            acc = api.c_weigthed_average()
            dfr = [torch.from_numpy(f).to(dev) for f in frames]
            wmap = torch.empty((H, W), dtype=torch.float32, device=dev)
            wblur = torch.empty((H, W), dtype=torch.float32, device=dev)

            fm = [capi.device_mat(t.data_ptr(), H, W, np.float32, cn=3) for t in dfr]
            mw, mb = capi.device_mat(wmap.data_ptr(), H, W, np.float32), capi.device_mat(wblur.data_ptr(), H, W, np.float32)
            rf, rw, rb = [C.byref(m) for m in fm], C.byref(mw), C.byref(mb)

            def one_on(i, m):
                capi.check(capi.lib.ssk_lpg(m, 6.0, 2.0, 0, 0, rw))
                capi.check(capi.lib.ssk_gaussian_blur(rw, 1.0, 1.0, rb))
                capi.check(capi.lib.ssk_acc_add(acc._h, m, rb, 0))

            def one(i):
                one_on(i, rf[i % POOL])

            h2d = W * H * 12
        # the frames are device-resident: the calls are stream-ordered (ssk_set_stream_ordered), the host enqueues ahead and
        # compute() at the end of the step waits for the chain
        api.set_stream_ordered(True)
        ocn = 1 if cfg == 4 else 3
        avg_host = torch.empty((H, W) if ocn == 1 else (H, W, ocn), dtype=torch.float32).pin_memory().numpy()     # the stack is read into pinned memory
        mask_host = torch.empty((H, W), dtype=torch.uint8).pin_memory().numpy()
        for i in range(args.warmup * 2):
            one(i)
        acc.clear()
        torch.cuda.synchronize()
        sampler.start()
        launches0 = capi.lib.ssk_kernel_launch_count()
        t0 = time.perf_counter()
        for s in range(args.steps):
            for i in range(B):
                one(s * B + i)
        capi.check(capi.lib.ssk_acc_compute(acc._h, C.byref(capi.mat(avg_host)), C.byref(capi.mat(mask_host)), 1.0))   # waits for the chain
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        accumulated = acc.accumulated_frames()
        value = args.steps * B / (ms * 1e-3)
        frames_total = args.steps * B
        t_k = ms * 1e-3 / frames_total
        FL = 1
        stage_ms = None
        # e2e: pinned host frames -> device (H2D inside the timed region, double-buffered against the previous frame's chain)
        # -> the same stream-ordered calls -> compute() of the stack into host memory
        acc.clear()
        pinned = [torch.from_numpy(f).pin_memory() for f in frames]
        dbuf = [torch.empty_like(dfr[0]) for _ in range(2)]
        dmat = [C.byref(capi.device_mat(t.data_ptr(), H, W, np.float32, cn=(1 if cfg == 4 else 3))) for t in dbuf]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(B):
            dbuf[i & 1].copy_(pinned[i % POOL], non_blocking=True)    # overlaps the chain of frame i - 1
            torch.cuda.current_stream().synchronize()
            api.device_synchronize()                                  # frame i - 1 done: its buffer is free for frame i + 1
            one_on(i, dmat[i & 1])
        capi.check(capi.lib.ssk_acc_compute(acc._h, C.byref(capi.mat(avg_host)), C.byref(capi.mat(mask_host)), 1.0))
        e2e_s = time.perf_counter() - t0
        api.set_stream_ordered(False)
        e2e = {"value": B / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": B * h2d, "d2h_bytes_per_step": W * H * (4 * (1 if cfg == 4 else 3) + 1)}
        # CPU baseline
        cv2.setNumThreads(ncores)
        from oracle import accumulation as oacc
        o = oacc.WeightedAverage()
        ns = 3
        t0 = time.perf_counter()
        if cfg == 4:
            from oracle import derotation as od
            for i in range(ns):
                od.jdr_derotate_and_add(o, frames[i], None, (W, H), center, axes, target, poses[i][0] - target[0], 1.0 / (1.0 + abs(times[i]) / wts),
                                        is_master=(times[i] == 0), lpg_opts=dict(k=2.0, p=2.0, dscale=2, uscale=6))
        else:
            from oracle import weights as ow
            for i in range(ns):
                o.add(frames[i], cv2.GaussianBlur(ow.lpg(frames[i], k=6.0, p=2.0, dscale=0, uscale=0), (0, 0), 1, None, 1, cv2.BORDER_REPLICATE))
        dt_cpu = time.perf_counter() - t0
        cpu = {"value": ns / dt_cpu, "unit": "frames/s", "cores": ncores, "kind": "port",
               "sample": "%d frames of the same workload in %.1f s (oracle/ over cv2 %s, cv2 threads = %d)" % (ns, dt_cpu, cv2.__version__, ncores)}

    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = capi.lib.ssk_kernel_launch_count() - launches0
    achieved = bytes_frame * FL / t_k / 1e9
    resident = None
    if cfg in (1, 3):
        # the accumulators stay on chip for the frames of a launch: what the kernel has to move is the frames once and the
        # accumulators (read + write) once per launch; section 8(d)'s per-frame figure counts their read-modify-write every frame,
        # so "frac" can exceed 1 - the stricter figure is reported next to it
        acc_rw = W * H * (48 if cfg == 3 else 16)
        need = FL * W * H * 2 + acc_rw
        resident = {"bytes_per_launch_resident_acc": need, "achieved_resident_acc": need / t_k / 1e9, "frac_resident_acc": need / t_k / 1e9 / peak}
    line = {
        "metric": "frames/sec register+warp+stack (config #%d)" % cfg, "value": value, "unit": "frames/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "frames_per_step": B, "resident_pool_frames": POOL, "accumulated_frames": accumulated,
                   "l2_policy": "inputs larger than L2: %d distinct frames (%.0f MB) cycled" % (POOL, POOL * bytes_frame / 1e6) if POOL * W * H * 2 > 126e6 else
                                "pool of %d distinct frames (%.0f MB of input); the accumulate stage streams %.1f MB per frame" % (POOL, POOL * W * H * 2 / 1e6, bytes_frame / 1e6)},
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "launch_ms": t_k * 1e3, "frames_per_launch": FL, "algorithmic_bytes_per_frame": bytes_frame,
                     **(resident or {})},
        "stage_ms_per_launch": stage_ms, "cpu_baseline": cpu, "clocks": sampler.summary(),
    }
    print(json.dumps(line))
    return 0

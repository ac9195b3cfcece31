"""Stage times of the stacking loop with the c_eccflow stage enabled (config #2 geometry), on the GPU box.
usage: python tools/flow_bench.py [frames per launch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from serstacker_b200 import api, capi, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H = 1920, 1080
frames, _, _ = synth.make_planet_sequence(W, H, 8, seed=2, radius=400, sigma_t=4.0, sigma_rot_deg=0.2, sigma_scale=0.002,
                                          blur_range=(0.8, 2.5), dtype="f32")
dev = [torch.from_numpy(f).cuda() for f in frames]
for flow in (0, 1):
    ro = api.registration_options(motion_type=3, interpolation=2, enable_eccflow_registration=flow, ecc=dict(ecc_method=3, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1, max_batch=B))
    p.set_reference(frames[0])
    mats = [capi.device_mat(dev[i % len(dev)].data_ptr(), H, W, np.float32) for i in range(B)]
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        p.add_frames(mats, want_results=False)
        p.sync(); torch.cuda.synchronize(); dt = time.time() - t0
        st = p.stage_times()
    print("eccflow=%d: %d frames per launch, wall %.2f ms, stages [prep, weights, ecc(+flow), warp+accumulate] = %s ms -> %.1f frames/s"
          % (flow, B, dt * 1e3, ["%.2f" % x for x in st], B / dt))
    del p

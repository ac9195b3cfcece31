"""Profiling driver (not a test, not a bench): a few device-resident steps of bench.py's workload (config #2, 128 frames per
step) with nothing else around them, for `ncu`.  Usage: python tools/prof_step.py [steps] [batch]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from serstacker_b200 import api, capi

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
pool = bench.make_frames_gpu(B + 1, 2, dev)
ro = api.registration_options(motion_type=capi.MOTION_AFFINE, interpolation=capi.INTER_CUBIC,
                              ecc=dict(ecc_method=capi.ECC_INVERSE_COMPOSITIONAL_LM, ecch_max_level=-1))
pipe = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=capi.STACK_WEIGHTED_AVERAGE, max_batch=B))
pipe.set_reference(capi.device_mat(pool[0].data_ptr(), bench.H, bench.W, np.float32))
frames = [capi.device_mat(pool[j].data_ptr(), bench.H, bench.W, np.float32) for j in range(1, B + 1)]
for s in range(steps):
    pipe.add_frames_async(frames)
    pipe.sync()
    print("step", s, "stage ms", [round(v, 3) for v in pipe.stage_times()], flush=True)
print("accumulated", pipe.accumulated_frames())

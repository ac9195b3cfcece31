"""Stage times of the stacking loop on 3-channel frames (config #2 shape, RGB): python tools/rgb_step.py [batch] [dtype f32|u16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from serstacker_b200 import api, capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
mono = bench.make_frames_gpu(B + 1, 2, dev)                       # (B + 1, H, W)
gains = torch.tensor([0.9, 1.0, 0.8], device=dev)
pool = (mono[..., None] * gains).contiguous()                     # BGR frames of the same scene
ro = api.registration_options(motion_type=capi.MOTION_AFFINE, interpolation=capi.INTER_CUBIC,
                              ecc=dict(ecc_method=capi.ECC_INVERSE_COMPOSITIONAL_LM, ecch_max_level=-1))
pipe = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=capi.STACK_WEIGHTED_AVERAGE, max_batch=B))
pipe.set_reference(capi.device_mat(pool[0].data_ptr(), bench.H, bench.W, np.float32, cn=3))
frames = [capi.device_mat(pool[j].data_ptr(), bench.H, bench.W, np.float32, cn=3) for j in range(1, B + 1)]
for s in range(3):
    pipe.add_frames_async(frames)
    pipe.sync()
    print("step", s, "stage ms", [round(v, 3) for v in pipe.stage_times()], flush=True)
print("accumulated", pipe.accumulated_frames())

"""Debug helper (not a test): ECC iteration statistics of the bench workload (config #2) on the GPU."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from serstacker_b200 import api, capi

dev = torch.device("cuda", 0)
pool = bench.make_frames_gpu(65, 2, dev)
ro = api.registration_options(motion_type=capi.MOTION_AFFINE, interpolation=capi.INTER_CUBIC,
                              ecc=dict(ecc_method=capi.ECC_INVERSE_COMPOSITIONAL_LM, ecch_max_level=-1))
so = api.stack_options(registration=ro, accumulation_method=capi.STACK_WEIGHTED_AVERAGE, max_batch=64)
pipe = api.c_image_stacking_pipeline(so)
pipe.set_reference(capi.device_mat(pool[0].data_ptr(), bench.H, bench.W, np.float32))
res = pipe.add_frames([capi.device_mat(pool[j].data_ptr(), bench.H, bench.W, np.float32) for j in range(1, 65)])
its = np.array([r["iterations"] for r in res])
print("iterations per frame: mean %.1f min %d max %d" % (its.mean(), its.min(), its.max()), "ok", sum(r["ok"] for r in res))
print("keys", list(res[0].keys()))
print(res[0])

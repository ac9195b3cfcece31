"""Debug helper: solver trials per pyramid level of the bench workload (config #2), from the library's trace records."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench
from serstacker_b200 import api, capi

dev = torch.device("cuda", 0)
pool = bench.make_frames_gpu(9, 2, dev).cpu().numpy()
ro = api.registration_options(motion_type=capi.MOTION_AFFINE, interpolation=capi.INTER_CUBIC,
                              ecc=dict(ecc_method=capi.ECC_INVERSE_COMPOSITIONAL_LM, ecch_max_level=-1))
reg = api.c_frame_registration(ro)
reg.setup_reference_frame(pool[0])
capi.lib.ssk_reg_set_trace.argtypes = [C.c_void_p, C.c_int]
for i in range(1, 9):
    capi.check(capi.lib.ssk_reg_set_trace(reg._h, 4096))
    reg.register_frame(pool[i])
    rec = np.zeros((4096, 40), np.float32)
    n = C.c_int(0)
    capi.check(capi.lib.ssk_reg_get_trace(reg._h, rec.ctypes.data_as(C.POINTER(C.c_float)), 4096, C.byref(n)))
    r = rec[:n.value]
    trials = r[r[:, 1] != 9]
    lv = trials[:, 0].astype(int)
    print("frame", i, "records", n.value, "trials per level (0 = finest):", np.bincount(lv, minlength=6).tolist(), "iterations", reg.status.num_iterations)

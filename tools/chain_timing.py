"""Wall time of the stateless entry points the config #5 chain calls (device buffers): where the per-frame time goes."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from serstacker_b200 import api, capi
W, H = 2448, 2048
dev = torch.device("cuda", 0)
f = torch.rand((H, W, 3), device=dev)
wmap = torch.empty((H, W), device=dev); wblur = torch.empty((H, W), device=dev)
acc = api.c_weigthed_average()
m = capi.device_mat(f.data_ptr(), H, W, np.float32, cn=3)
mw, mb = capi.device_mat(wmap.data_ptr(), H, W, np.float32), capi.device_mat(wblur.data_ptr(), H, W, np.float32)
def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
print("ssk_lpg           %.1f us" % t(lambda: capi.check(capi.lib.ssk_lpg(C.byref(m), 6.0, 2.0, 0, 0, C.byref(mw)))))
print("ssk_gaussian_blur %.1f us" % t(lambda: capi.check(capi.lib.ssk_gaussian_blur(C.byref(mw), 1.0, 1.0, C.byref(mb)))))
print("ssk_acc_add       %.1f us" % t(lambda: capi.check(capi.lib.ssk_acc_add(acc._h, C.byref(m), C.byref(mb), 0))))
print("empty sync        %.1f us" % t(lambda: torch.cuda.synchronize()))
api.set_stream_ordered(True)
print("stream-ordered:")
print("ssk_lpg           %.1f us" % t(lambda: capi.check(capi.lib.ssk_lpg(C.byref(m), 6.0, 2.0, 0, 0, C.byref(mw)))))
print("ssk_gaussian_blur %.1f us" % t(lambda: capi.check(capi.lib.ssk_gaussian_blur(C.byref(mw), 1.0, 1.0, C.byref(mb)))))
print("ssk_acc_add       %.1f us" % t(lambda: capi.check(capi.lib.ssk_acc_add(acc._h, C.byref(m), C.byref(mb), 0))))
def chain():
    capi.check(capi.lib.ssk_lpg(C.byref(m), 6.0, 2.0, 0, 0, C.byref(mw)))
    capi.check(capi.lib.ssk_gaussian_blur(C.byref(mw), 1.0, 1.0, C.byref(mb)))
    capi.check(capi.lib.ssk_acc_add(acc._h, C.byref(m), C.byref(mb), 0))
print("chain             %.1f us" % t(chain, 50))
api.device_synchronize()
api.set_stream_ordered(False)

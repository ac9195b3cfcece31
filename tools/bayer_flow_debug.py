"""Debug helper: bayer_average + eccflow in ssk_stack against the oracle - separates the map difference (eccflow's rounding
envelope on this scene) from the gather (oracle accumulator fed with the DEVICE maps must equal the device stack)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, cv2
from serstacker_b200 import api, synth
from oracle import pipeline as opl, registration as oreg, accumulation as oacc
from oracle.debayer import debayer_nn2
import test_gpu_eccflow as T

W, H, N = [int(v) for v in (sys.argv[1:4] + [320, 224, 5][len(sys.argv) - 1:])]
frames, bpp = T._bayer_turbulent_sequence(W, H, N, seed=41)
oo = T._flow_registration_options(0, 3, cv2.INTER_LINEAR)
ro = oreg.FrameRegistration(oo)
bgr = [opl.to_float_frame(debayer_nn2(f, 8), bpp) for f in frames]
raw = [opl.to_float_frame(f, bpp) for f in frames]
ro.setup_reference_frame(bgr[0], None)
rg = api.c_frame_registration(api.registration_options(motion_type=0, interpolation=1, enable_eccflow_registration=1,
                                                      ecc=dict(ecc_method=3, ecch_max_level=-1)))
rg.setup_reference_frame(bgr[0])
acc_oo, acc_og = oacc.BayerAverage(), oacc.BayerAverage()
acc_oo.set_bayer_pattern(8); acc_og.set_bayer_pattern(8)
for f, r in zip(bgr, raw):
    assert ro.register_frame(f, None) and rg.register_frame(f)
    mo, mg = ro.current_remap, rg.current_remap()
    d = np.abs(mo - mg).max(axis=-1)
    print("map diff: max %.3g  99.9%% %.3g  mean %.3g" % (d.max(), np.quantile(d, .999), d.mean()))
    _, mask_o = ro.custom_remap(mo, f, None, oo.interpolation, oo.border_mode, oo.border_value)
    _, mask_g = ro.custom_remap(mg, f, None, oo.interpolation, oo.border_mode, oo.border_value)
    acc_oo.set_remap(mo); acc_oo.add(r, mask_o)
    acc_og.set_remap(mg); acc_og.add(r, mask_g)
avg_oo, m_oo = acc_oo.compute()
avg_og, m_og = acc_og.compute()
ropt = api.registration_options(motion_type=0, interpolation=1, enable_eccflow_registration=1, ecc=dict(ecc_method=3, ecch_max_level=-1))
p = api.c_image_stacking_pipeline(api.stack_options(registration=ropt, accumulation_method=2, bayer_colorid=8, max_batch=4))
p.set_reference(frames[0], bpp=bpp)
p.add_frames(frames)
avg_g, m_g = p.compute()
def rel(a, b, m):
    return float(np.sqrt(((a[m] - b[m]) ** 2).sum()) / np.sqrt((b[m] ** 2).sum()))
m = (m_oo > 0) & (m_g > 0) & (m_og > 0)
print("stack vs oracle(own maps):    rel-L2 %.3g   mask mismatch %.3g" % (rel(avg_g, avg_oo, m), (m_g != m_oo).mean()))
print("stack vs oracle(device maps): rel-L2 %.3g   max|d| %.3g  mask mismatch %.3g" % (rel(avg_g, avg_og, m), np.abs(avg_g - avg_og)[m].max(), (m_g != m_og).mean()))
print("oracle(own) vs oracle(device maps): rel-L2 %.3g" % rel(avg_og, avg_oo, m))

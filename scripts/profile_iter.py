"""Short driver for ncu: config #5 rig (64 cameras x 100k frames), one warm-up solve and one 3-iteration LM solve."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rig = synth.make_config(5, n_frame=frames)
s = m.Solver(device=0)
s.set_rig(rig)
s.set_parameters(rig["params_init"])
s.save_parameters()
for it in (5, 3):
    s.restore_parameters()
    rep = s.solve(mode=m.capi.MODE_LM, crit_type=1, max_count=it)
    print("iters", rep["iterations"], "device_ms", rep["device_ms"], "kernels", rep["kernel_launches"], "cost", rep["cost"])
    if os.environ.get("MCCBA_PROFILE") == "1":      # per-phase CUDA-event times of the plain-stream path (ms per iteration)
        print("   phases [schur, reduce, exchange, decide+solve+camera, frame_update, resid_jac] ms:", np.round(s.last_kernel_ms(), 4))
print("k1 ms", s.time_eval(5))
s.close()

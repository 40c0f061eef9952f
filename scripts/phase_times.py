"""Per-phase CUDA-event times of one iteration outside the graph (MCCBA_PROFILE=1): config #5 rig, reference schedule with a
COUNT criterion, so that every launch does work.  python scripts/phase_times.py [frames] [cams]"""
import os, sys
os.environ["MCCBA_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rig = synth.make_config(5, n_frame=frames)
s = m.Solver(device=0)
s.set_rig(rig)
for rep_i in range(2):
    s.set_parameters(rig["params_init"])
    rep = s.solve(mode=0, crit_type=1, max_count=30)
    t = s.last_kernel_ms() * 1e3 * 31 / 30
    print("us per iteration: schur %.1f  reduce %.1f  exchange %.1f  decide+solve+camera %.1f  frame_update %.1f  resid_jac %.1f  | sum %.1f" % (*t, t.sum()))
s.close()

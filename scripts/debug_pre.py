import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth
rig = synth.make_rig(n_cam=8, n_frame=400, seed=1002)
for pre in ("none", "eval", "reduced", "error"):
    s = m.Solver(device=0)
    s.set_rig(rig); s.set_parameters(rig["params_init"])
    if pre == "eval": s.eval()
    if pre == "reduced": s.reduced_system(1e-3)
    if pre == "error": s.reproj_error()
    rep = s.solve(mode=0, crit_type=1, max_count=3, check=False)
    print(pre, rep["rc"], rep["iterations"], rep["cost"])
    s.close()

"""Short driver for ncu / timing of the omnidir::calibrate path: config #3 (single Mei camera, 5000 frames).
python scripts/omni_iter.py [frames] [iterations]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
r3 = synth.make_config(3, n_frame=frames)
n3 = r3["n_frame"]
pt = r3["params_true"].reshape(-1, 6)
K5, D, xi = r3["cam_K5"][0], r3["cam_dist8"][0][:4], r3["cam_xi"][0]
poses = np.array([pt[r3["edge_pv"][e] - 1] for e in range(n3)])
p3 = np.concatenate([poses.ravel(), [K5[0] * 1.03, K5[1] * 1.03, K5[4], K5[2], K5[3], xi + 0.1], np.zeros(4)])
s = m.Solver(device=0)
s.omni_set_observations(r3["edge_off"], r3["obj"], r3["img"])
for rep_i in range(3):
    s.omni_set_parameters(p3)
    rep = s.omni_solve(0, 1, iters, 0.0)
    print("iterations", rep["iterations"], "device_ms", rep["device_ms"], "us/iter", rep["device_ms"] * 1e3 / max(rep["iterations"], 1), "rms", rep["rms"])
s.close()

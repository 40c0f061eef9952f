import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_camera_calibration_b200 as m
s = m.Solver(device=0)
rng = np.random.default_rng(0)
n = 378
M = rng.standard_normal((n, n)); S = M @ M.T + n * np.eye(n); g = rng.standard_normal(n)
x, ms = s.debug_solve_dense(S, g, 2)
x, ms = s.debug_solve_dense(S, g, 2)
print("ms", ms, "rel err", np.abs(x - np.linalg.solve(S, g)).max() / np.abs(x).max())

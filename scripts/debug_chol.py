import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_camera_calibration_b200 as m
s = m.Solver(device=0)
rng = np.random.default_rng(0)
for n in [6, 18, 31, 32, 33, 42, 90, 127, 160, 378, 420]:
    M = rng.standard_normal((n, n)); S = M @ M.T + n * np.eye(n); g = rng.standard_normal(n)
    ref = np.linalg.solve(S, g)
    for blocked in (2, 1, 0):
        x, ms = s.debug_solve_dense(S, g, blocked)
        x, ms = s.debug_solve_dense(S, g, blocked)
        print("n", n, "blocked", blocked, "rel err %.2e" % (np.abs(x - ref).max() / np.abs(ref).max()), "ms %.4f" % ms)
try:
    s.debug_solve_dense(-np.eye(5), np.ones(5))
except m.MccbaError as e:
    print("expected failure:", e)

"""Times the residual / Jacobian pass alone on the config #5 rig (development aid): python scripts/k1_time.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rig = synth.make_config(5, n_frame=frames)
for prec, name in ((1, "mixed"), (2, "fast32"), (0, "fp64")):
    s = m.Solver(device=0, precision=prec)
    s.set_rig(rig)
    s.set_parameters(rig["params_init"])
    t = min(s.time_eval(20) for _ in range(3))
    rep = s.solve(mode=1, crit_type=1, max_count=10)
    print("%s %-7s k1 %.1f us  (%.0f GB/s algorithmic)  iteration %.1f us  cost %.9e" % (os.environ.get("MCCBA_LIB", "default").split("/")[-1], name, t * 1e3, 20.0 * rig["n_points"] / t / 1e6, rep["device_ms"] * 1e3 / 10, rep["cost"]))
    s.close()

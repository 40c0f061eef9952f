"""Deviation of the first Gauss-Newton iterates of the double-sided problem (tests/test_double_side.py) from the oracle,
per precision policy (development aid; MCCBA_LIB selects a kernel variant):  python scripts/ds_dev.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multi_camera_calibration_b200 as m
from tests import test_double_side as t
from oracle import dense_reenact as dr
r, prob = t._problem(40)
rec = []
dr.ds_optimize(prob, r["edge_back"], r["cam_pose"], r["ds_params_init"], 1, 4, 0.0, record=rec)
print("obj z range", float(np.abs(r["obj"][:,2]).max()))
for name, prec in (("fp64", m.capi.PRECISION_FP64), ("mixed", m.capi.PRECISION_MIXED), ("fast32", m.capi.PRECISION_FAST32)):
    s = m.Solver(device=0, precision=prec)
    s.set_rig(r); s.ds_set_problem(r["edge_back"], r["cam_pose"])
    out=[]
    for k in (1,2,4):
        s.ds_set_parameters(r["ds_params_init"]); rep = s.ds_solve(1, k, 0.0); p = s.ds_get_parameters(); ref = rec[k-1]["params"]
        d = np.abs(p - ref) / np.maximum(np.abs(ref), 1.0)
        out.append((float(d.max()), int(d.argmax())))
    print(os.environ.get("MCCBA_LIB","default").split("/")[-1], name, out)
    s.close()

import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import rigs
from tests.test_parity_gpu import RIGS, _param_rel
import multi_camera_calibration_b200 as m

for name in ["omni3", "pinhole8"]:
    rig = rigs.make_rig(**RIGS[name])
    O = rigs.to_oracle_rig(rig)
    s = m.Solver(device=0)
    s.set_rig(rig)
    p0 = rig["params_init"]
    ref = O.solve(p0, mode=1, crit_type=3, max_count=60, eps=1e-7, trace_cap=60)
    print(name, "oracle iters", ref["iters"], "cost", ref["cost"], "change", ref["change"])
    print(ref["trace"])
    for k in range(1, ref["iters"] + 3):
        s.set_parameters(p0)
        rep = s.solve(mode=1, crit_type=1, max_count=k)
        r2 = O.solve(p0, mode=1, crit_type=1, max_count=k)
        print(k, "gpu acc/rej", rep["accepted"], rep["rejected"], "cost %.12e vs %.12e" % (rep["cost"], r2["cost"]),
              "change %.3e vs %.3e" % (rep["change"], r2["change"]), "lam %.3e %.3e" % (rep["lam"], r2["lam"]),
              "prel %.2e" % _param_rel(s.get_parameters(), r2["params"]))
    s.set_parameters(p0)
    rep = s.solve(mode=1, crit_type=3, max_count=60, eps=1e-7)
    print("converged gpu", rep, "prel", _param_rel(s.get_parameters(), ref["params"]))
    s.close()

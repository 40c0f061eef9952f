import sys, os, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth
from tests import rigs
from oracle import oracle as orc
rig = synth.make_config(5)
O = rigs.to_oracle_rig(rig)
s = m.Solver(device=0); s.set_rig(rig)
nC = rig["n_cam"]
def rel(a, b): return np.abs(a - b) / np.maximum(np.abs(b), 1.0)
for mode in (1, 0):
    for k in (1, 2, 4, 8):
        s.set_parameters(rig["params_init"]); rep = s.solve(mode=mode, crit_type=1, max_count=k)
        p = s.get_parameters()
        orc.set_num_threads(16); ref = O.solve(rig["params_init"], mode=mode, crit_type=1, max_count=k)
        orc.set_num_threads(1); ref1 = O.solve(rig["params_init"], mode=mode, crit_type=1, max_count=k) if k <= 2 else ref
        r = rel(p, ref["params"]); r1 = rel(ref1["params"], ref["params"])
        ncp = 6 * (nC - 1)
        w = int(np.argmax(r))
        print("mode %d k %d cam_rel %.2e frame_rel %.2e (n>1e-7: %d, >1e-9: %d) worst idx %d (vertex %d comp %d) gpu %.12g ref %.12g | oracle 16t vs 1t: cam %.2e frame %.2e | cost rel %.2e"
              % (mode, k, r[:ncp].max(), r[ncp:].max(), int((r > 1e-7).sum()), int((r > 1e-9).sum()), w, w // 6 + 1, w % 6, p[w], ref["params"][w],
                 r1[:ncp].max(), r1[ncp:].max(), abs(rep["cost"] - ref["cost"]) / ref["cost"]), flush=True)

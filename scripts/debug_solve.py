import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import rigs
from tests.test_parity_gpu import RIGS
import multi_camera_calibration_b200 as m
from oracle import oracle as orc

name = sys.argv[1] if len(sys.argv) > 1 else "mixed4_v3_ragged"
rig = rigs.make_rig(**RIGS[name])
O = rigs.to_oracle_rig(rig)
for graph in (True, False):
    s = m.Solver(device=0, use_graph=graph)
    s.set_rig(rig)
    s.set_parameters(rig["params_init"])
    S, gs = s.reduced_system(0.0)
    w = np.linalg.eigvalsh(S)
    print("graph", graph, "S eig min/max", w.min(), w.max(), "sym", np.abs(S - S.T).max())
    dc = np.linalg.solve(S, gs)
    print(" |dc|", np.linalg.norm(dc))
    rep = s.solve(mode=0, crit_type=1, max_count=2, check=False)
    print(" report", rep)
    ref = O.solve(rig["params_init"], mode=0, crit_type=1, max_count=2)
    print(" param diff", np.abs(s.get_parameters() - ref["params"]).max())
    s.close()

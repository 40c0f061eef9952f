import sys, os, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_camera_calibration_b200 import synth
from tests import rigs, harness
from oracle import oracle as orc
rig = synth.make_config(5)
O = rigs.to_oracle_rig(rig)
p = rig["params_init"]
POL = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for lam in (1e-3,):
    t = time.time(); out = harness.rig_step(rig, p, lam); print("harness", time.time() - t)
    O.eval(p, policy=POL); rc, step, S, gs = O.solve_normal(p, lam)
    r = np.abs(out["step"] - step) / np.maximum(np.abs(p), 1.0)
    ncp = 6 * 63
    print("lam", lam, "cam rel", r[:ncp].max(), "frame rel", r[ncp:].max(), "n>1e-9", (r > 1e-9).sum(), "n>1e-7", (r > 1e-7).sum())
    print("  S rel", np.abs(out["S"] - S).max() / np.abs(S).max() if lam > 0 else None, "gs rel", np.abs(out["gs"] - gs).max() / np.abs(gs).max() if lam > 0 else None)
    o = np.argsort(-r)
    for i in o[:5]:
        v = i // 6
        print("   idx %d vertex %d comp %d step_h %.10g step_o %.10g |om| %.6f" % (i, v + 1, i % 6, out["step"][i], step[i], np.linalg.norm(p[6 * v:6 * v + 3])))
    np.save("/tmp/step_h_%g.npy" % lam, out["step"]); np.save("/tmp/step_o_%g.npy" % lam, step)

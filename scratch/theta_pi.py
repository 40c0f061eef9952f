import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_camera_calibration_b200 import synth
from tests import rigs, harness
from oracle import oracle as orc
rig = synth.make_rig(n_cam=64, n_frame=4000, seed=1005)
O = rigs.to_oracle_rig(rig)
p = rig["params_init"]
out = harness.rig_step(rig, p, 1e-3)
cost = O.eval(p)
g6o = O.blocks(1); H6o = O.blocks(0).reshape(-1, 6, 6)
def skew(v): return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
def Jl(om):
    th = np.linalg.norm(om); K = skew(om)
    return np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * (K @ K)
E = rig["edge_cam"].size
res = []
for e in range(E):
    c, pv = int(rig["edge_cam"][e]), int(rig["edge_pv"][e])
    omP, tP = p[6 * (pv - 1):6 * (pv - 1) + 3], p[6 * (pv - 1) + 3:6 * pv]
    omC, tC = (p[6 * (c - 1):6 * (c - 1) + 3], p[6 * (c - 1) + 3:6 * c]) if c > 0 else (np.zeros(3), np.zeros(3))
    om3, _, _ = orc.compose_motion(omP, tP, omC, tC)
    D = np.eye(6); D[:3, :3] = Jl(om3)
    g = D.T @ out["blocks"][e, 21:27]
    rel = np.abs(g - g6o[e]).max() / np.abs(g6o[e]).max()
    res.append((np.pi - np.linalg.norm(om3), rel, np.linalg.norm(omP)))
res = np.array(res)
o = np.argsort(res[:, 0])
print("closest to pi (pi - theta3, g6 rel diff, |omP|):")
for i in o[:8]: print("  %.3e  %.3e  %.4f" % tuple(res[i]))
print("median rel diff", np.median(res[:, 1]), "max", res[:, 1].max(), "at pi-theta3 =", res[np.argmax(res[:, 1]), 0])
# frames near pi in |omP|
o2 = np.argsort(np.abs(np.pi - res[:, 2]))
print("closest |omP| to pi:")
for i in o2[:5]: print("  pi-|omP| %.3e  pi-th3 %.3e rel %.3e" % (np.pi - res[i, 2], res[i, 0], res[i, 1]))
print("---- step comparison")
for lam in (0.0, 1e-3):
    rc, step, S, gs = O.solve_normal(p, lam)
    out = harness.rig_step(rig, p, lam)
    d = np.abs(out["step"] - step)
    sc = np.maximum(np.abs(p), 1.0)
    r = d / sc
    o = np.argsort(-r)
    print("lam", lam, "max rel step diff (vs max(|p|,1))", r.max(), "n>1e-9:", (r > 1e-9).sum())
    for i in o[:6]:
        v = i // 6
        print("   idx %d vertex %d comp %d step_h %.10g step_o %.10g |om| %.6f p %.6f" % (i, v + 1, i % 6, out["step"][i], step[i], np.linalg.norm(p[6 * v:6 * v + 3]), p[i]))

"""How the error of the MIXED and FAST32 policies against the FP64 policy depends on the board: corner count vs angular
extent (4-camera rig, 400 frames, boards 1.2-2 m away, 15 LM iterations).  Output: profiles/r2_precision_vs_board.txt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multi_camera_calibration_b200 as m
from tests import rigs

def study(name, rig, skw):
    res = {}
    for prec in (m.capi.PRECISION_FP64, m.capi.PRECISION_MIXED, m.capi.PRECISION_FAST32):
        s = m.Solver(device=0, precision=prec)
        s.set_rig(rig); s.set_parameters(rig["params_init"])
        rep = s.solve(**skw)
        res[prec] = (s.get_parameters(), rep["cost"])
        s.close()
    nC = rig["n_cam"]
    p64 = res[0][0].reshape(-1, 6)
    out = []
    for prec in (1, 2):
        p = res[prec][0].reshape(-1, 6)
        rel = np.abs(p - p64) / np.maximum(np.abs(p64), 1.0)
        out.append("%s: cams %.1e frames rot %.1e trans %.1e" % ("mixed" if prec == 1 else "fast32", rel[:nC - 1].max(), rel[nC - 1:, :3].max(), rel[nC - 1:, 3:].max()))
    print("%-34s %s" % (name, " | ".join(out)), flush=True)

lm = dict(mode=1, crit_type=1, max_count=15, lambda0=1e-3)
orig_board = rigs.board
for nx, ny, pitch in ((9, 6, 40.0), (4, 3, 40.0), (4, 3, 107.0), (9, 6, 13.0), (9, 6, 80.0), (7, 5, 40.0), (13, 10, 40.0)):
    rigs.board = lambda a=9, b=6, pitch_=pitch, p=40.0: orig_board(a, b, pitch_)
    try:
        rig = rigs.make_rig(n_cam=4, n_frame=400, cam_models=[0, 0, 1, 0], seed=5, nx=nx, ny=ny)
    except Exception as e:
        print("board %dx%d pitch %.0f: generator refused (%s)" % (nx, ny, pitch, str(e)[:50])); continue
    study("board %dx%d pitch %.0f mm (%3d corners)" % (nx, ny, pitch, nx * ny), rig, lm)
rigs.board = orig_board

"""Host-side breakdown of one end-to-end step on config #5 (MCCBA_TIMING=1 prints the phases of mccba_set_observations)."""
import os, sys, time
os.environ.setdefault("MCCBA_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth
rig = synth.make_config(5)
pin = {k: torch.from_numpy(np.ascontiguousarray(rig[k])).pin_memory().numpy() for k in ("obj", "img", "params_init")}
s = m.Solver(device=0)
s.set_cameras(rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])
kw = dict(mode=m.capi.MODE_LM, crit_type=m.capi.CRIT_COUNT, max_count=20, lambda0=1e-3, lambda_up=10.0, lambda_down=1.0 / 3.0)
for i in range(4):
    t0 = time.perf_counter()
    s.set_observations(rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], pin["obj"], pin["img"])
    t1 = time.perf_counter()
    s.set_parameters(pin["params_init"])
    t2 = time.perf_counter()
    r = s.solve(**kw)
    t3 = time.perf_counter()
    p = s.get_parameters()
    t4 = time.perf_counter()
    print("step %d: set_observations %.2f ms, set_parameters %.2f, solve %.2f (device %.2f), get_parameters %.2f, total %.2f" %
          (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, r["device_ms"], (t4 - t3) * 1e3, (t4 - t0) * 1e3), flush=True)
s.close()

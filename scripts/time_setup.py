import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth
rig = synth.make_config(5)
pin = {k: torch.from_numpy(np.ascontiguousarray(rig[k])).pin_memory().numpy() for k in ("obj", "img", "params_init")}
s = m.Solver(device=0)
s.set_cameras(rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])
for rep in range(3):
    t0 = time.perf_counter()
    s.set_observations(rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], pin["obj"], pin["img"])
    t1 = time.perf_counter()
    s.set_parameters(pin["params_init"])
    t2 = time.perf_counter()
    r = s.solve(mode=1, crit_type=1, max_count=20)
    t3 = time.perf_counter()
    p = s.get_parameters()
    t4 = time.perf_counter()
    print("set_obs %.1f ms set_par %.1f solve %.1f (device %.1f) get %.1f" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, r["device_ms"], (t4-t3)*1e3), flush=True)

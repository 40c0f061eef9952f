import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def fresh_id():
    """every communicator needs its own ncclUniqueId"""
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(m.capi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().tolist())
rig = synth.make_rig(n_cam=8, n_frame=400, seed=1002)
sh = synth.shard_rig(rig, rank, world)
for pre in ("none", "eval", "allreduce", "reduced", "reduced2", "error"):
    s = m.Solver(device=local, rank=rank, nranks=world, nccl_id=fresh_id(), use_graph=True)
    s.set_rig(sh)
    s.set_parameters(sh["params_init"])
    extra = ""
    if pre == "eval":
        s.eval()
    if pre == "allreduce":
        extra = str(s.allreduce_sum(np.ones(4) * (rank + 1)))
    if pre in ("reduced", "reduced2"):
        S1, g1 = s.reduced_system(1e-3)
        if pre == "reduced2":
            S2, g2 = s.reduced_system(1e-3)
            extra = "S self-consistency %.2e" % np.abs(S1 - S2).max()
    if pre == "error":
        extra = str(s.reproj_error()["rms"])
    rep = s.solve(mode=0, crit_type=1, max_count=3, check=False)
    if rank == 0:
        print("pre", pre, "->", rep["rc"], rep["iterations"], rep["cost"], extra, flush=True)
    s.close()
dist.barrier()
dist.destroy_process_group()

"""Per-phase timing of the LM iteration under torchrun (MCCBA_PROFILE=1: events around K2 | K3a | exchange | solve | K4 | K1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.tensor(list(m.capi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(idt, 0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rig = synth.make_config(5, n_frame=frames, frame_stream=rank)
s = m.Solver(device=local, rank=rank, nranks=world, nccl_id=bytes(idt.cpu().tolist()))
s.set_rig(rig)
s.set_parameters(rig["params_init"])
s.save_parameters()
for it in (5, 10):
    s.restore_parameters()
    dist.barrier()
    rep = s.solve(mode=m.capi.MODE_LM, crit_type=1, max_count=it)
ms = s.last_kernel_ms()
if rank == 0:
    print("exchange_mode", s.exchange_mode(), "iters", rep["iterations"], "device_ms", rep["device_ms"])
    print("per launch us: K2 %.1f | K3a %.1f | exchange %.1f | decide+solve+camera %.1f | K4 %.1f | K1 %.1f" % tuple(1e3 * v for v in ms[:6]))
s.close()
dist.barrier()
dist.destroy_process_group()

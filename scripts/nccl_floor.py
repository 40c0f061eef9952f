"""How long does the per-iteration collective take on its own?  ncclAllReduce of the packed reduced system (143 266
doubles at 64 cameras), back to back, CUDA-event timed (development aid)."""
import os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 378 * 378 + 378 + 4
x = torch.ones(n, dtype=torch.float64, device="cuda")
for _ in range(20): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
if rank == 0: print("world %d: all_reduce of %d doubles: %.1f us per call" % (world, n, e0.elapsed_time(e1) * 1000 / 200))
dist.destroy_process_group()

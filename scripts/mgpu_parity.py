"""Multi-GPU parity: frames shard across ranks (one process per GPU, torchrun), the library sums the reduced camera
system with one ncclAllReduce per iteration, every rank solves it redundantly.  Rank 0 gathers the parameters and
checks them against the CPU oracle run on the WHOLE rig.

Tolerance: 1e-6 relative (the north star gate).  The all-reduce changes the summation order of S, and the undamped
Gauss-Newton system has condition ~1e9, so parameters agree to ~1e-8..1e-7 rather than the 1e-12 of one GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import multi_camera_calibration_b200 as m
from multi_camera_calibration_b200 import synth


def main():
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
    ok = True
    for name, kw, mode in (("pinhole8", dict(n_cam=8, n_frame=400, seed=1002), 0),
                           ("mixed6", dict(n_cam=6, n_frame=301, seed=77, models=[0, 1, 0, 1, 0, 1], views_per_frame=3), 1)):
        rig = synth.make_rig(**kw)
        sh = synth.shard_rig(rig, rank, world)
        s = m.Solver(device=local, rank=rank, nranks=world, nccl_id=fresh_id())
        s.set_rig(sh)
        s.set_parameters(sh["params_init"])
        S, gs = s.reduced_system(1e-3)
        rep = s.solve(mode=mode, crit_type=1, max_count=6)
        xmode = s.exchange_mode()
        p_local = s.get_parameters()
        err = s.reproj_error()
        s.close()
        # gather frame parameters (cameras are replicated)
        nC = rig["n_cam"]
        fr = torch.from_numpy(p_local[6 * (nC - 1):]).cuda()
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([fr.numel()], dtype=torch.int64, device="cuda"))
        mx = int(max(int(x) for x in sizes))
        pad = torch.zeros(mx, dtype=torch.float64, device="cuda"); pad[:fr.numel()] = fr
        parts = [torch.zeros(mx, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(parts, pad)
        cams = torch.from_numpy(p_local[:6 * (nC - 1)]).cuda()
        cam_all = [torch.zeros_like(cams) for _ in range(world)]
        dist.all_gather(cam_all, cams)
        if rank == 0:
            from oracle import oracle as orc    # checker only
            from tests import rigs
            O = rigs.to_oracle_rig(rig)
            ref = O.solve(rig["params_init"], mode=mode, crit_type=1, max_count=6)
            full = np.concatenate([p_local[:6 * (nC - 1)]] + [parts[r][:int(sizes[r])].cpu().numpy() for r in range(world)])
            scale = np.maximum(np.abs(ref["params"]), 1.0)
            rel = float(np.max(np.abs(full - ref["params"]) / scale))
            O.eval(rig["params_init"])
            _, _, So, gso = O.solve_normal(rig["params_init"], 1e-3)
            rs = float(np.abs(S - So).max() / np.abs(So).max())
            same_cams = all(torch.equal(cam_all[0], c) for c in cam_all)
            eo = O.error(ref["params"])
            line = "%s world=%d exchange=%s iters=%d param_rel=%.2e S_rel=%.2e cams_bit_identical=%s rms=%.9f/%.9f cost=%.9e/%.9e" % (
                name, world, {0: "none", 1: "nccl", 2: "peer"}[xmode], rep["iterations"], rel, rs, same_cams, err["rms"], eo["rms"], rep["cost"], ref["cost"])
            print(line, flush=True)
            ok = ok and rel < 1e-6 and rs < 1e-9 and same_cams and abs(err["rms"] - eo["rms"]) < 1e-9 * eo["rms"] \
                and abs(rep["cost"] - ref["cost"]) < 1e-9 * ref["cost"] and rep["iterations"] == 6
    if rank == 0:
        print("MGPU_PARITY_OK" if ok else "MGPU_PARITY_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Failure path of the peer-memory exchange (2 ranks, torchrun): the ranks launch DIFFERENT numbers of iterations, so the
rank that goes on finds no partner in the exchange.  Its bounded spin (MCCBA_P2P_TIMEOUT_MS) must end the solve with
MCCBA_ERR_NCCL instead of hanging the GPU, and after the epoch is re-agreed at the start of the next solve both ranks must
solve normally again (camera parameters bit-identical, equal to the first clean run)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ["MCCBA_P2P_TIMEOUT_MS"] = "300"
os.environ.pop("MCCBA_P2P", None)


def main():
    import torch
    import torch.distributed as dist
    import multi_camera_calibration_b200 as m
    from multi_camera_calibration_b200 import synth
    from scripts import mgpu_parity
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rig = synth.make_rig(n_cam=8, n_frame=400, seed=1002)
    sh = synth.shard_rig(rig, rank, world)
    s = m.Solver(device=local, rank=rank, nranks=world, nccl_id=mgpu_parity._fresh_id(m, dist, rank))
    s.set_rig(sh)
    res = dict(rank=rank, exchange=s.exchange_mode())
    # clean run
    s.set_parameters(sh["params_init"])
    s.solve(mode=0, crit_type=1, max_count=6)
    p_clean = s.get_parameters()
    # unequal launch counts: rank 0 goes on after rank 1 has stopped
    s.set_parameters(sh["params_init"])
    t0 = time.perf_counter()
    err = None
    try:
        s.solve(mode=0, crit_type=1, max_count=6 if rank == 0 else 3)
    except m.MccbaError as e:
        err = str(e)
    res["seconds"] = time.perf_counter() - t0
    res["error"] = err
    dist.barrier()
    # both ranks again, in step: the epoch is re-agreed at the start of the solve
    s.set_parameters(sh["params_init"])
    rep = s.solve(mode=0, crit_type=1, max_count=6)
    p_again = s.get_parameters()
    res["recovered_bit_identical"] = bool(np.array_equal(p_clean, p_again)) and rep["iterations"] == 6
    s.close()
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        ok = (out[0]["exchange"] == 2 and out[0]["error"] is not None and "exchange timed out" in out[0]["error"] and out[0]["seconds"] < 20.0
              and out[1]["error"] is None and all(o["recovered_bit_identical"] for o in out))
        print("mgpu_timeout " + json.dumps(out), flush=True)
        print("MGPU_TIMEOUT_OK" if ok else "MGPU_TIMEOUT_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

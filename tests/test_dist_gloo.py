"""World-size-2 CPU test (gloo) of the multi-GPU data path: frames shard across ranks, every rank builds the partial
reduced camera system of ITS frames, ONE all-reduce of the packed buffer [S | g | cost] sums it, and every rank
then solves the identical system redundantly.  The per-shard arithmetic here is the product's own device math
compiled for the host (tests/harness); the collective is torch.distributed/gloo instead of NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, seed, lam, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from multi_camera_calibration_b200 import synth
    from tests import harness
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rig = synth.make_rig(n_cam=4, n_frame=37, seed=seed, models=[0, 1, 0, 0], views_per_frame=2)
    sh = synth.shard_rig(rig, rank, world)
    out = harness.rig_step(sh, sh["params_init"], lam)
    ns = 6 * (rig["n_cam"] - 1)
    buf = torch.from_numpy(np.concatenate([out["S"].ravel(), out["gs"], [out["blocks"][:, 27].sum()]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)          # the single collective of one LM iteration
    b = buf.numpy()
    S, gs, cost = b[:ns * ns].reshape(ns, ns), b[ns * ns:ns * ns + ns], b[-1]
    dc = np.linalg.solve(S, gs)                         # redundant solve on every rank
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), S=S, gs=gs, cost=cost, dc=dc, f0=sh["global_frame_range"][0],
             f1=sh["global_frame_range"][1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lam", [1e-3])
def test_frame_sharded_reduced_system(tmp_path, lam):
    import torch.multiprocessing as mp
    from multi_camera_calibration_b200 import synth
    from tests import harness
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world, seed = 2, 31
    mp.spawn(_worker, args=(world, port, seed, lam, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz"); r1 = np.load(tmp_path / "rank1.npz")
    # identical inputs after the all-reduce => bit-identical redundant solves, no broadcast needed
    assert np.array_equal(r0["S"], r1["S"]) and np.array_equal(r0["dc"], r1["dc"])
    assert int(r0["f0"]) == 0 and int(r0["f1"]) == int(r1["f0"]) and int(r1["f1"]) == 37
    rig = synth.make_rig(n_cam=4, n_frame=37, seed=seed, models=[0, 1, 0, 0], views_per_frame=2)
    whole = harness.rig_step(rig, rig["params_init"], lam)
    assert np.abs(r0["S"] - whole["S"]).max() <= 1e-12 * np.abs(whole["S"]).max()
    assert np.abs(r0["gs"] - whole["gs"]).max() <= 1e-12 * np.abs(whole["gs"]).max()
    assert abs(float(r0["cost"]) - whole["blocks"][:, 27].sum()) <= 1e-12 * float(r0["cost"])
    dc = np.linalg.solve(whole["S"], whole["gs"])
    assert np.abs(r0["dc"] - dc).max() <= 1e-9 * np.abs(dc).max()

#!/usr/bin/env python
"""Benchmark of the calibration bundle-adjustment hot path (BASELINE.json metric: LM iterations/sec and
residual+Jacobian evals/sec at 1/2/4/8 B200 vs host CPU).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json configs[4] -- synthetic 64-camera pinhole rig, 100k frames, 9x6 board,
2 views per frame = 10.8 M corner observations PER GPU (frames shard across ranks; every rank holds the 64 cameras and
its own 100k frames: weak scaling, one exchange of the packed reduced camera system per iteration -- the library's
NVLink peer-memory kernel, or ncclAllReduce as its fallback; `exchange` in the JSON line says which).
A "step" = one optimizeExtrinsics call = `iters` (default 20, the reference's TermCriteria(COUNT, 20)) full LM
iterations on that rig.  value = corner observations processed per second through full LM iterations, whole job.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lm_corner_observations_per_sec"
UNIT = "corner observations/s through full LM iterations (whole job)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _pipe_block(M, k1_ms, clocks):
    """Arithmetic-pipe view of the dominant kernel: instruction counts per corner from the ncu source page of the committed
    capture (profiles/r2_summary.md), pipe rates measured on the box with scripts/ubench (profiles/r2_pipe_peaks.txt)."""
    p = os.path.join(ROOT, "profiles", "k1_pipes.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None
    mhz = (clocks or {}).get("sm_mhz") or d.get("sm_mhz", 1965.0)
    sm = d.get("sms", 148)
    out = dict(d)
    # lane-instructions per clock per SM measured for realistic operand patterns (3 distinct registers)
    fp64_us = d["fp64_instr_per_corner"] * M / (d["dfma_lane_instr_per_clk_sm"] * sm * mhz * 1e6) * 1e6
    fma_us = d["f32x2_instr_per_corner"] * M / (d["ffma2_lane_instr_per_clk_sm"] * sm * mhz * 1e6) * 1e6
    out.update(fp64_pipe_floor_us=fp64_us, fma_pipe_floor_us=fma_us, kernel_us=k1_ms * 1e3,
               frac_of_binding_pipe=max(fp64_us, fma_us) / (k1_ms * 1e3))
    return out


def _k1_traffic():
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "25"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = []
        for line in open(self.f.name):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 9:
                rows.append(parts)
        os.unlink(self.f.name)
        if not rows:
            return out
        try:
            sm = sorted(float(r[1]) for r in rows)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][2])
            out["power_w_max"] = max(float(r[3]) for r in rows)
        except Exception:
            pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in rows:
            for i, n in enumerate(names):
                if r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(rows)
        return out


def _rig(args, rank):
    from multi_camera_calibration_b200 import synth
    cams = synth.make_cameras(args.cams, args.seed)
    return synth.make_rig(n_cam=args.cams, n_frame=args.frames, seed=args.seed, cameras=cams, frame_stream=rank)


def _oracle_rig(rig):
    from oracle import oracle as orc                      # cpu_baseline / reference arm only
    return orc, orc.Rig(rig["n_cam"], rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], rig["obj"],
                        rig["img"], rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])


def config_dict(args, world):
    return {"workload": "BASELINE configs[4]: synthetic %d-camera pinhole rig, %d frames per GPU, 9x6 board, 2 views "
                        "per frame; one step = %d LM iterations (TermCriteria COUNT)" % (args.cams, args.frames, args.iters),
            "n_cameras": args.cams, "frames_per_gpu": args.frames, "corners_per_gpu": args.frames * 2 * 54,
            "reduced_system_n": 6 * (args.cams - 1), "lm_iterations_per_step": args.iters, "sharding": "frames x%d" % world,
            "l2": "inputs larger than L2 (216 MB of observations + 90 MB of per-edge blocks per pass vs 126 MB L2)",
            "seed": args.seed}


def _cpu_threads(orc):
    """torchrun exports OMP_NUM_THREADS=1: the CPU arm asks for every host core explicitly and reports what it got."""
    want = os.cpu_count() or 1
    try:
        want = len(os.sched_getaffinity(0))
    except Exception:
        pass
    orc.set_num_threads(want)
    return orc.num_threads()


def _bind_to_gpu_numa(local):
    """Multi-GPU runs: pin this rank to the CPUs of its GPU's NUMA node before the pinned host buffers are allocated and
    first touched, so that the per-step uploads of the N ranks do not cross the socket interconnect (what `numactl
    --cpunodebind --membind` does for a launcher that knows the topology).  Returns what was done, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else local
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(dev + "/numa_node").read())
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None


def run_reference(args):
    """CPU arm: the restated reference algorithm (oracle port -- the reference itself needs OpenCV C++ and Eigen and
    cannot be built in this image) with all host threads, on a bounded sample of the same workload: rank 0's shard of
    the rig (the per-GPU unit of the weak-scaling job; corner observations/s does not depend on how many such shards
    the job holds), `--ref-iters` LM iterations per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    rig = _rig(args, 0)
    orc, O = _oracle_rig(rig)
    cores = _cpu_threads(orc)
    iters = args.ref_iters
    M = rig["n_points"]
    kw = dict(mode=1, crit_type=1, max_count=iters, lambda0=1e-3, lambda_up=10.0, lambda_down=1.0 / 3.0)
    for _ in range(args.warmup):
        O.solve(rig["params_init"], **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.solve(rig["params_init"], **kw)
    dt = (time.perf_counter() - t0) / args.steps
    value = M * iters / dt
    sample = "one %d-frame shard (%d corners), %d LM iterations per step instead of %d" % (args.frames, M, iters, args.iters)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world), "lm_iters_per_sec": iters / dt,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def _time_solves(s, kw, steps, warmup, world, barrier):
    """(ms per solve, max over ranks, timed by the library with CUDA events on its own stream; last report)"""
    import torch
    import torch.distributed as dist
    rep = None
    for _ in range(warmup):
        s.restore_parameters()
        rep = s.solve(**kw)
    barrier()
    ms = 0.0
    for _ in range(steps):
        s.restore_parameters()
        rep = s.solve(**kw)
        ms += rep["device_ms"]
    barrier()
    t = torch.tensor([ms / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), rep


def _sharded_leg(m, dist, rig_local, kw, world, rank, local, barrier, steps=10, warmup=3):
    from scripts import mgpu_parity
    nccl_id = mgpu_parity._fresh_id(m, dist, rank) if world > 1 else None
    s2 = m.Solver(device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    s2.set_rig(rig_local)
    s2.set_parameters(rig_local["params_init"])
    s2.save_parameters()
    ms, rep = _time_solves(s2, kw, steps, warmup, world, barrier)
    err = s2.reproj_error()
    s2.close()
    return ms, rep, err


def _extra_legs(args, m, s, rig, pin, kw, world, rank, local, barrier):
    """Secondary measurements, each one a few solves: (a) the dense-graph reduced solver (tile DAG, MCCBA_CHOL=2) on the
    headline rig, (b) strong scaling of config #5 as written (args.frames frames in TOTAL, split over the ranks),
    (c) BASELINE configs[3] (16-camera mixed rig, 10k frames in total, split over the ranks), (d) at N > 1 the
    multi-GPU parity check against the whole-rig oracle (checker only, outside every timed region)."""
    import torch.distributed as dist
    from multi_camera_calibration_b200 import synth
    out = {}
    M = rig["n_points"]
    # (a) dense-graph solver on the same rig: what a rig without the ring topology would pay per iteration
    old = os.environ.get("MCCBA_CHOL")
    os.environ["MCCBA_CHOL"] = "2"
    try:
        s.set_observations(rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], pin["obj"], pin["img"])
        s.set_parameters(pin["params_init"])
        s.save_parameters()
        ms, rep = _time_solves(s, kw, 5, 3, world, barrier)
        out["dense_graph_solver"] = {"us_per_lm_iteration": ms * 1e3 / max(rep["iterations"], 1),
                                     "note": "same rig, reduced system solved by the dense tile-DAG Cholesky (MCCBA_CHOL=2)"}
    finally:
        if old is None:
            os.environ.pop("MCCBA_CHOL", None)
        else:
            os.environ["MCCBA_CHOL"] = old
    # (b) strong scaling: BASELINE configs[4] as written -- args.frames frames in total
    if world > 1:
        fr = args.frames // world
        cams = synth.make_cameras(args.cams, args.seed)
        rl = synth.make_rig(n_cam=args.cams, n_frame=fr, seed=args.seed, cameras=cams, frame_stream=100 + rank)
        ms, rep, err = _sharded_leg(m, dist, rl, kw, world, rank, local, barrier)
        it = max(rep["iterations"], 1)
        out["strong"] = {"workload": "%d-camera rig, %d frames in TOTAL split over %d GPUs (%d per GPU)" % (args.cams, fr * world, world, fr),
                         "us_per_lm_iteration": ms * 1e3 / it, "lm_iters_per_sec": it / (ms * 1e-3),
                         "corner_obs_per_s": world * rl["n_points"] * it / (ms * 1e-3), "rms_px": err["rms"]}
    # (c) config #4: 16-camera mixed pinhole / Mei rig, 10k frames in total
    fr4 = 10000 // world
    models = [1 if c % 2 else 0 for c in range(16)]
    cams4 = synth.make_cameras(16, 1004, models)
    r4 = synth.make_rig(n_cam=16, n_frame=fr4, seed=1004, cameras=cams4, frame_stream=rank)
    ms, rep, err = _sharded_leg(m, dist, r4, kw, world, rank, local, barrier)
    it = max(rep["iterations"], 1)
    out["config4"] = {"workload": "BASELINE configs[3]: 16-camera mixed pinhole/omnidir rig, %d frames in total over %d GPU(s)" % (fr4 * world, world),
                      "us_per_lm_iteration": ms * 1e3 / it, "lm_iters_per_sec": it / (ms * 1e-3),
                      "corner_obs_per_s": world * r4["n_points"] * it / (ms * 1e-3), "rms_px": err["rms"]}
    # (c2) the layouts off the tuned path: frames seen by three cameras (views beyond the second are not
    # TMA-staged in the Schur kernel) and a board of 130 corners per view (longer edges, fewer of them per tile of work)
    if world == 1:
        cl = {}
        for name, rk in (("views3_16cam_10k_frames", dict(n_cam=16, n_frame=10000, seed=1014, views_per_frame=3)),
                         ("board13x10_16cam_4k_frames", dict(n_cam=16, n_frame=4000, seed=1015, board_shape=(13, 10, 25.0)))):
            rc_ = synth.make_rig(**rk)
            msc, repc, errc = _sharded_leg(m, dist, rc_, kw, world, rank, local, barrier)
            itc = max(repc["iterations"], 1)
            cl[name] = {"corners": rc_["n_points"], "us_per_lm_iteration": msc * 1e3 / itc,
                        "corner_obs_per_s": rc_["n_points"] * itc / (msc * 1e-3), "rms_px": errc["rms"]}
        cl["note"] = "compare corner_obs_per_s with config4 (2 views per frame, 54 corners per view, same camera count)"
        out["other_layouts"] = cl
    # (e) the other precision policies on the headline rig (rank-local timing; the default policy is the headline itself)
    if world == 1:
        pol = {}
        for name, prec in (("fp64", m.capi.PRECISION_FP64), ("mixed", m.capi.PRECISION_MIXED), ("fast32", m.capi.PRECISION_FAST32)):
            sp = m.Solver(device=local, precision=prec)
            sp.set_cameras(rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])
            sp.set_observations(rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], pin["obj"], pin["img"])
            sp.set_parameters(pin["params_init"])
            sp.save_parameters()
            ms, rep = _time_solves(sp, kw, 5, 3, 1, barrier)
            k1 = min(sp.time_eval(reps=10) for _ in range(2))
            e = sp.reproj_error()
            pol[name] = {"resjac_kernel_us": k1 * 1e3, "resjac_gbs": 20.0 * M / (k1 * 1e-3) / 1e9,
                         "us_per_lm_iteration": ms * 1e3 / max(rep["iterations"], 1), "final_cost": rep["cost"], "rms_px": e["rms"],
                         "accepted": rep["accepted"], "rejected": rep["rejected"]}
            sp.close()
        pol["default"] = "mixed"
        pol["parity"] = ("fp64: oracle to 1e-8; mixed: cameras 1e-8, worst pattern-pose parameter of config #5 1.0e-6 "
                         "(tests/test_full_size_gpu.py); fast32: ~2e-6, outside the gate, opt-in")
        out["precision_policies"] = pol
    # (f) BASELINE configs[2]: single-camera Mei calibration (omnidir::calibrate loop), 5k frames, intrinsics + distortion
    if world == 1:
        r3 = synth.make_config(3)
        n3 = r3["n_frame"]
        pt = r3["params_true"].reshape(-1, 6)
        K5, D, xi = r3["cam_K5"][0], r3["cam_dist8"][0][:4], r3["cam_xi"][0]
        poses = np.array([pt[r3["edge_pv"][e] - 1] for e in range(n3)])
        p3 = np.concatenate([poses.ravel(), [K5[0] * 1.03, K5[1] * 1.03, K5[4], K5[2], K5[3], xi + 0.1], np.zeros(4)])
        so = m.Solver(device=local)
        so.omni_set_observations(r3["edge_off"], r3["obj"], r3["img"])
        its = 50
        t3 = []
        for _ in range(4):
            so.omni_set_parameters(p3)
            t3.append(so.omni_solve(0, 1, its, 0.0)["device_ms"])
        ms3 = float(np.median(t3[1:]))
        out["config3"] = {"workload": "BASELINE configs[2]: single Mei camera, %d frames, %d corners, 6n+10 = %d parameters; %d iterations of the omnidir::calibrate schedule" % (n3, r3["n_points"], 6 * n3 + 10, its),
                          "us_per_iteration": ms3 * 1e3 / its, "iters_per_sec": its / (ms3 * 1e-3),
                          "resjac_evals_per_sec": r3["n_points"] * its / (ms3 * 1e-3)}
        so.close()
    # (f2) the two small "next" paths of SURVEY 8(f): omnidir stereo bundle adjustment (+ uncertainties) and the double-sided
    # board calibration; per-iteration times on synthetic problems (parity: tests/test_stereo_ba.py, tests/test_double_side.py)
    if world == 1:
        try:
            rng = np.random.default_rng(9)
            cams2 = synth.make_cameras(2, 77, models=[1, 1])
            rs = synth.make_rig(n_cam=2, n_frame=400, seed=77, cameras=cams2, board_distance=(300.0, 700.0), tilt_max_deg=45.0,
                                lateral=200.0, min_depth=150.0)
            nF = rs["n_frame"]; offs = rs["edge_off"]
            E0 = np.nonzero(rs["edge_cam"] == 0)[0]; E1 = np.nonzero(rs["edge_cam"] == 1)[0]
            sl = lambda e: slice(offs[e], offs[e + 1])
            objs = np.concatenate([rs["obj"][sl(e)] for e in E0]); i1 = np.concatenate([rs["img"][sl(e)] for e in E0])
            i2 = np.concatenate([rs["img"][sl(e)] for e in E1])
            foff = np.concatenate([[0], np.cumsum([offs[e + 1] - offs[e] for e in E0])]).astype(np.int64)
            pts = rs["params_true"].reshape(-1, 6)
            intr = lambda c: np.concatenate([[rs["cam_K5"][c][0], rs["cam_K5"][c][1], rs["cam_K5"][c][4], rs["cam_K5"][c][2],
                                              rs["cam_K5"][c][3], rs["cam_xi"][c]], rs["cam_dist8"][c][:4]])
            p0s = np.concatenate([pts[0], pts[rs["edge_pv"][E0] - 1].ravel(), intr(0), intr(1)])
            p0s[:6] += np.array([0.01] * 3 + [4.0] * 3) * rng.standard_normal(6)
            ss = m.Solver(device=local)
            ss.stereo_set_observations(foff, objs, i1, i2)
            its = 100
            tt = []
            for _ in range(3):
                ss.stereo_set_parameters(p0s)
                tt.append(ss.stereo_solve(0, 1, its, 0.0)["device_ms"])
            t0 = time.perf_counter(); ss.stereo_uncertainties(0); tu = (time.perf_counter() - t0) * 1e3
            ss.close()
            out["stereo_ba"] = {"workload": "omnidir stereo pair, %d frames, %d corners per view, 6 + 6n + 20 = %d parameters; %d iterations" % (nF, int(foff[-1]), 6 + 6 * nF + 20, its),
                                "us_per_iteration": float(np.median(tt[1:])) * 1e3 / its, "uncertainties_ms": tu}
            rd = synth.make_double_side_rig(2000, seed=4002)
            sd = m.Solver(device=local)
            sd.set_rig(rd)
            sd.ds_set_problem(rd["edge_back"], rd["cam_pose"])
            itd = 20
            td = []
            for _ in range(3):
                sd.ds_set_parameters(rd["ds_params_init"])
                td.append(sd.ds_solve(1, itd, 0.0)["device_ms"])
            sd.close()
            out["double_side"] = {"workload": "3 fixed cameras, double-sided board, %d frames, %d corners, 6 + 6n = %d parameters; %d iterations" % (rd["n_frame"], rd["n_points"], 6 + 6 * rd["n_frame"], itd),
                                  "us_per_iteration": float(np.median(td[1:])) * 1e3 / itd}
        except Exception as e:                      # informative legs only
            out["stereo_ba"] = out.get("stereo_ba") or {"unavailable": str(e)[:200]}
    # (g) second end-to-end figure: the drop-in CLASS (file ingest, indexing, initialisation, optimisation, XML output) on
    # BASELINE configs[3] (16 cameras, 10k frames): MultiCameraCalibration(...).run(); writeParameters(...)
    if world == 1:
        import tempfile
        from multi_camera_calibration_b200 import multicalib, obsfile
        r4f = synth.make_config(4)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "rig.mccb")
            obsfile.write_rig(path, r4f)
            tt = []
            for _ in range(3):
                t0 = time.perf_counter()
                mc = multicalib.MultiCameraCalibration(0, 16, path, 360.0, 200.0, criteria=(1, 20, 1e-7), mode=m.capi.MODE_LM, device=local)
                mc.run()
                mc.writeParameters(os.path.join(td, "out.xml"))
                tt.append(time.perf_counter() - t0)
                st4 = mc.stats()
                mc.close()
            out["e2e_host_class"] = {"workload": "MultiCameraCalibration(16 cameras, 10k frames).run() + writeParameters(): observation file -> XML, 20 LM iterations",
                                     "seconds": float(np.median(tt[1:])), "file_mb": os.path.getsize(path) / 1e6, "rms_px": st4["rms"],
                                     "device_ms_of_the_solve": st4["device_ms"]}
    # (d) parity of the sharded path (both exchanges) against the oracle on the whole rig
    if world > 1:
        from scripts import mgpu_parity
        res = mgpu_parity.check(dist, rank, world, local, verbose=False)
        omni = mgpu_parity.check_omni(dist, rank, world, local, verbose=False)     # frame-sharded omnidir::calibrate path
        if rank == 0:
            out["parity"] = {"ok": bool(res["ok"] and omni["ok"]), "param_rel": res["param_rel"], "S_rel": res["S_rel"], "cost_rel": res["cost_rel"],
                             "cams_bit_identical": res["cams_bit_identical"], "exchange": res["exchange"],
                             "cases": [c["case"] + ":" + c["exchange"] for c in res["cases"]],
                             "omni_calibrate": {k: omni[k] for k in ("case", "iters", "oracle_iters", "param_rel", "intrinsics_bit_identical", "ok")},
                             "checker": "oracle/ on the whole rig (rank 0), tolerance 1e-6"}
    return out


def _literal_reference_toy():
    """BASELINE configs[0] substitute, part (c): the reference's LITERAL algorithm -- dense 2M x P Jacobian, dense J^T J, full
    P x P solve per iteration (src/multicalib.cpp:593-703) -- timed where it still fits: a 3-camera x 60-frame rig (the
    author's own use case: 3 serials, samples/multi_cameras_calibration.cpp:53), through the numpy re-enactment
    (oracle/dense_reenact.py, cv2 where the reference calls OpenCV).  CPU baseline leg only."""
    from multi_camera_calibration_b200 import synth
    from oracle import dense_reenact as dr
    rig = synth.make_rig(n_cam=3, n_frame=60, seed=1001)
    nC = rig["n_cam"]
    K = np.zeros((nC, 3, 3))
    for c in range(nC):
        fx, fy, cx, cy, sk = rig["cam_K5"][c]
        K[c] = [[fx, sk, cx], [0, fy, cy], [0, 0, 1]]
    dist = [rig["cam_dist8"][c][:rig["cam_ndist"][c]] for c in range(nC)]
    edges = []
    for e in range(rig["edge_cam"].size):
        a, b = rig["edge_off"][e], rig["edge_off"][e + 1]
        edges.append((int(rig["edge_cam"][e]), int(rig["edge_pv"][e]), rig["obj"][a:b], rig["img"][a:b]))
    prob = dr.RigProblem(rig["cam_model"], K, dist, rig["cam_xi"], edges, nC + rig["n_frame"])
    t0 = time.perf_counter()
    p, it, ch = dr.optimize_extrinsics(prob, rig["params_init"], 1, 5, 1e-7, policy="fp64")
    dt = (time.perf_counter() - t0) / 5
    return {"workload": "3 cameras x 60 frames, %d corners, P = %d parameters: dense J (%d x %d), dense J^T J, full solve" %
                        (rig["n_points"], rig["params_init"].size, 2 * rig["n_points"], rig["params_init"].size),
            "seconds_per_iteration": dt, "corner_obs_per_s": rig["n_points"] / dt, "kind": "literal re-enactment (numpy + cv2)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import multi_camera_calibration_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(m.capi.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    numa = _bind_to_gpu_numa(local) if world > 1 else None
    rig = _rig(args, rank)
    M = rig["n_points"]
    # pinned host copies for the end-to-end leg
    pin = {}
    for k in ("obj", "img", "params_init"):
        t = torch.from_numpy(np.ascontiguousarray(rig[k])).pin_memory()
        pin[k] = t.numpy()
    s = m.Solver(device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    s.set_cameras(rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])
    s.set_observations(rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], pin["obj"], pin["img"])
    s.set_parameters(pin["params_init"])
    s.save_parameters()
    kw = dict(mode=m.capi.MODE_LM, crit_type=m.capi.CRIT_COUNT, max_count=args.iters, lambda0=1e-3, lambda_up=10.0,
              lambda_down=1.0 / 3.0)

    # ---- device-resident leg: inputs already in HBM --------------------------------------------------------
    for _ in range(args.warmup):
        s.restore_parameters()
        rep = s.solve(**kw)
    barrier()
    if world > 1:
        s.exchange_stats()                       # reset the exchange counters: the timed region only
    sampler = ClockSampler(local) if rank == 0 else None
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    launches = 0
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.restore_parameters()
        rep = s.solve(**kw)                      # timed on the library's own stream with CUDA events (device_ms)
        dev_ms += rep["device_ms"]
        launches += rep["kernel_launches"]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if sampler else None
    xstats = None
    if world > 1:
        # per-launch %globaltimer stamps of the peer-memory exchange: posting the stores vs waiting for the last peer's words
        xs = s.exchange_stats()
        tx = torch.tensor([xs["store_us"], xs["wait_us"], xs["max_wait_us"]], dtype=torch.float64, device="cuda")
        tmax = tx.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tmin = tx.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        xstats = {"launches_per_rank": xs["launches"], "store_us_mean_max_over_ranks": float(tmax[0]),
                  "wait_us_mean_min_over_ranks": float(tmin[1]), "wait_us_mean_max_over_ranks": float(tmax[1]),
                  "wait_us_worst_launch": float(tmax[2]),
                  "note": "wait = arrival skew of the ranks + one NVLink trip; the rank that arrives last sees only the trip (min over ranks)"}
    t = torch.tensor([dev_ms / args.steps, wall_ms / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, wall_step = float(t[0]), float(t[1])
    iters_done = rep["iterations"]
    value = world * M * iters_done / (ms_step * 1e-3)

    # ---- dominant kernel alone (roofline) --------------------------------------------------------------------
    k1_ms = s.time_eval(reps=20)
    peak, peak_src = _peaks()
    k1_gbs = 20.0 * M / (k1_ms * 1e-3) / 1e9
    iter_gbs = 20.0 * M / (ms_step * 1e-3 / max(iters_done, 1)) / 1e9
    traffic = _k1_traffic()

    # ---- end-to-end leg: host buffers in, parameters out, every step -----------------------------------------
    h2d = pin["obj"].nbytes + pin["img"].nbytes + pin["params_init"].nbytes + rig["edge_cam"].nbytes + \
        rig["edge_pv"].nbytes + rig["edge_off"].nbytes
    d2h = pin["params_init"].nbytes
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    p_host = torch.empty(pin["params_init"].size, dtype=torch.float64).pin_memory().numpy()   # result lands in pinned memory

    def e2e_step():
        s.set_observations(rig["n_frame"], rig["edge_cam"], rig["edge_pv"], rig["edge_off"], pin["obj"], pin["img"])
        s.set_parameters(pin["params_init"])
        r = s.solve(**kw)
        p = s.get_parameters(out=p_host)
        return r, p

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r2, p_out = e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    e2e_value = world * M * r2["iterations"] / (e2e_ms * 1e-3)
    err = s.reproj_error()

    # ---- secondary legs (extra keys, not the headline): other BASELINE configs and solver variants ------------------
    extra = {}
    if not args.no_extra:
        extra = _extra_legs(args, m, s, rig, pin, kw, world, rank, local, barrier)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        orc, O = _oracle_rig(rig)
        cores = _cpu_threads(orc)
        ci = args.ref_iters
        ckw = dict(mode=1, crit_type=1, max_count=ci, lambda0=1e-3, lambda_up=10.0, lambda_down=1.0 / 3.0)
        O.solve(rig["params_init"], **ckw)                       # warm-up (page faults, thread pool)
        reps, cdt = 0, 0.0
        while cdt < 10.0 and reps < 50:                          # about 10 s of CPU work
            t0 = time.perf_counter()
            O.solve(rig["params_init"], **ckw)
            cdt += time.perf_counter() - t0
            reps += 1
        cpu = {"value": M * ci * reps / cdt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "full rig (%d corners), %d solves of %d LM iterations (%.1f s of CPU work)" % (M, reps, ci, cdt),
               "lm_iters_per_sec": ci * reps / cdt}
    if cpu is not None:
        try:
            cpu["literal_reference_toy"] = _literal_reference_toy()
        except Exception as e:                     # cv2 missing on the box: the toy leg is informative only
            cpu["literal_reference_toy"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": {0: "f64", 1: "f64 residual/cost/Schur/solve + packed f32 Jacobian products (MIXED policy, chosen by the default AUTO policy for this rig; parity gate 1e-6)",
                          2: "f32 (FAST32 policy)"}[s.effective_precision()],
                "data": "synthetic", "config": config_dict(args, world),
                "lm_iters_per_sec": iters_done / (ms_step * 1e-3), "us_per_lm_iteration": ms_step * 1e3 / max(iters_done, 1),
                "resjac_evals_per_sec": world * M / (k1_ms * 1e-3), "wall_ms_per_step": wall_step,
                "lm": {"iterations": iters_done, "accepted": rep["accepted"], "rejected": rep["rejected"],
                       "final_cost": rep["cost"], "rms_px": err["rms"]},
                "roofline": {"bound": "hbm", "achieved": k1_gbs, "peak": peak, "unit": "GB/s", "frac": k1_gbs / peak,
                             "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                             "kernel": {0: "resid_jac_accum_kernel (FP64)", 1: "resid_jac_accum_f32_kernel<true> (MIXED)",
                                        2: "resid_jac_accum_f32_kernel<false> (FAST32)"}[s.effective_precision()], "kernel_ms": k1_ms,
                             "algorithmic_bytes_per_launch": 20.0 * M,
                             "whole_iteration_gbs": iter_gbs, "whole_iteration_frac": iter_gbs / peak,
                             "traffic_source": "static: ncu --set full capture of this kernel (profiles/k1_traffic.json)",
                             "pipes": _pipe_block(M, k1_ms, clocks),
                             "note": "MIXED policy: residual in double, Jacobian products in packed float32; the kernel is bound by "
                                     "the FMA / FP64 pipes, not by HBM (DESIGN.md section 4)"},
                "exchange": {0: "none", 1: "ncclAllReduce", 2: "nvlink-peer-memory"}[s.exchange_mode()],
                "exchange_stats": xstats,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_ms, "steps": e2e_steps},
                "gpu_launches": int(launches), "clocks": clocks}
        if numa:
            line["host_numa_binding"] = numa     # rank 0's; every rank binds to its own GPU's node
        line.update(extra)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cams", type=int, default=64)
    ap.add_argument("--frames", type=int, default=100000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ref-iters", type=int, default=10, help="LM iterations per step of the CPU arm (bounded sample)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1005)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary legs (dense solver, strong scaling, config #4, parity)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

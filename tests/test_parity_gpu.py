"""GPU parity tests proper: the CUDA path, called through the C ABI (include/mccba.h via ctypes), against the CPU
oracle (oracle/mccba_oracle.c) on the same seeded inputs.

Tolerances (stated per assert):
  per-edge blocks, reduced system   1e-9 relative (fp64 on both sides, different formulations: tangent vs rvec)
  iterates / final parameters       1e-8 relative to max(|a|, scale) (north star gate is 1e-6)
  integer / index work              exact
"""
import numpy as np
import pytest

from tests import rigs

pytestmark = pytest.mark.gpu


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _Jl(om):
    th = np.linalg.norm(om)
    K = _skew(om)
    if th < 1e-6:
        return np.eye(3) + 0.5 * K + K @ K / 6
    return np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (th - np.sin(th)) / th ** 3 * (K @ K)


def _unpack(tri):
    H = np.zeros((6, 6))
    k = 0
    for i in range(6):
        for j in range(i, 6):
            H[i, j] = H[j, i] = tri[k]
            k += 1
    return H


def _rel(a, b, scale=None):
    s = np.abs(b).max() if scale is None else scale
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(s, 1e-300)


def _param_rel(a, b):
    """|a-b| <= tol * max(|a|, scale): scale = 1 rad for rotations, median |t| for translations."""
    a = a.reshape(-1, 6); b = b.reshape(-1, 6)
    tscale = np.median(np.linalg.norm(b[:, 3:], axis=1))
    rr = np.abs(a[:, :3] - b[:, :3]) / np.maximum(np.abs(b[:, :3]), 1.0)
    tt = np.abs(a[:, 3:] - b[:, 3:]) / np.maximum(np.abs(b[:, 3:]), tscale)
    return max(rr.max(), tt.max())


RIGS = {
    "pinhole3": dict(n_cam=3, n_frame=40, cam_models=[0, 0, 0], seed=11),
    "omni3": dict(n_cam=3, n_frame=40, cam_models=[1, 1, 1], seed=12),
    "mixed4_v3_ragged": dict(n_cam=4, n_frame=70, cam_models=[0, 1, 0, 1], seed=13, views_per_frame=3, ragged=True),
    "rational2": dict(n_cam=2, n_frame=33, cam_models=[0, 0], seed=14, ndist=8),
    "pinhole8": dict(n_cam=8, n_frame=200, cam_models=[0] * 8, seed=15),
    "nodist5": dict(n_cam=5, n_frame=64, cam_models=[0] * 5, seed=16, ndist=4),
    # 130 corners per image: a 32-edge chunk no longer fits a TMA stage -> observations read straight from global memory
    "bigboard3": dict(n_cam=3, n_frame=70, cam_models=[0, 1, 0], seed=17, nx=13, ny=10),
}


@pytest.fixture(scope="module")
def solver():
    import multi_camera_calibration_b200 as m
    s = m.Solver(device=0, precision=m.capi.PRECISION_FP64)       # parity at rounding level: the fp64 policy (MIXED: tests/test_precision_gpu.py)
    yield s
    s.close()


@pytest.fixture(scope="module", params=sorted(RIGS))
def case(request, oracle_lib):
    rig = rigs.make_rig(**RIGS[request.param])
    return request.param, rig, rigs.to_oracle_rig(rig)


def test_eval_blocks(solver, case, oracle_lib):
    name, rig, O = case
    p = rig["params_init"]
    solver.set_rig(rig)
    solver.set_parameters(p)
    out = solver.eval()
    cost = O.eval(p)
    assert abs(out["cost"] - cost) <= 1e-11 * cost
    assert _rel(out["edge_cost"], O.blocks(7)[:, 0]) < 1e-11
    H6o = O.blocks(0).reshape(-1, 6, 6)
    g6o = O.blocks(1)
    nC = rig["n_cam"]
    for e in range(rig["edge_cam"].size):
        c, pv = int(rig["edge_cam"][e]), int(rig["edge_pv"][e])
        omP, tP = p[6 * (pv - 1):6 * (pv - 1) + 3], p[6 * (pv - 1) + 3:6 * pv]
        omC, tC = (p[6 * (c - 1):6 * (c - 1) + 3], p[6 * (c - 1) + 3:6 * c]) if c > 0 else (np.zeros(3), np.zeros(3))
        om3, _, _ = oracle_lib.compose_motion(omP, tP, omC, tC)
        D = np.eye(6); D[:3, :3] = _Jl(om3)          # phi3 = J_l(om3) d om3
        Ht = _unpack(out["H6"][e])
        assert _rel(D.T @ Ht @ D, H6o[e]) < 1e-9, (name, e)
        assert _rel(D.T @ out["g6"][e], g6o[e]) < 1e-9, (name, e)


@pytest.mark.parametrize("lam", [1e-3, 0.7])
def test_reduced_system(solver, case, lam):
    name, rig, O = case
    p = rig["params_init"]
    solver.set_rig(rig)
    solver.set_parameters(p)
    S, gs = solver.reduced_system(lam)
    O.eval(p)
    rc, step, So, gso = O.solve_normal(p, lam)
    assert rc == 0
    assert _rel(S, So) < 1e-9
    assert _rel(gs, gso) < 1e-9
    assert _rel(S, S.T) < 1e-12


def test_reference_schedule_iterates(solver, case):
    """Reference step-scaled Gauss-Newton (src/multicalib.cpp:473-507): same iteration count, same iterates."""
    name, rig, O = case
    p0 = rig["params_init"]
    solver.set_rig(rig)
    for k in (1, 2, 5):
        solver.set_parameters(p0)
        rep = solver.solve(mode=0, crit_type=1, max_count=k)
        ref = O.solve(p0, mode=0, crit_type=1, max_count=k, trace_cap=k)
        assert rep["iterations"] == k == ref["iters"]
        assert _param_rel(solver.get_parameters(), ref["params"]) < 1e-8, (name, k)
        assert abs(rep["change"] - ref["change"]) <= 1e-6 * ref["change"]
        assert abs(rep["cost"] - ref["cost"]) <= 1e-8 * ref["cost"]


def test_reference_schedule_converged(solver, case):
    """TermCriteria(COUNT+EPS, 200, 1e-7), the MyMultiCameraCalibration default (mymulticalib.hpp:95)."""
    name, rig, O = case
    p0 = rig["params_init"]
    solver.set_rig(rig)
    solver.set_parameters(p0)
    rep = solver.solve(mode=0, crit_type=3, max_count=200, eps=1e-7)
    ref = O.solve(p0, mode=0, crit_type=3, max_count=200, eps=1e-7)
    assert rep["iterations"] == ref["iters"], (rep, ref["iters"])
    pg = solver.get_parameters()
    assert _param_rel(pg, ref["params"]) < 1e-8
    err = solver.reproj_error()
    eo = O.error(ref["params"])
    assert abs(err["rms"] - eo["rms"]) <= 1e-9 * eo["rms"]
    assert abs(err["mean_reproj_error"] - eo["mean_reproj_error"]) <= 1e-9 * eo["mean_reproj_error"]
    assert _rel(err["per_edge"], eo["per_edge"]) < 1e-8
    assert err["n_points"] == eo["n_points"]


def test_lm(solver, case):
    """LM with a fixed iteration count: every accept/reject decision is far from the rounding floor, so the
    device-side sequence and the iterates must equal the oracle's."""
    name, rig, O = case
    p0 = rig["params_init"]
    solver.set_rig(rig)
    kw = dict(lambda0=1e-3, lambda_up=10.0, lambda_down=1.0 / 3.0)
    for k in (1, 3, 6):
        solver.set_parameters(p0)
        rep = solver.solve(mode=1, crit_type=1, max_count=k, **kw)
        ref = O.solve(p0, mode=1, crit_type=1, max_count=k, trace_cap=k, **kw)
        acc = int(ref["trace"][:, 3].sum())
        assert rep["iterations"] == ref["iters"] == k
        assert rep["accepted"] == acc and rep["rejected"] == k - acc
        assert _param_rel(solver.get_parameters(), ref["params"]) < 1e-8
        assert abs(rep["cost"] - ref["cost"]) <= 1e-9 * ref["cost"]
        assert abs(rep["lam"] - ref["lam"]) <= 1e-12 * ref["lam"]


def test_lm_converged(solver, case):
    """LM run to convergence (COUNT+EPS).  Near the optimum the cost differences fall below rounding, so the
    accept/reject tail (and hence the iteration count) is not comparable; the converged parameters are: within the
    north star's 1e-6."""
    name, rig, O = case
    p0 = rig["params_init"]
    solver.set_rig(rig)
    solver.set_parameters(p0)
    rep = solver.solve(mode=1, crit_type=3, max_count=60, eps=1e-7)
    ref = O.solve(p0, mode=1, crit_type=3, max_count=60, eps=1e-7)
    assert rep["status"] == 0 and ref["status"] == 0
    assert _param_rel(solver.get_parameters(), ref["params"]) < 1e-6
    assert abs(rep["cost"] - ref["cost"]) <= 1e-9 * ref["cost"]
    # same minimum as the reference schedule (LM stops at change <= 1e-7, i.e. still ~1e-6 away in ill-conditioned
    # directions, so compare costs and leave a loose bound on the parameters)
    gn = O.solve(p0, mode=0, crit_type=3, max_count=200, eps=1e-9)
    assert abs(rep["cost"] - gn["cost"]) <= 1e-7 * gn["cost"]
    assert _param_rel(solver.get_parameters(), gn["params"]) < 1e-4


def test_lm_rejections(solver, oracle_lib):
    """A badly perturbed start (0.3 rad / 200 mm) makes LM reject steps while the costs are still far apart
    (oracle sequence 1 1 1 0 0 0 0 0 1 1 0 0 1 1): the device-side accept/reject decisions, the damping and the
    iterates after a rebuild must equal the oracle's."""
    rig = rigs.make_rig(n_cam=3, n_frame=30, cam_models=[0, 0, 0], seed=24, init_rot=0.3, init_trans=200.0)
    O = rigs.to_oracle_rig(rig)
    p0 = rig["params_init"]
    solver.set_rig(rig)
    kw = dict(lambda0=1e-6, lambda_up=10.0, lambda_down=0.1)
    full = O.solve(p0, mode=1, crit_type=1, max_count=14, trace_cap=14, **kw)
    assert full["trace"][:, 3].astype(int).tolist() == [1, 1, 1, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 1]
    for k in (3, 4, 8, 9, 12, 14):
        solver.set_parameters(p0)
        rep = solver.solve(mode=1, crit_type=1, max_count=k, **kw)
        ref = O.solve(p0, mode=1, crit_type=1, max_count=k, trace_cap=k, **kw)
        acc = int(ref["trace"][:, 3].sum())
        assert (rep["iterations"], rep["accepted"], rep["rejected"]) == (k, acc, k - acc), (k, rep)
        assert abs(rep["lam"] - ref["lam"]) <= 1e-12 * ref["lam"]
        assert abs(rep["cost"] - ref["cost"]) <= 1e-7 * ref["cost"], k
        assert _param_rel(solver.get_parameters(), ref["params"]) < 1e-7, k


@pytest.mark.parametrize("chol", ["2", "3"])
def test_every_reduced_solver_matches(case, chol, monkeypatch):
    """The reduced camera system has two solvers (MCCBA_CHOL, read when the observations are set): 2 the one-launch tile DAG
    (dense camera graphs), 3 the block cyclic reduction (the default when the camera graph is block-banded, which every
    small rig here is; "3" = leave the choice to the library).  Both must reproduce the oracle's LM iterates."""
    import multi_camera_calibration_b200 as m
    name, rig, O = case
    monkeypatch.setenv("MCCBA_CHOL", chol)
    s = m.Solver(device=0, precision=m.capi.PRECISION_FP64)
    try:
        s.set_rig(rig)
        s.set_parameters(rig["params_init"])
        rep = s.solve(mode=1, crit_type=1, max_count=5)
        ref = O.solve(rig["params_init"], mode=1, crit_type=1, max_count=5)
        assert rep["iterations"] == 5 == ref["iters"]
        assert _param_rel(s.get_parameters(), ref["params"]) < 1e-8
        assert abs(rep["cost"] - ref["cost"]) <= 1e-9 * ref["cost"]
    finally:
        s.close()


def test_no_graph_path_matches(case):
    import multi_camera_calibration_b200 as m
    name, rig, O = case
    s = m.Solver(device=0, use_graph=False, precision=m.capi.PRECISION_FP64)
    try:
        s.set_rig(rig)
        s.set_parameters(rig["params_init"])
        rep = s.solve(mode=0, crit_type=1, max_count=4)
        ref = O.solve(rig["params_init"], mode=0, crit_type=1, max_count=4)
        assert rep["iterations"] == 4
        assert _param_rel(s.get_parameters(), ref["params"]) < 1e-8
    finally:
        s.close()


def test_bad_arguments(solver):
    import multi_camera_calibration_b200 as m
    rig = rigs.make_rig(n_cam=2, n_frame=6, seed=3)
    solver.set_cameras(rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])
    bad_pv = rig["edge_pv"].copy(); bad_pv[0] = 0
    with pytest.raises(m.MccbaError) as ei:
        solver.set_observations(rig["n_frame"], rig["edge_cam"], bad_pv, rig["edge_off"], rig["obj"], rig["img"])
    assert ei.value.code == m.capi.ERR_ARG
    with pytest.raises(m.MccbaError) as ei:
        solver.set_cameras([1], [[600, 600, 900, 500, 0]], np.zeros((1, 8)), [5], [1.0])   # Mei needs 4 coefficients
    assert ei.value.code == m.capi.ERR_ARG
    solver.set_rig(rig)
    with pytest.raises(m.MccbaError) as ei:
        solver.set_parameters(np.zeros(5))
    assert ei.value.code == m.capi.ERR_ARG

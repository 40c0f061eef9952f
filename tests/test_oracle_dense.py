"""Tier 2: the block-sparse / Schur C oracle computes the same iterates as the dense literal re-enactment of the
reference loop (committed golden iterates, and live when cv2 is importable)."""
import os

import numpy as np
import pytest

from tests import rigs

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "rig_dense.npz"))
CASES = {"pin": dict(n_cam=3, n_frame=8, cam_models=[0, 0, 0], seed=41),
         "omni": dict(n_cam=2, n_frame=7, cam_models=[1, 1], seed=42),
         "mixed": dict(n_cam=3, n_frame=9, cam_models=[0, 1, 0], seed=43, views_per_frame=3, ragged=True)}


@pytest.mark.parametrize("name", sorted(CASES))
def test_first_step_equals_dense_solve(oracle_lib, name):
    rig = rigs.make_rig(**CASES[name])
    O = rigs.to_oracle_rig(rig)
    O.eval(rig["params_init"])
    rc, step, S, gs = O.solve_normal(rig["params_init"], 0.0)
    assert rc == 0
    x0 = G[name + "_x0"]
    assert np.abs(step - x0).max() <= 1e-9 * np.abs(x0).max()


@pytest.mark.parametrize("name", sorted(CASES))
def test_fp64_iterates(oracle_lib, name):
    rig = rigs.make_rig(**CASES[name])
    O = rigs.to_oracle_rig(rig)
    it = int(G[name + "_fp64_iters"])
    r = O.solve(rig["params_init"], mode=0, crit_type=3, max_count=12, eps=1e-7, policy=0, trace_cap=12)
    assert r["iters"] == it
    ref = G[name + "_fp64_params"]
    assert np.abs(r["params"] - ref).max() <= 1e-10 * np.abs(ref).max()
    tr = G[name + "_fp64_trace"]
    assert np.abs(r["trace"][:, 0] - tr[:, 0]).max() <= 1e-9 * tr[:, 0].max()
    assert np.abs(r["trace"][:, 1] - tr[:, 1]).max() <= 1e-6 * tr[:, 1].max()
    for k in (1, 2):      # intermediate iterates too
        rk = O.solve(rig["params_init"], mode=0, crit_type=1, max_count=k)
        assert np.abs(rk["params"] - G[name + "_fp64_iterates"][k - 1]).max() <= 1e-10 * np.abs(ref).max()
    e = O.error(r["params"], 0)
    ge = G[name + "_fp64_err"]
    assert abs(e["mean_reproj_error"] - ge[0]) <= 1e-10 * ge[0]
    assert abs(e["rms"] - ge[1]) <= 1e-10 * ge[1]
    assert e["n_points"] == int(ge[2])
    assert np.abs(e["per_edge"] - G[name + "_fp64_per_edge"]).max() < 1e-9


@pytest.mark.parametrize("name", sorted(CASES))
def test_faithful_f32_iterates_bit_exact_enough(oracle_lib, name):
    """The reference's float32 round trips.  Parameters live in float32, so agreement with the
    dense re-enactment is expected to within one float32 ulp; the drift against fp64 is reported, not asserted."""
    rig = rigs.make_rig(**CASES[name])
    O = rigs.to_oracle_rig(rig)
    r = O.solve(rig["params_init"], mode=0, crit_type=3, max_count=12, eps=1e-7, policy=1)
    ref = G[name + "_faithful_f32_params"]
    assert r["iters"] == int(G[name + "_faithful_f32_iters"])
    ulp = np.maximum(np.abs(ref), 1e-3) * 2.0 ** -22
    assert np.all(np.abs(r["params"] - ref) <= ulp)
    drift = np.abs(ref - G[name + "_fp64_params"]).max() / np.abs(ref).max()
    assert drift < 5e-6


def test_live_dense_reenactment(oracle_lib):
    pytest.importorskip("cv2")
    from oracle import dense_reenact as dr
    rig = rigs.make_rig(n_cam=3, n_frame=6, cam_models=[1, 0, 0], seed=77)
    P = rigs.to_dense_problem(rig)
    O = rigs.to_oracle_rig(rig)
    pd, itd, chd = dr.optimize_extrinsics(P, rig["params_init"], 1, 3, 1e-7, "fp64")
    r = O.solve(rig["params_init"], mode=0, crit_type=1, max_count=3)
    assert np.abs(pd - r["params"]).max() <= 1e-10 * np.abs(pd).max()


def test_lm_reaches_the_same_minimum(oracle_lib):
    rig = rigs.make_rig(n_cam=3, n_frame=10, seed=9)
    O = rigs.to_oracle_rig(rig)
    gn = O.solve(rig["params_init"], mode=0, crit_type=3, max_count=200, eps=1e-9)
    lm = O.solve(rig["params_init"], mode=1, crit_type=3, max_count=100, eps=1e-9)
    assert abs(gn["cost"] - lm["cost"]) <= 1e-8 * gn["cost"]
    assert lm["cost"] < O.eval(rig["params_true"])     # the optimum fits the noisy data better than the truth


def test_empty_and_ragged_edges(oracle_lib):
    rig = rigs.make_rig(n_cam=3, n_frame=9, seed=4, ragged=True)
    O = rigs.to_oracle_rig(rig)
    c = O.eval(rig["params_init"])
    assert np.isclose(O.blocks(7).sum(), c)
    bad = rig["edge_pv"].copy(); bad[0] = 1
    with pytest.raises(ValueError):
        oracle_lib.Rig(rig["n_cam"], rig["n_frame"], rig["edge_cam"], bad, rig["edge_off"], rig["obj"], rig["img"],
                       rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"])

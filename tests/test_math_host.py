"""The product's device arithmetic (multi_camera_calibration_b200/csrc/mccba_math.cuh, __host__ __device__) compiled
for the host and re-enacted sequentially, against the oracle.  Two independent formulations meet here: the oracle
differentiates through compose_motion in Rodrigues-vector coordinates like the reference, the product works in
left-perturbation tangent coordinates."""
import numpy as np
import pytest

from tests import rigs
from tests import harness

CASES = [dict(n_cam=3, n_frame=12, cam_models=[0, 0, 0], seed=5),
         dict(n_cam=3, n_frame=12, cam_models=[1, 1, 1], seed=6),
         dict(n_cam=4, n_frame=14, cam_models=[0, 1, 0, 1], seed=7, ndist=8, views_per_frame=3, ragged=True),
         dict(n_cam=2, n_frame=9, cam_models=[0, 0], seed=8, ndist=4)]


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("lam", [0.0, 1e-3, 0.5])
def test_step_and_reduced_system(oracle_lib, kw, lam):
    rig = rigs.make_rig(**kw)
    O = rigs.to_oracle_rig(rig)
    p = rig["params_init"]
    out = harness.rig_step(rig, p, lam)
    cost = O.eval(p)
    rc, step, S, gs = O.solve_normal(p, lam)
    assert out["bad"] == 0 and rc == 0
    assert abs(out["blocks"][:, 27].sum() - cost) <= 1e-12 * cost
    assert np.abs(out["step"] - step).max() <= 1e-9 * np.abs(step).max()
    if lam > 0:     # for lam == 0 the oracle's S is in Rodrigues coordinates; the step is coordinate-free
        assert np.abs(out["S"] - S).max() <= 1e-10 * np.abs(S).max()
        assert np.abs(out["gs"] - gs).max() <= 1e-10 * np.abs(gs).max()


def test_rodrigues_and_left_jacobian_inverse(oracle_lib):
    rng = np.random.default_rng(3)
    for sc in (1e-9, 1e-5, 1e-3, 0.2, 0.26, 1.0, 2.5):
        om = rng.standard_normal(3); om *= sc / np.linalg.norm(om)
        R, J = oracle_lib.rodrigues(om)
        assert np.abs(harness.rodrigues(om) - R).max() < 1e-14
        # J_l^-1 psi: d exp(om)/d om [delta] = [J_l delta]_x R, so for psi = J_l delta the map must return delta
        delta = rng.standard_normal(3)
        K = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
        th = np.linalg.norm(om)
        Jl = np.eye(3) + ((1 - np.cos(th)) / th ** 2 if th > 1e-4 else 0.5) * K + \
            ((th - np.sin(th)) / th ** 3 if th > 1e-4 else 1 / 6) * (K @ K)
        back = harness.left_jacobian_inv_apply(om, Jl @ delta)
        assert np.abs(back - delta).max() < 1e-7 * max(1.0, 1.0)

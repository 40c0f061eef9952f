"""Oracle (oracle/mccba_oracle.c) against the committed golden vectors: tier 0 (OpenCV 4.13 via cv2) and tier 1
(transcription of src/omnidir.cpp:126-243), plus finite-difference checks of both camera-model Jacobians."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(G, "cv2_primitives.npz"))


def test_rodrigues_matches_cv2(oracle_lib, prim):
    for om, R, J in zip(prim["rod_om"], prim["rod_R"], prim["rod_J"]):
        Ro, Jo = oracle_lib.rodrigues(om)
        assert np.abs(Ro - R).max() < 1e-14
        # OpenCV's own Jacobian loses digits to cancellation below ~1e-6 rad; 1e-9 absolute covers it
        assert np.abs(Jo - J).max() < (1e-9 if np.linalg.norm(om) < 1e-5 else 1e-13)
        assert np.abs(oracle_lib.rodrigues_inv(R) - om).max() < 1e-12


def test_rodrigues_small_and_near_pi(oracle_lib):
    for th in (0.0, 1e-12, 1e-9, np.pi - 1e-4):
        om = np.array([0.3, -0.5, 0.81]); om = om / np.linalg.norm(om) * th
        R, _ = oracle_lib.rodrigues(om)
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-14
        assert np.abs(oracle_lib.rodrigues_inv(R) - om).max() < 1e-7


def test_compose_motion_matches_composeRT(oracle_lib, prim):
    for vin, vout in zip(prim["compose_in"], prim["compose_out"]):
        om3, T3, d = oracle_lib.compose_motion(vin[0:3], vin[3:6], vin[6:9], vin[9:12])
        assert np.abs(om3 - vout[0:3]).max() < 1e-13
        assert np.abs(T3 - vout[3:6]).max() < 1e-10
        ref = vout[6:].reshape(8, 3, 3)     # dr3dr1 dr3dt1 dr3dr2 dr3dt2 dt3dr1 dt3dt1 dt3dr2 dt3dt2
        for i in range(8):
            assert np.abs(d[i] - ref[i]).max() <= 1e-12 * max(1.0, np.abs(ref[i]).max()), i


@pytest.mark.parametrize("nd", [0, 4, 5, 8])
def test_pinhole_matches_projectPoints(oracle_lib, prim, nd):
    p, j = oracle_lib.project_pinhole(prim["pin_obj"], prim["pin_om"], prim["pin_T"], prim["pin_K5"], prim["pin_dist8"][:nd])
    assert np.abs(p - prim["pin_proj_%d" % nd]).max() < 1e-10
    ref = prim["pin_jac_%d" % nd]
    assert np.abs(j - ref).max() <= 1e-12 * np.abs(ref).max()


def test_pinhole_unsupported_length(oracle_lib, prim):
    with pytest.raises(ValueError):
        oracle_lib.project_pinhole(prim["pin_obj"], prim["pin_om"], prim["pin_T"], prim["pin_K5"], np.zeros(12))


def test_camodocal_known_answers(oracle_lib, prim):
    """camodocal/PinholeCamera_test.cc:10-23, 65-85: P=(0,0,1) -> (cx,cy); P=(1,-1,4) -> (538.70955939, 75.58850928)."""
    p, _ = oracle_lib.project_pinhole(np.array([[0.0, 0, 1], [1.0, -1, 4]]), np.zeros(3), np.zeros(3), prim["camodocal_K5"],
                                      prim["camodocal_D"], want_jac=False)
    assert np.abs(p[0] - prim["camodocal_K5"][2:4]).max() < 1e-10
    assert np.abs(p[1] - np.array([538.70955939, 75.58850928])).max() < 1e-7
    assert np.abs(p - prim["camodocal_uv"]).max() < 1e-10


def test_omnidir_matches_transcription(oracle_lib):
    g = np.load(os.path.join(G, "omni_points.npz"))
    p, j = oracle_lib.project_omnidir(g["obj"], g["om"], g["T"], g["K5"], float(g["xi"]), g["D"])
    assert np.abs(p - g["proj"]).max() < 1e-10
    assert np.abs(j - g["jac"]).max() <= 1e-12 * np.abs(g["jac"]).max()


def test_omnidir_jacobian_finite_difference(oracle_lib):
    """The author's verification method (commented-out FD checks, src/multicalib.cpp:644-668), with a threshold."""
    g = np.load(os.path.join(G, "omni_points.npz"))
    base = np.concatenate([g["om"], g["T"], g["K5"][[0, 1, 4, 2, 3]], [float(g["xi"])], g["D"]])  # jac column order

    def f(v):
        K5 = np.array([v[6], v[7], v[9], v[10], v[8]])
        return oracle_lib.project_omnidir(g["obj"], v[0:3], v[3:6], K5, v[11], v[12:16], want_jac=False)[0].reshape(-1)

    _, jac = oracle_lib.project_omnidir(g["obj"], g["om"], g["T"], g["K5"], float(g["xi"]), g["D"])
    for c in range(16):
        h = 1e-6 * max(1.0, abs(base[c]))
        vp = base.copy(); vp[c] += h
        vm = base.copy(); vm[c] -= h
        fd = (f(vp) - f(vm)) / (2 * h)
        assert np.abs(fd - jac[:, c]).max() <= 2e-6 * max(1.0, np.abs(jac[:, c]).max()), c


def test_pinhole_jacobian_finite_difference(oracle_lib, prim):
    base = np.concatenate([prim["pin_om"], prim["pin_T"]])
    f = lambda v: oracle_lib.project_pinhole(prim["pin_obj"], v[:3], v[3:], prim["pin_K5"], prim["pin_dist8"], False)[0].reshape(-1)
    _, jac = oracle_lib.project_pinhole(prim["pin_obj"], prim["pin_om"], prim["pin_T"], prim["pin_K5"], prim["pin_dist8"])
    for c in range(6):
        h = 1e-6 * max(1.0, abs(base[c]))
        vp = base.copy(); vp[c] += h
        vm = base.copy(); vm[c] -= h
        fd = (f(vp) - f(vm)) / (2 * h)
        assert np.abs(fd - jac[:, c]).max() <= 2e-6 * max(1.0, np.abs(jac[:, c]).max()), c


def test_live_cv2_agreement(oracle_lib):
    """Same checks against the cv2 in this image, if present (guards against a stale fixture)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    om = rng.standard_normal(3) * 0.4
    R, J = cv2.Rodrigues(om)
    Ro, Jo = oracle_lib.rodrigues(om)
    assert np.abs(Ro - R).max() < 1e-14 and np.abs(Jo - J).max() < 1e-13

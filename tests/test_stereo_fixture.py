"""BASELINE configs[0] substitute, part (b): the only real multi-camera data the reference
ships -- tutorials/data/omni_stereocalib_data.xml, two 704x576 omnidirectional cameras x 39 frames x 48 corners -- taken
through the reference's own sequence: per-camera omnidir::calibrate, spanning-tree initialisation, 20 iterations of the
rig loop.  Golden vectors: tests/golden/stereo_rig_fixture.npz, written by tests/golden/make_golden.py with the literal
dense re-enactment; they reproduce the survey's independent numbers to every printed digit (asserted below):
cam0 37/39 frames, RMS 0.446242, xi 5.7313; cam1 36/39, RMS 0.406593, xi 1.5353; rig 40 vertices, 73 edges, 234
parameters, 3504 corners, RMS 0.540273 -> 0.456979 -> 0.456756 -> 0.456754, mean |e| 0.390572, camera-1 pose
rvec (-0.054500, -0.062763, 0.110670), tvec (-158.800, -19.662, -5.356)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "stereo_rig_fixture.npz"))


def test_golden_reproduces_the_survey_numbers():
    assert G["cam0_idx"].size == 37 and G["cam1_idx"].size == 36
    assert abs(float(G["cam0_rms"]) - 0.446242) < 1e-6 and abs(float(G["cam1_rms"]) - 0.406593) < 1e-6
    assert abs(G["cam0_params"][6 * 37 + 5] - 5.7313) < 1e-4 and abs(G["cam1_params"][6 * 36 + 5] - 1.5353) < 1e-4
    assert abs(G["cam0_params"][6 * 37] - 1626.52) < 1e-2 and abs(G["cam1_params"][6 * 36 + 1] - 649.80) < 1e-2
    assert G["rig_vertex_timestamp"].size == 40 and G["rig_edge_cam"].size == 73 and G["rig_p_init"].size == 234
    assert int(G["rig_edge_off"][-1]) == 3504
    assert abs(float(G["rig_rms_init"]) - 0.540273) < 1e-5        # start: float32 poses, printed to 6 digits by the survey
    assert np.allclose(G["rig_rms_seq"][:3], [0.456979, 0.456756, 0.456754], atol=1e-6)
    assert abs(float(G["rig_rms"]) - 0.456754) < 1e-6 and abs(float(G["rig_mean_error"]) - 0.390572) < 1e-6
    assert np.allclose(G["rig_p_final"][:3], [-0.054500, -0.062763, 0.110670], atol=1e-6)
    assert np.allclose(G["rig_p_final"][3:6], [-158.800, -19.662, -5.356], atol=1e-3)


def _oracle_rig(oracle_lib):
    nC, nV = 2, G["rig_vertex_timestamp"].size
    K5 = np.array([[G["rig_K"][c][0, 0], G["rig_K"][c][1, 1], G["rig_K"][c][0, 2], G["rig_K"][c][1, 2], G["rig_K"][c][0, 1]] for c in range(nC)])
    d8 = np.zeros((nC, 8)); d8[:, :4] = G["rig_dist"]
    args = (nC, nV - nC, G["rig_edge_cam"], G["rig_edge_pv"], G["rig_edge_off"], G["rig_obj"], G["rig_img"],
            np.ones(nC, dtype=np.int32), K5, d8, np.full(nC, 4, dtype=np.int32), G["rig_xi"])
    return oracle_lib.Rig(*args), args


def test_oracle_per_camera_calibration(oracle_lib):
    for c, n in ((0, 37), (1, 36)):
        off, obj, img = G["cam%d_off" % c], G["cam%d_obj" % c].astype(np.float64), G["cam%d_img" % c].astype(np.float64)
        r = oracle_lib.omni_solve(off, obj, img, G["cam%d_p0" % c], 0, 3, 300, 1e-7, dense=True)
        ref = G["cam%d_params" % c]
        assert r["iters"] == int(G["cam%d_iters" % c]) == 300
        # cam0 is the xi / focal-length degenerate lens of the survey (hits the iteration cap): 1e-5; cam1: 1e-6
        tol = 1e-5 if c == 0 else 1e-6
        assert np.max(np.abs(r["params"] - ref) / np.maximum(np.abs(ref), 1.0)) < tol
        assert abs(r["rms"] - float(G["cam%d_rms" % c])) < 1e-7


def test_oracle_rig_loop(oracle_lib):
    O, _ = _oracle_rig(oracle_lib)
    ref = O.solve(G["rig_p_init"], mode=0, crit_type=1, max_count=20)
    assert ref["iters"] == 20
    assert np.max(np.abs(ref["params"] - G["rig_p_final"]) / np.maximum(np.abs(G["rig_p_final"]), 1.0)) < 1e-7
    e = O.error(ref["params"])
    assert abs(e["rms"] - float(G["rig_rms"])) < 1e-8 and abs(e["rms"] - 0.456754) < 1e-6
    assert abs(e["mean_reproj_error"] - float(G["rig_mean_error"])) < 1e-8        # OMNIDIRECTIONAL: N points per edge
    r3 = O.solve(G["rig_p_init"], mode=0, crit_type=1, max_count=3)
    assert np.max(np.abs(r3["params"] - G["rig_iter3"]) / np.maximum(np.abs(G["rig_iter3"]), 1.0)) < 1e-8


@pytest.mark.gpu
def test_gpu_per_camera_calibration():
    import multi_camera_calibration_b200 as m
    s = m.Solver(device=0)
    for c, n in ((0, 37), (1, 36)):
        s.omni_set_observations(G["cam%d_off" % c], G["cam%d_obj" % c], G["cam%d_img" % c])
        s.omni_set_parameters(G["cam%d_p0" % c])
        rep = s.omni_solve(0, 3, 300, 1e-7)
        p = s.omni_get_parameters()
        ref = G["cam%d_params" % c]
        assert rep["iterations"] == 300
        tol = 1e-5 if c == 0 else 1e-6
        assert np.max(np.abs(p - ref) / np.maximum(np.abs(ref), 1.0)) < tol
        assert abs(rep["rms"] - float(G["cam%d_rms" % c])) < 1e-7
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
def test_gpu_rig_loop(oracle_lib, precision):
    """The rig loop on the GPU (both precision policies), 20 iterations of the reference schedule, against the golden
    vectors of the literal re-enactment: parameters 1e-6 (the gate), RMS, mean error, camera-1 pose of the survey."""
    import multi_camera_calibration_b200 as m
    _, args = _oracle_rig(oracle_lib)
    nC, nF, ec, ep, eo, obj, img, model, K5, d8, nd, xi = args
    s = m.Solver(device=0, precision=precision)
    s.set_cameras(model, K5, d8, nd, xi)
    s.set_observations(nF, ec, ep, eo, obj, img)
    s.set_parameters(G["rig_p_init"])
    rep = s.solve(mode=0, crit_type=1, max_count=20)
    p = s.get_parameters()
    e = s.reproj_error()
    s.close()
    ref = G["rig_p_final"]
    assert rep["iterations"] == 20
    assert np.max(np.abs(p - ref) / np.maximum(np.abs(ref), 1.0)) < (1e-8 if precision == 0 else 1e-6)
    assert abs(e["rms"] - 0.456754) < 1e-6 and abs(e["rms"] - float(G["rig_rms"])) < 1e-8
    assert abs(e["mean_reproj_error"] - 0.390572) < 1e-6
    assert np.allclose(p[3:6], [-158.800, -19.662, -5.356], atol=1e-3)

"""Host-side mirror of the reference class (include/mccba_host.hpp): indexing, initialisation, XML output on the CPU;
the full run() / outlier loop on the GPU."""
import os

import numpy as np
import pytest

from multi_camera_calibration_b200 import multicalib, obsfile, synth
from oracle import indexing
from tests import rigs


def _rig_file(tmp_path, **kw):
    rig = synth.make_rig(**kw)
    path = str(tmp_path / "rig.mccb")
    return rig, path


def test_indexing_is_bit_exact_with_the_reference_rules(tmp_path):
    rig, path = _rig_file(tmp_path, n_cam=4, n_frame=30, seed=9)
    # add two single-view images (must be dropped by the multi-camera filter) and out-of-order timestamps
    B = synth.board().astype(np.float32)
    extra = [dict(camera=1, timestamp=900, transform=np.eye(4), obj=B, img=B[:, :2] + 500),
             dict(camera=3, timestamp=7777, transform=np.eye(4), obj=B, img=B[:, :2] + 300)]
    images = obsfile.write_rig(path, rig, seed=1, extra_images=extra)
    mc = multicalib.MultiCameraCalibration(multicalib.PINHOLE, 4, path, 360.0, 200.0)
    mc.loadImages()
    ix = mc.indexing()
    files = [[] for _ in range(4)]
    for im in images:
        files[im["camera"]].append(("cam%d/%012d.yaml" % (im["camera"], im["timestamp"]), im["timestamp"]))
    ref = indexing.build_indexing(4, files)
    assert ix["edge_cam"].tolist() == [e[0] for e in ref["edges"]]
    assert ix["edge_pv"].tolist() == [e[1] for e in ref["edges"]]
    assert ix["photo_index"].tolist() == [e[2] for e in ref["edges"]]
    assert ix["vertex_timestamp"].tolist() == ref["vertex_timestamp"]
    assert 900 not in ix["vertex_timestamp"] and 7777 not in ix["vertex_timestamp"]
    mc.close()


def test_initialize_chains_poses_close_to_truth(tmp_path):
    rig, path = _rig_file(tmp_path, n_cam=5, n_frame=40, seed=10)
    obsfile.write_rig(path, rig, seed=2)
    mc = multicalib.MultiCameraCalibration(0, 5, path, 360.0, 200.0)
    mc.loadImages()
    mc.initialize()
    p0 = mc.initialParameters()
    assert p0.size == rig["params_true"].size
    assert np.array_equal(p0, p0.astype(np.float32).astype(np.float64))        # CV_32F parameter vector
    a, b = p0.reshape(-1, 6), rig["params_true"].reshape(-1, 6)
    Ra, Rb = synth.rodrigues_batch(a[:, :3]), synth.rodrigues_batch(b[:, :3])
    ang = np.arccos(np.clip((np.einsum("nij,nij->n", Ra, Rb) - 1) / 2, -1, 1))   # rotation angle of Ra^T Rb
    assert ang.max() < 0.1 and np.abs(a[:, 3:] - b[:, 3:]).max() < 80            # PnP-like noise chained along the tree
    mc.close()


def test_write_parameters_is_opencv_filestorage(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rig, path = _rig_file(tmp_path, n_cam=3, n_frame=8, seed=11, models=[1, 1, 1])
    obsfile.write_rig(path, rig, seed=3)
    mc = multicalib.MultiCameraCalibration(multicalib.OMNIDIRECTIONAL, 3, path, 360.0, 200.0)
    mc.loadImages()
    mc.initialize()
    out = str(tmp_path / "out.xml")
    mc.writeParameters(out)
    ix = mc.indexing()
    p0 = mc.initialParameters()
    mc.close()
    txt = open(out).read()
    # key order of src/multicalib.cpp:1092-1127
    keys = [k for k in __import__("re").findall(r"^<([a-zA-Z_0-9]+)", txt, flags=__import__("re").M) if k != "opencv_storage"]
    expect = ["nCameras"]
    for c in range(3):
        expect += ["camera_matrix_%d" % c, "camera_distortion_%d" % c, "xi_%d" % c, "camera_pose_%d" % c]
    expect += ["meanReprojectError"] + ["pose_timestamp_%d" % t for t in ix["vertex_timestamp"][3:]]
    assert keys == expect
    fs = cv2.FileStorage(out, cv2.FILE_STORAGE_READ)
    assert int(fs.getNode("nCameras").real()) == 3
    for c in range(3):
        K = fs.getNode("camera_matrix_%d" % c).mat()
        assert K.dtype == np.float32 and K.shape == (3, 3)
        fx, fy, cx, cy, s = rig["cam_K5"][c]
        assert np.array_equal(K, np.array([[fx, s, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32))
        D = fs.getNode("camera_distortion_%d" % c).mat()
        assert np.array_equal(D.ravel(), rig["cam_dist8"][c][:4].astype(np.float32))
        assert np.float32(fs.getNode("xi_%d" % c).real()) == np.float32(rig["cam_xi"][c])
        P = fs.getNode("camera_pose_%d" % c).mat()
        assert P.dtype == np.float32 and P.shape == (4, 4)
        if c > 0:
            assert np.allclose(P[:3, 3], p0[6 * (c - 1) + 3:6 * c], atol=0)
    # the same content written by cv2 itself parses to identical values (text layout may differ in whitespace only)
    ref = str(tmp_path / "ref.xml")
    w = cv2.FileStorage(ref, cv2.FILE_STORAGE_WRITE)
    w.write("camera_pose_1", fs.getNode("camera_pose_1").mat())
    w.release()
    r = cv2.FileStorage(ref, cv2.FILE_STORAGE_READ)
    assert np.array_equal(r.getNode("camera_pose_1").mat(), fs.getNode("camera_pose_1").mat())
    body = lambda t, k: t[t.index("<" + k):t.index("</" + k + ">")]
    assert body(open(ref).read(), "camera_pose_1") == body(txt, "camera_pose_1")       # byte-identical matrix block


@pytest.mark.gpu
def test_run_matches_oracle_from_the_same_start(tmp_path):
    rig, path = _rig_file(tmp_path, n_cam=4, n_frame=60, seed=12, models=[0, 0, 0, 0])
    obsfile.write_rig(path, rig, seed=4)
    mc = multicalib.MultiCameraCalibration(0, 4, path, 360.0, 200.0, criteria=(3, 200, 1e-7))
    mc.loadImages()
    mc.initialize()
    p0 = mc.initialParameters()
    err = mc.optimizeExtrinsics()
    st = mc.stats()
    p = mc.parameters()
    out = str(tmp_path / "out.xml")
    mc.writeParameters(out)
    mc.close()
    O = rigs.to_oracle_rig(rig)
    ref = O.solve(p0, mode=0, crit_type=3, max_count=200, eps=1e-7)
    eo = O.error(ref["params"])
    assert st["iterations"] == ref["iters"]
    scale = np.maximum(np.abs(ref["params"]), 1.0)
    assert np.max(np.abs(p - ref["params"]) / scale) < 1e-6      # the gate; the host class runs the default policy (AUTO: MIXED for this rig)
    assert abs(err - eo["mean_reproj_error"]) <= 1e-9 * err
    assert abs(st["rms"] - eo["rms"]) <= 1e-9 * eo["rms"] and 0.38 < st["rms"] < 0.46
    cv2 = pytest.importorskip("cv2")
    fs = cv2.FileStorage(out, cv2.FILE_STORAGE_READ)
    assert abs(fs.getNode("meanReprojectError").real() - err) == 0.0


@pytest.mark.gpu
def test_outlier_loop(tmp_path):
    rig, path = _rig_file(tmp_path, n_cam=3, n_frame=40, seed=13)
    rig["img"] = rig["img"].copy()
    a, b = rig["edge_off"][5], rig["edge_off"][6]
    rig["img"][a:b] += 6.0                       # one corrupted image
    obsfile.write_rig(path, rig, seed=5)
    mc = multicalib.MultiCameraCalibration(0, 3, path, 360.0, 200.0, criteria=(3, 200, 1e-7))
    e1 = mc.run()
    n = mc.removeOutlier(0.5)                    # per-edge mean error above 0.5 px (src/mymulticalib.cpp:406-423)
    assert n >= 1
    mc.reset(); mc.loadImages(); mc.initialize()
    e2 = mc.optimizeExtrinsics()
    assert e2 < e1
    # dropped images are skipped at reload; timestamps left with a single view are dropped too (mymulticalib.cpp:374-376)
    assert mc.indexing()["edge_cam"].size <= rig["edge_cam"].size - n
    mc.close()

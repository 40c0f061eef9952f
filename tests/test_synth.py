import numpy as np

from multi_camera_calibration_b200 import synth
from tests import rigs


def test_deterministic_and_float32_storage():
    a = synth.make_rig(n_cam=5, n_frame=40, seed=11)
    b = synth.make_rig(n_cam=5, n_frame=40, seed=11)
    for k in ("obj", "img", "params_init", "edge_cam", "edge_pv", "cam_K5"):
        assert np.array_equal(a[k], b[k]), k
    assert a["obj"].dtype == np.float32 and a["img"].dtype == np.float32
    assert np.array_equal(a["params_init"], a["params_init"].astype(np.float32).astype(np.float64))
    assert np.array_equal(a["cam_K5"], a["cam_K5"].astype(np.float32).astype(np.float64))


def test_respects_the_reference_asserts():
    r = synth.make_config(4, n_frame=300)
    t = r["params_true"].reshape(-1, 6)[:, 3:]
    n = np.linalg.norm(t[r["n_cam"] - 1:], axis=1)
    assert n.min() > 300 and n.max() < 3000                       # src/multicalib.cpp:107-113
    assert r["img"].min() >= 0 and r["img"][:, 0].max() < 1920 and r["img"][:, 1].max() < 1080   # :704-715
    assert set(np.unique(r["cam_model"])) == {0, 1}
    assert (r["cam_ndist"][r["cam_model"] == 1] == 4).all()


def test_converges_to_noise_level(oracle_lib):
    r = synth.make_config(2, n_frame=120)
    O = rigs.to_oracle_rig(r)
    out = O.solve(r["params_init"], mode=0, crit_type=3, max_count=200, eps=1e-7)
    rms = np.sqrt(out["cost"] / r["n_points"])
    assert 0.38 < rms < 0.45                                       # 0.3 px per axis -> 0.424 px
    assert out["iters"] < 20


def test_shards_partition_the_rig(oracle_lib):
    r = synth.make_rig(n_cam=4, n_frame=50, seed=5)
    sh = [synth.shard_rig(r, k, 3) for k in range(3)]
    assert sum(s["n_frame"] for s in sh) == r["n_frame"]
    assert sorted(np.concatenate([s["edge_index"] for s in sh]).tolist()) == list(range(r["edge_cam"].size))
    O = rigs.to_oracle_rig(r)
    total = O.eval(r["params_init"], want_blocks=False)
    part = sum(rigs.to_oracle_rig(s).eval(s["params_init"], want_blocks=False) for s in sh)
    assert abs(total - part) <= 1e-12 * total
    for s in sh:
        assert s["edge_pv"].min() >= r["n_cam"] and s["edge_pv"].max() < r["n_cam"] + s["n_frame"]

"""Bit-exact indexing contract: vertices, edges, parameter slots, row offsets."""
import numpy as np

from oracle import indexing
from multi_camera_calibration_b200 import synth


def test_first_seen_order_and_glob_sorting():
    # camera 0 sees timestamps 2 and 10; cv::glob sorts PATH STRINGS, so "10.yaml" is visited before "2.yaml"
    files = [[("d/a/10.yaml", 10), ("d/a/2.yaml", 2), ("d/a/7.yaml", 7)],
             [("d/b/2.yaml", 2), ("d/b/10.yaml", 10), ("d/b/5.yaml", 5)],
             [("d/c/5.yaml", 5), ("d/c/7.yaml", 7)]]
    ix = indexing.build_indexing(3, files)
    assert ix["vertex_timestamp"] == [-1, -1, -1, 10, 2, 7, 5]
    assert ix["edges"] == [(0, 3, 0), (0, 4, 1), (0, 5, 2), (1, 3, 0), (1, 4, 1), (1, 6, 2), (2, 6, 0), (2, 5, 1)]
    assert ix["points_location"] == [108 * i for i in range(9)]
    assert ix["timestamp_cnt"][3:] == [2, 2, 2, 2]


def test_single_view_timestamps_are_dropped():
    files = [[("a/1.yaml", 1), ("a/2.yaml", 2)], [("b/2.yaml", 2), ("b/3.yaml", 3)]]
    ix = indexing.build_indexing(2, files)
    assert ix["vertex_timestamp"] == [-1, -1, 2]
    assert ix["edges"] == [(0, 2, 1), (1, 2, 0)]          # photoIndex counts the skipped file too


def test_param_slots():
    assert indexing.param_slot(0) is None
    assert indexing.param_slot(1) == (0, 6) and indexing.param_slot(7) == (36, 42)


def test_generator_follows_the_contract():
    rig = synth.make_rig(n_cam=4, n_frame=23, seed=3, views_per_frame=2)
    nC = 4
    # rebuild from "files": zero-padded timestamps so lexicographic == numeric
    files = [[] for _ in range(nC)]
    for k in range(23):
        for c in {k % nC, (k + 1) % nC}:
            files[c].append(("cam%d/%06d.yaml" % (c, k), k))
    ix = indexing.build_indexing(nC, files)
    assert [e[0] for e in ix["edges"]] == rig["edge_cam"].tolist()
    assert [e[1] for e in ix["edges"]] == rig["edge_pv"].tolist()
    assert ix["vertex_timestamp"][nC:] == rig["timestamps"].tolist()
    assert (np.array(ix["points_location"]) == 2 * rig["edge_off"]).all()

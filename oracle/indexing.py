"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's vertex / edge / row indexing.

Integers only; this is the "correspondence and indexing must be bit-exact" part of the north star.

Reference (relative to /root/reference):
  vertex / edge structs     include/opencv2/ccalib/multicalib.hpp:86-122
  getPhotoVertex            src/multicalib.cpp:323-346   (linear scan, first-seen order, timestampCnt++)
  multi-camera filter       src/mymulticalib.cpp:314-347 (identifyMultiCameraTimestamps)
  edge creation             src/mymulticalib.cpp:360-403 (cameras outer loop; images in cv::glob order =
                            lexicographically sorted path strings; timestamp = stoi(stem))
  parameter slots           src/multicalib.cpp:422-440, 1058-1075  ([rvec|tvec] at 6*(v-1), vertex 0 = gauge)
  row offsets               src/multicalib.cpp:597-603   (pointsLocation[e+1] = pointsLocation[e] + 2*N_e)
"""
from __future__ import annotations


def build_indexing(n_cam, files_per_camera, n_points_per_file=None):
    """files_per_camera[c] = list of (path_string, timestamp) in ANY order; they are sorted by path string here
    because cv::glob returns sorted paths (so '10.yaml' sorts before '2.yaml').
    n_points_per_file[c][path] = corner count (for the row offsets); default 54.

    Returns dict(vertex_timestamp, edges=[(cameraVertex, photoVertex, photoIndex)], points_location,
                 timestamp_cnt)."""
    sorted_files = [sorted(files_per_camera[c], key=lambda ft: ft[0]) for c in range(n_cam)]
    # identifyMultiCameraTimestamps: a timestamp is kept iff some OTHER camera also has it
    seen_by = {}
    for c in range(n_cam):
        for _, ts in sorted_files[c]:
            seen_by.setdefault(ts, set()).add(c)
    multi = {ts for ts, cams in seen_by.items() if len(cams) >= 2}
    vertex_ts = [-1] * n_cam            # camera vertices carry timestamp -1 (multicalib.hpp:116-120)
    ts_cnt = [1] * n_cam
    edges = []
    for c in range(n_cam):
        for photo_index, (path, ts) in enumerate(sorted_files[c]):
            if ts not in multi:
                continue                  # mymulticalib.cpp:374-376
            pv = -1
            for i, vts in enumerate(vertex_ts):   # getPhotoVertex scans ALL vertices incl. cameras (ts == -1)
                if vts == ts:
                    pv = i
                    ts_cnt[i] += 1
                    break
            if pv < 0:
                vertex_ts.append(ts)
                ts_cnt.append(1)
                pv = len(vertex_ts) - 1
            edges.append((c, pv, photo_index))
    loc = [0]
    for (c, pv, pi) in edges:
        n = 54
        if n_points_per_file is not None:
            n = n_points_per_file[c][sorted_files[c][pi][0]]
        loc.append(loc[-1] + 2 * n)
    return dict(vertex_timestamp=vertex_ts, edges=edges, points_location=loc, timestamp_cnt=ts_cnt)


def param_slot(vertex):
    """[6(v-1), 6(v-1)+6) = [rvec | tvec]; vertex 0 is the gauge and has no slot."""
    if vertex <= 0:
        return None
    return (6 * (vertex - 1), 6 * (vertex - 1) + 6)

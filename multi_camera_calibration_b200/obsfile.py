"""Observation file ("MCCBOBS1") read by the host class's loadImages(): what MyMultiCameraCalibration reads from
<data>/<serial>/<ts>.yaml (corners, objects) and <cfg>/<serial>.xml (Intrinsics, Distortion), plus the per-image PnP
transform, in one little-endian binary file.

  char[8]  "MCCBOBS1"
  int32    n_cam, n_image, 0, 0
  n_cam  x { int32 model, int32 ndist, float64 K5[5] (fx fy cx cy skew), float64 dist8[8], float64 xi }
  n_image x { int32 camera, int32 timestamp, int32 n_points, int32 0, float32 transform[16] }   (load order: cameras
            outer loop, each camera's files in cv::glob order)
  for each image: float32 obj[3 n], float32 img[2 n]
"""
from __future__ import annotations

import struct

import numpy as np


def write(path, cam_model, cam_K5, cam_dist8, cam_ndist, cam_xi, images):
    """images: list of dicts(camera, timestamp, transform (4x4), obj (n,3) f32, img (n,2) f32) in load order."""
    nC = len(cam_model)
    with open(path, "wb") as f:
        f.write(b"MCCBOBS1")
        f.write(struct.pack("<4i", nC, len(images), 0, 0))
        for c in range(nC):
            f.write(struct.pack("<2i", int(cam_model[c]), int(cam_ndist[c])))
            f.write(np.asarray(cam_K5[c], dtype="<f8").tobytes())
            f.write(np.asarray(cam_dist8[c], dtype="<f8").tobytes())
            f.write(struct.pack("<d", float(cam_xi[c])))
        for im in images:
            n = int(np.asarray(im["obj"]).reshape(-1, 3).shape[0])
            f.write(struct.pack("<4i", int(im["camera"]), int(im["timestamp"]), n, 0))
            f.write(np.asarray(im["transform"], dtype="<f4").reshape(16).tobytes())
        for im in images:
            f.write(np.asarray(im["obj"], dtype="<f4").reshape(-1).tobytes())
            f.write(np.asarray(im["img"], dtype="<f4").reshape(-1).tobytes())


def images_from_rig(rig, seed=0, rot_noise=0.01, trans_noise=5.0):
    """One image per edge of a synthetic rig (synth.make_rig), with a PnP-like pattern->camera transform: the true
    composed transform perturbed by N(0, rot_noise rad) / N(0, trans_noise mm), stored float32."""
    from .synth import rodrigues_batch
    rng = np.random.default_rng(seed)
    nC = rig["n_cam"]
    p = rig["params_true"].reshape(-1, 6)
    out = []
    for e in range(rig["edge_cam"].size):
        c, pv = int(rig["edge_cam"][e]), int(rig["edge_pv"][e])
        Rp = rodrigues_batch(p[pv - 1, :3]); tp = p[pv - 1, 3:]
        Rc = rodrigues_batch(p[c - 1, :3]) if c > 0 else np.eye(3)
        tc = p[c - 1, 3:] if c > 0 else np.zeros(3)
        R = rodrigues_batch(rot_noise * rng.standard_normal(3)) @ Rc @ Rp
        t = Rc @ tp + tc + trans_noise * rng.standard_normal(3)
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
        a, b = rig["edge_off"][e], rig["edge_off"][e + 1]
        out.append(dict(camera=c, timestamp=int(rig["timestamps"][pv - nC]), transform=T.astype(np.float32),
                        obj=rig["obj"][a:b], img=rig["img"][a:b]))
    return out


def write_rig(path, rig, seed=0, extra_images=()):
    """Writes a synthetic rig; extra_images (e.g. single-view timestamps) are merged in load order."""
    images = images_from_rig(rig, seed) + list(extra_images)
    images.sort(key=lambda im: (im["camera"], "%012d" % im["timestamp"]))     # cameras outer, glob order inner
    write(path, rig["cam_model"], rig["cam_K5"], rig["cam_dist8"], rig["cam_ndist"], rig["cam_xi"], images)
    return images

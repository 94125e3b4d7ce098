"""Trajectory I/O and gauge alignment around the optimiser.

Python mirror of the reference's helpers in src/OptimizationUtils.cpp (the C++
versions with the reference's signatures live in host/TrajectoryIO.cpp):

  read_camera_intrinsics_from_file   :146-158  last complete 9-number record -> (fx, fy, cx, cy)
  write_keyframe_poses_to_file       :160-172  "timestamp tx ty tz qx qy qz qw", ostream default
                                               formatting (6 significant digits, %g)
  get_first_pose                     :323-375  skips THREE header lines (:332-334), nearest
                                               timestamp (first minimum, src/nearest_interp_1d.cpp),
                                               quaternion normalised by the Sophus constructor
  pose_offset                        :377-384  T <- (initial * T_0^-1) * T  for every keyframe

Poses are 7-vectors (qx, qy, qz, qw, tx, ty, tz): Sophus storage order (headers/sophus/se3.hpp:356-365).
"""
import numpy as np

from . import se3


def read_camera_intrinsics_from_file(path):
    """(fx, fy, cx, cy) of the last complete `fx fy cx cy d0 d1 d2 d3 d4` record; a trailing
    partial record overwrites the leading values it reaches, exactly like the chained
    `infile >> fx >> fy ...` of the reference (a token that does not parse stores 0 and stops)."""
    v = [0.0] * 9
    with open(path) as f:
        toks = f.read().split()
    k = 0
    for t in toks:
        try:
            v[k % 9] = float(t)
        except ValueError:
            v[k % 9] = 0.0
            break
        k += 1
    return np.array(v[:4], dtype=np.float64)


def _g(x):
    return "%g" % x  # == operator<<(double) with the default precision 6


def write_keyframe_poses_to_file(path, timestamps, pose7):
    """One line per keyframe: `timestamp tx ty tz qx qy qz qw` (timestamp is the keyframe's string)."""
    pose7 = np.asarray(pose7, dtype=np.float64).reshape(-1, 7)
    with open(path, "w") as f:
        for ts, p in zip(timestamps, pose7):
            f.write("%s %s %s %s %s %s %s %s\n" % (ts, _g(p[4]), _g(p[5]), _g(p[6]), _g(p[0]), _g(p[1]), _g(p[2]), _g(p[3])))


def read_ground_truth(path):
    """Rows `timestamp tx ty tz qx qy qz qw` after the three header lines the reference skips."""
    with open(path) as f:
        lines = f.read().split("\n")[3:]
    toks = " ".join(lines).split()
    rows = []
    for i in range(0, len(toks) - 7, 8):
        try:
            rows.append([float(t) for t in toks[i:i + 8]])
        except ValueError:
            break
    return np.array(rows, dtype=np.float64).reshape(-1, 8)


def nearest_index(xd, x):
    """Index of the nearest data point, first one on ties (src/nearest_interp_1d.cpp:63-73)."""
    return int(np.argmin(np.abs(x - np.asarray(xd, dtype=np.float64))))


def get_first_pose(first_timestamp, ground_truth_path):
    gt = read_ground_truth(ground_truth_path)
    r = gt[nearest_index(gt[:, 0], float(first_timestamp))]
    q = r[4:8] / np.sqrt(np.dot(r[4:8], r[4:8]))  # SO3(quaternion) normalises (so3.hpp:203-205)
    return np.concatenate([q, r[1:4]])


def pose_offset(pose7, initial_pose):
    """In place: every pose pre-multiplied by initial_pose * pose7[0]^-1 (:377-384)."""
    pose7 = np.asarray(pose7, dtype=np.float64).reshape(-1, 7)
    delta = se3.mul(np.asarray(initial_pose, dtype=np.float64), se3.inverse(pose7[0]))
    for i in range(pose7.shape[0]):
        pose7[i] = se3.mul(delta, pose7[i])
    return pose7

// TrajectoryIO.cpp -- the I/O and gauge helpers that surround windowOptimize in the
// reference's src/OptimizationUtils.cpp, with the reference's signatures
// (headers/OptimizationUtils.h): SURVEY.md 8f row N2.  Plain host C++ (no CUDA).
//
//   read_camera_intrinsics_from_file   :146-158
//   write_keyframe_poses_to_file       :160-172
//   getFirstPose                       :323-375 (+ src/nearest_interp_1d.cpp:63-73)
//   poseOffset                         :377-384
//
// Inside the reference tree compile with -DBA_USE_REFERENCE_HEADERS (real Sophus / Eigen).
#ifdef BA_USE_REFERENCE_HEADERS
#include "OptimizationUtils.h"
#else
#include "compat/reference_types.h"
#endif

#include "se3_raw.h"

#include <cmath>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

using std::string;
using std::vector;

// The last complete "fx fy cx cy d0 d1 d2 d3 d4" record wins; a trailing partial record
// overwrites the fields it reaches (chained stream extraction, :150-152).
Vector4d read_camera_intrinsics_from_file(const string &file_path) {
  std::ifstream in(file_path);
  double rec[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; in >> rec[k % 9]; ++k) {
  }
  return Vector4d(rec[0], rec[1], rec[2], rec[3]);
}

// "timestamp tx ty tz qx qy qz qw" per keyframe, default ostream formatting (:160-172)
void write_keyframe_poses_to_file(const string &file_path, const vector<KeyFrame> &keyframes) {
  std::ofstream out(file_path);
  for (const KeyFrame &kf : keyframes) {
    const double *p = kf.T_w_c.data();  // (qx, qy, qz, qw, tx, ty, tz): se3.hpp:356-365
    out << kf.timestamp << " " << p[4] << " " << p[5] << " " << p[6] << " " << p[0] << " " << p[1] << " " << p[2] << " " << p[3]
        << "\n";
  }
}

// nearest ground-truth pose to first_timestamp; the first THREE lines of the file are skipped (:332-334)
Sophus::SE3d getFirstPose(const string &first_timestamp, const string &ground_truth_file_path) {
  std::ifstream in(ground_truth_file_path);
  string header;
  for (int i = 0; i < 3; ++i) std::getline(in, header);
  vector<double> rows;
  double v[8];
  for (;;) {
    int k = 0;
    while (k < 8 && (in >> v[k])) ++k;
    if (k < 8) break;
    rows.insert(rows.end(), v, v + 8);
  }
  const double want = std::stod(first_timestamp);
  size_t best = 0;
  double d = rows.empty() ? 0.0 : std::fabs(want - rows[0]);
  for (size_t i = 1; i * 8 < rows.size(); ++i) {  // first minimum wins (nearest_interp_1d.cpp:66-71)
    const double d2 = std::fabs(want - rows[8 * i]);
    if (d2 < d) {
      best = i;
      d = d2;
    }
  }
  if (rows.empty()) return Sophus::SE3d();
  const double *r = &rows[8 * best];
  // SE3(quaternion, translation) normalises the quaternion (so3.hpp:203-205)
  return Sophus::SE3d(Eigen::Quaterniond(r[7], r[4], r[5], r[6]), Sophus::SE3d::Point(r[1], r[2], r[3]));
}

// every pose pre-multiplied by initial_pose * T_0^-1 (:377-384)
void poseOffset(vector<KeyFrame> &keyframes, const Sophus::SE3d &initial_pose) {
  if (keyframes.empty()) return;
  const Sophus::SE3d delta = initial_pose * keyframes[0].T_w_c.inverse();
  for (KeyFrame &kf : keyframes) kf.T_w_c = delta * kf.T_w_c;
}

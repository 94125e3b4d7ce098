// reference_types.h -- stand-ins for the third-party and first-party types that
// appear in the signatures of the reference's optimiser entry points
// (headers/OptimizationUtils.h:42,55).  ONLY used when this repository is built
// on its own (no Eigen / Sophus / OpenCV / Ceres in the image);
// inside the reference's tree the wrapper is compiled against the real headers
// (define BA_USE_REFERENCE_HEADERS).
//
// What is mirrored, and from where:
//   KeyFrame / Landmark / Map3D        headers/CommonTypes.h:13-43 (member names)
//   ceresGlobalProblem                 headers/BundleAdjustmentConfig.h:44-69
//   Sophus::SE3d                       headers/sophus/se3.hpp, so3.hpp: storage order
//     (qx,qy,qz,qw,tx,ty,tz) :356-365, inverse :186-189, product :317-321 with
//     the 2/(1+|q|^2) renormalisation so3.hpp:339-356, point action :299-301
//   ceres::Solver::Options             the five fields the reference sets (:62-66)
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace Eigen {
template <int N>
struct VecN {
  double v[N];
  VecN() { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  double &operator()(int i) { return v[i]; }
  const double &operator()(int i) const { return v[i]; }
  double &operator[](int i) { return v[i]; }
  const double &operator[](int i) const { return v[i]; }
  double *data() { return v; }
  const double *data() const { return v; }
};
struct Vector2d : VecN<2> {
  Vector2d() {}
  Vector2d(double a, double b) { v[0] = a; v[1] = b; }
};
struct Vector3d : VecN<3> {
  Vector3d() {}
  Vector3d(double a, double b, double c) { v[0] = a; v[1] = b; v[2] = c; }
};
struct Vector4d : VecN<4> {
  Vector4d() {}
  Vector4d(double a, double b, double c, double d) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
};
// Eigen::Quaterniond(w, x, y, z): the one constructor the wrapper uses (argument order of Eigen 3.4.0)
struct Quaterniond {
  double w_, x_, y_, z_;
  Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
  double w() const { return w_; }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
};
}  // namespace Eigen
using Eigen::Vector2d;
using Eigen::Vector3d;
using Eigen::Vector4d;

namespace Sophus {
class SE3d {
 public:
  static const int num_parameters = 7;
  using Point = Eigen::Vector3d;
  // the constructors below are the ones headers/sophus/se3.hpp:407-455 has (default, copy, quaternion + translation);
  // raw storage is reached through data() only (:469-476), as in the real class
  SE3d(const Eigen::Quaterniond &q, const Point &t) {  // SO3(quaternion) normalises (so3.hpp:203-205)
    const double n = std::sqrt(q.x() * q.x() + q.y() * q.y() + q.z() * q.z() + q.w() * q.w());
    p_[0] = q.x() / n; p_[1] = q.y() / n; p_[2] = q.z() / n; p_[3] = q.w() / n;
    p_[4] = t[0]; p_[5] = t[1]; p_[6] = t[2];
  }
  SE3d() { p_[0] = p_[1] = p_[2] = 0.0; p_[3] = 1.0; p_[4] = p_[5] = p_[6] = 0.0; }
  double *data() { return p_; }
  const double *data() const { return p_; }
  SE3d inverse() const {
    double q[4] = {-p_[0], -p_[1], -p_[2], p_[3]};
    const double len = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (double &c : q) c /= len;
    const double nt[3] = {p_[4] * -1.0, p_[5] * -1.0, p_[6] * -1.0};
    SE3d r;
    rotate(q, nt, r.p_ + 4);
    std::memcpy(r.p_, q, sizeof(q));
    return r;
  }
  SE3d operator*(const SE3d &b) const {
    SE3d r(*this);
    double rt[3];
    rotate(p_, b.p_ + 4, rt);
    for (int i = 0; i < 3; ++i) r.p_[4 + i] += rt[i];
    const double ax = p_[0], ay = p_[1], az = p_[2], aw = p_[3];
    const double bx = b.p_[0], by = b.p_[1], bz = b.p_[2], bw = b.p_[3];
    double o[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                   aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
    const double n2 = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3];
    if (n2 != 1.0) {
      const double s = 2.0 / (1.0 + n2);
      for (double &c : o) c *= s;
    }
    std::memcpy(r.p_, o, sizeof(o));
    return r;
  }
  Vector3d operator*(const Vector3d &x) const {
    Vector3d r;
    rotate(p_, x.data(), r.data());
    for (int i = 0; i < 3; ++i) r[i] += p_[4 + i];
    return r;
  }

 private:
  // Eigen _transformVector: v + w*uv + q.vec x uv with uv = 2 (q.vec x v)
  static void rotate(const double *q, const double *v, double *o) {
    double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    for (double &c : uv) c += c;
    const double c3[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
    for (int i = 0; i < 3; ++i) o[i] = v[i] + q[3] * uv[i] + c3[i];
  }
  double p_[7];
};
}  // namespace Sophus

namespace cv {
struct Point2f {
  float x = 0.f, y = 0.f;
};
struct Point2d {
  double x = 0.0, y = 0.0;
};
struct KeyPoint {
  Point2f pt;
};
struct Mat {};  // descriptors: never read by the optimiser
}  // namespace cv

namespace ceres {
enum LinearSolverType { DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
struct Solver {
  struct Options {
    LinearSolverType linear_solver_type = SPARSE_SCHUR;
    bool minimizer_progress_to_stdout = false;
    int max_num_iterations = 50;
    double eta = 1e-1;
    double function_tolerance = 1e-6;   // (Ceres 2.0.0 defaults, solver.h)
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
    int num_threads = 1;
  };
};
}  // namespace ceres

typedef int LandmarkId;
typedef unsigned int uint;

struct KeyFrame {
  uint frame_id = 0;
  std::string timestamp;
  Sophus::SE3d T_w_c;                             // camera -> world
  std::vector<cv::KeyPoint> keypoints;            // pixel of every local feature
  cv::Mat descriptors;
  std::vector<Vector3d> points3d_local;           // back-projected features, camera frame (z = depth)
  std::unordered_map<int, LandmarkId> global_points_map;  // local feature id -> landmark id
};
typedef std::pair<int, cv::Point2d> Observation;
typedef std::vector<Observation> Observations;
struct Landmark {
  Observations observations;  // not read by the optimiser
  Vector3d point;             // world
};
typedef std::unordered_map<LandmarkId, Landmark> Map3D;

class ceresGlobalProblem {
 public:
  const double HUB_P_REPR = 1e-3;
  const double WEIGHT_INTRINSICS = 1e-6;
  const double WEIGHT_UNPR = 10;
  const double HUB_P_UNPR = 1e-3;
  const int frame_frequency = 10;
  const int window_size = 0;
  ceres::Solver::Options options;
  ceresGlobalProblem() {
    options.linear_solver_type = ceres::SPARSE_SCHUR;
    options.minimizer_progress_to_stdout = true;
    options.max_num_iterations = 75;
    options.eta = 1e-6;
    options.trust_region_strategy_type = ceres::LEVENBERG_MARQUARDT;
  }
};

// I/O and gauge helpers around the hot path (headers/OptimizationUtils.h), host/TrajectoryIO.cpp
Vector4d read_camera_intrinsics_from_file(const std::string &file_path);
void write_keyframe_poses_to_file(const std::string &file_path, const std::vector<KeyFrame> &keyframes);
Sophus::SE3d getFirstPose(const std::string &first_timestamp, const std::string &ground_truth_file_path);
void poseOffset(std::vector<KeyFrame> &keyframes, const Sophus::SE3d &initial_pose);

// the two entry points of the hot path (headers/OptimizationUtils.h:42, 55)
int countConstraints(const Map3D &map, const std::vector<KeyFrame> &keyframes, int kf_i, int kf_f);
bool windowOptimize(ceresGlobalProblem &globalProblem, int kf_i, int kf_f, std::vector<KeyFrame> &keyframes, Map3D &map,
                    const Vector4d &intrinsics_initial, Vector4d &intrinsics_optimized);

// The reference's optimisation schedule (src/main.cpp:161-182) over a recorded synthetic sequence, through
//     bool windowOptimize(ceresGlobalProblem&, int, int, vector<KeyFrame>&, Map3D&, const Vector4d&, Vector4d&)
// (headers/OptimizationUtils.h:55) -- WHICHEVER implementation is linked:
//   * the reference's own src/OptimizationUtils.cpp (its cost functors, ceres::Problem / ceres::Solve): the true CPU
//     reference arm SURVEY.md 8(d) and BASELINE.md 3 promise for hosts that have Ceres 2.0.0 + Eigen + OpenCV
//     (bench/CMakeLists.txt, built only when find_package(Ceres) succeeds -- none of them is in this repository's image);
//   * this repository's host/OptimizationUtils_gpu.cpp + libba_gpu.so (make -C bench): the same program, the same
//     containers, the same schedule on the B200.
// Both runs write the optimised poses / landmarks / intrinsics to a file: scripts/dump_sequence.py --compare a b checks
// them against each other (north-star tolerance: poses 1e-6 m / 1e-6 rad) -- the end-to-end parity check against real
// Ceres that cannot be run inside this repository's container.
//
//   ceres_baseline <sequence.bin> <result.bin> [--window 20] [--every 10] [--iterations 10] [--passes 2] [--global] [--tolerances]
//
// sequence.bin is written by scripts/dump_sequence.py (little endian): int32 magic 0xBA5E0001, n_kf, n_obs, n_lm;
// int32 kf_ptr[n_kf+1]; int32 lm[n_obs] (landmark id per local feature, insertion order); float uv[2 n_obs];
// double depth[n_obs]; double pose7[7 n_kf] (Sophus storage: qx qy qz qw tx ty tz); int32 lm_id[n_lm]; double lm_pt[3 n_lm];
// double intrinsics[4].
#ifdef BA_USE_REFERENCE_HEADERS
#include "OptimizationUtils.h"
#else
#include "../3dsmc-bundle-adjustment_b200/host/compat/reference_types.h"
#endif

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {
template <class T>
bool rd(FILE *f, std::vector<T> &v, size_t n) {
  v.resize(n);
  return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}
}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s sequence.bin result.bin [--window N] [--every N] [--iterations N] [--passes N] [--global] [--tolerances]\n", argv[0]);
    return 2;
  }
  int window = 20, every = 10, iterations = 10, passes = 2;  // (pass 1 warms up: device context, allocations, caches)
  bool global = false, tolerances = false;
  for (int a = 3; a < argc; ++a) {
    const std::string s = argv[a];
    if (s == "--window" && a + 1 < argc) window = std::atoi(argv[++a]);
    else if (s == "--every" && a + 1 < argc) every = std::atoi(argv[++a]);
    else if (s == "--iterations" && a + 1 < argc) iterations = std::atoi(argv[++a]);
    else if (s == "--passes" && a + 1 < argc) passes = std::max(1, std::atoi(argv[++a]));
    else if (s == "--global") global = true;
    else if (s == "--tolerances") tolerances = true;
  }
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) {
    std::perror(argv[1]);
    return 1;
  }
  int32_t hdr[4];
  if (fread(hdr, 4, 4, f) != 4 || (uint32_t)hdr[0] != 0xBA5E0001u) {
    std::fprintf(stderr, "%s: not a sequence dump\n", argv[1]);
    return 1;
  }
  const int n_kf = hdr[1], n_obs = hdr[2], n_lm = hdr[3];
  std::vector<int32_t> kf_ptr, lm, lm_id;
  std::vector<float> uv;
  std::vector<double> depth, pose7, lm_pt, intr;
  if (!rd(f, kf_ptr, (size_t)n_kf + 1) || !rd(f, lm, (size_t)n_obs) || !rd(f, uv, (size_t)2 * n_obs) || !rd(f, depth, (size_t)n_obs) ||
      !rd(f, pose7, (size_t)7 * n_kf) || !rd(f, lm_id, (size_t)n_lm) || !rd(f, lm_pt, (size_t)3 * n_lm) || !rd(f, intr, 4)) {
    std::fprintf(stderr, "%s: truncated\n", argv[1]);
    return 1;
  }
  std::fclose(f);
  if (window > n_kf) window = n_kf;

  double seconds = 0.0;
  int calls = 0;
  bool ok = true;
  std::vector<KeyFrame> keyframes;
  Map3D map;
  Eigen::Vector4d intr0(intr[0], intr[1], intr[2], intr[3]), intr1 = intr0;
  for (int pass = 0; pass < passes; ++pass) {  // (every pass starts from the recorded state; the last one is reported)
  seconds = 0.0;
  calls = 0;
  keyframes.clear();
  map.clear();
  intr1 = intr0;
  // containers as the tracking front end leaves them (src/main.cpp:25-82, src/Map3D.cpp:29-74): key points, local 3-D points
  // (only z is read by the optimiser), local feature id -> landmark id in insertion order, landmarks by id
  std::unordered_map<int, int> lm_row;
  for (int l = 0; l < n_lm; ++l) lm_row[lm_id[l]] = l;
  keyframes.reserve(n_kf);
  auto arrive = [&](int k) {
    keyframes.emplace_back();
    KeyFrame &kf = keyframes[k];
    kf.frame_id = (uint)k;
    std::memcpy(kf.T_w_c.data(), pose7.data() + (size_t)7 * k, 7 * sizeof(double));
    const int a = kf_ptr[k], b = kf_ptr[k + 1];
    kf.keypoints.resize(b - a);
    kf.points3d_local.resize(b - a);
    for (int i = a; i < b; ++i) {
      kf.keypoints[i - a].pt.x = uv[2 * (size_t)i];
      kf.keypoints[i - a].pt.y = uv[2 * (size_t)i + 1];
      kf.points3d_local[i - a] = Eigen::Vector3d(0.0, 0.0, depth[i]);
      kf.global_points_map.insert({i - a, lm[i]});
      if (map.find(lm[i]) == map.end()) {
        const int l = lm_row.at(lm[i]);
        Landmark L;
        L.point = Eigen::Vector3d(lm_pt[3 * (size_t)l], lm_pt[3 * (size_t)l + 1], lm_pt[3 * (size_t)l + 2]);
        map.insert({lm[i], L});
      }
    }
  };

  ceresGlobalProblem gp;
  gp.options.max_num_iterations = iterations;
  if (!tolerances)  // fixed iteration count: the throughput metric of BASELINE.json (LM iterations per second)
    gp.options.function_tolerance = gp.options.gradient_tolerance = gp.options.parameter_tolerance = 0.0;
  auto run = [&](int kf_i, int kf_f) {
    const auto t0 = std::chrono::steady_clock::now();
    ok = windowOptimize(gp, kf_i, kf_f, keyframes, map, intr0, intr1) && ok;
    seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ++calls;
  };
  if (global) {
    for (int k = 0; k < n_kf; ++k) arrive(k);
    run(0, n_kf - 1);  // src/main.cpp:178-182
  } else {
    for (int size = 1; size <= n_kf; ++size) {
      arrive(size - 1);
      if (size % every == 0 && size >= window) run(size - window, size - 1);  // :161-166
    }
    if (n_kf % every != 0) run(n_kf - window, n_kf - 1);  // leftovers, :169-175
  }
  }  // passes

  // result: poses, landmarks (rows of the input table; NaN for landmarks nobody referenced), intrinsics
  FILE *o = std::fopen(argv[2], "wb");
  if (!o) {
    std::perror(argv[2]);
    return 1;
  }
  const int32_t ohdr[4] = {(int32_t)0xBA5E0002u, n_kf, n_lm, calls};
  fwrite(ohdr, 4, 4, o);
  for (int k = 0; k < n_kf; ++k) fwrite(keyframes[k].T_w_c.data(), sizeof(double), 7, o);
  for (int l = 0; l < n_lm; ++l) {
    double p[3] = {NAN, NAN, NAN};
    auto it = map.find(lm_id[l]);
    if (it != map.end())
      for (int j = 0; j < 3; ++j) p[j] = it->second.point(j);
    fwrite(p, sizeof(double), 3, o);
  }
  const double k4[4] = {intr1(0), intr1(1), intr1(2), intr1(3)};
  fwrite(k4, sizeof(double), 4, o);
  std::fclose(o);
  std::printf("{\"windows\": %d, \"ok\": %s, \"seconds\": %.6f, \"windows_per_s\": %.3f, \"lm_iterations_per_s_if_fixed\": %.3f, "
              "\"keyframes\": %d, \"window\": %d, \"iterations\": %d, \"global\": %s}\n",
              calls, ok ? "true" : "false", seconds, calls / seconds, tolerances ? 0.0 : calls * (double)iterations / seconds, n_kf,
              global ? n_kf : window, iterations, global ? "true" : "false");
  return ok ? 0 : 1;
}

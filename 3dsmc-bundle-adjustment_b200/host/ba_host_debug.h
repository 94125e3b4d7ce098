// Introspection of the last problem windowOptimize handed to the GPU (tests only).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/ba_gpu.h"

struct BaHostLastProblem {
  std::vector<int32_t> cam_idx, pt_idx;
  std::vector<int> landmark_of_pt;
  int admissible_obs = 0;
  ba_gpu_summary summary{};
  // host wall clock of the last call, per phase: container walk + frame change, ba_gpu_upload, ba_gpu_solve,
  // ba_gpu_download, write-back into the caller's containers
  double ms_extract = 0, ms_upload = 0, ms_solve = 0, ms_download = 0, ms_writeback = 0;
};
const BaHostLastProblem &ba_host_last_problem();
// measurement only: tolerances off, so that every call runs exactly options.max_num_iterations LM iterations
void ba_host_fixed_iterations(bool on);

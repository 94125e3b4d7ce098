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
};
const BaHostLastProblem &ba_host_last_problem();

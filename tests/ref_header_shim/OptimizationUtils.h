// TEST FIXTURE: stands in for the reference's headers/OptimizationUtils.h when host/*.cpp are compiled with
// -DBA_USE_REFERENCE_HEADERS (the build INTEGRATION.md section 1 prescribes inside the reference tree).
// Eigen / Ceres / OpenCV are not in the image, so the real header cannot be included; this shim exposes the
// SAME names with ONLY the members the real types have -- in particular Sophus::SE3d has the constructors of
// headers/sophus/se3.hpp:407-455 (default, copy, quaternion + translation) and data() (:469-476) and NO
// constructor from a raw pointer -- and it reproduces the reference header's `using namespace` directives
// (headers/OptimizationUtils.h:13, headers/CommonTypes.h) so that name clashes show up here.
#pragma once
#include "../../3dsmc-bundle-adjustment_b200/host/compat/reference_types.h"
using namespace std;
using namespace Eigen;
int findLocalPointIndex(const KeyFrame &keyframe, const int landmarkId);
bool optimizeDebug(ceresGlobalProblem &globalProblem, const Vector4d &intrinsics_initial, Vector4d &intrinsics_optimized,
                   int opt_type);

// se3_raw.h -- pose <-> 7 contiguous doubles (qx,qy,qz,qw,tx,ty,tz), through the ONLY raw-storage access the
// reference's vendored Sophus offers: SE3::data() (headers/sophus/se3.hpp:356-365 documents the storage order; the
// reference itself hands pose.data() to Ceres at src/OptimizationUtils.cpp:251).  There is no SE3(const double*)
// constructor in headers/sophus/se3.hpp:407-476, so none is used here or in compat/reference_types.h.
#pragma once
#include <cstring>

inline Sophus::SE3d se3_from_raw(const double *p7) {
  Sophus::SE3d T;
  std::memcpy(T.data(), p7, 7 * sizeof(double));
  return T;
}

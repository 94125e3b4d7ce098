/*
 * ba_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the windowed bundle-adjustment hot path of
 * martinxluptak/3dsmc-bundle-adjustment (src/OptimizationUtils.cpp:21-137,
 * 215-313) together with the third-party behaviour it invokes:
 *   - ceres-solver 2.0.0 (conanfile.txt:4): AutoDiffCostFunction (Jets),
 *     HuberLoss + Corrector, LocalParameterization chain rule, Jacobi column
 *     scaling, Levenberg-Marquardt trust region, Schur elimination,
 *     ITERATIVE_SCHUR conjugate gradients.
 *   - Sophus SE3/SO3 (headers/sophus/se3.hpp, so3.hpp) and the Eigen 3.4.0
 *     quaternion formulas they call.
 * Ceres and Eigen are NOT in /root/reference and NOT installed in the build
 * image, so their published algorithms are restated from the call sites.
 *
 * PARITY STATUS: "parity unpinned" at the Ceres boundary -- the reference has
 * no test, fixture or golden vector for windowOptimize, residuals or Jacobians
 * (SURVEY.md section 4 / 8c).  What pins this oracle instead: 50-digit mpmath
 * golden vectors for residuals / Jacobians / SE3 ops (tests/golden/, generated
 * by tests/golden/make_golden.py) and noise-free synthetic problems with known
 * optimum.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call this file.  The product
 * (3dsmc-bundle-adjustment_b200/) never does.
 */
#ifndef BA_ORACLE_H
#define BA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Problem in the flat layout of the C-ABI upload (include/ba_gpu.h).
 * pose7: (qx,qy,qz,qw,tx,ty,tz) camera->world, Sophus storage order
 * (headers/sophus/se3.hpp:356-365).  Observations are in the canonical
 * reference order (camera-major). */
typedef struct ora_problem {
  int32_t n_cam, n_pt, n_obs, fixed_cam; /* fixed_cam < 0: none fixed */
  double *pose7;                         /* [n_cam*7]  in/out */
  double *pt3;                           /* [n_pt*3]   in/out */
  const int32_t *cam_idx;                /* [n_obs] */
  const int32_t *pt_idx;                 /* [n_obs] */
  const double *uv2;                     /* [n_obs*2] */
  const double *depth;                   /* [n_obs] or NULL */
  double intr[4];                        /* fx fy cx cy, in/out */
  double intr_prior[4];
} ora_problem;

/* DENSE_SCHUR: dense S + dense Cholesky; IMPLICIT_PCG: ITERATIVE_SCHUR-equivalent; SPARSE_SCHUR: the reference's own
 * setting (headers/BundleAdjustmentConfig.h:62) -- explicit S in envelope storage + sparse Cholesky (exact step) */
enum { ORA_SOLVER_DENSE_SCHUR = 0, ORA_SOLVER_IMPLICIT_PCG = 1, ORA_SOLVER_SPARSE_SCHUR = 2 };

typedef struct ora_options {
  /* reference knobs, headers/BundleAdjustmentConfig.h:47-50,64-65 */
  double huber_repr, huber_unpr, weight_unpr, weight_intrinsics;
  int32_t max_num_iterations;
  double eta;
  /* cost-model switches (REF mode = both 1; NS mode = both 0) */
  int32_t use_depth_prior, optimize_intrinsics;
  int32_t solver;
  int64_t n_obs_total; /* weight normaliser; 0 -> problem.n_obs */
  /* Ceres 2.0.0 defaults (solver.h), SURVEY Appendix A */
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_radius, max_radius, min_radius, min_relative_decrease;
  double min_lm_diagonal, max_lm_diagonal;
  int32_t max_consecutive_invalid_steps, jacobi_scaling;
  int32_t max_pcg_iterations, min_pcg_iterations, residual_reset_period;
  int32_t num_threads; /* OpenMP threads; Ceres default is 1 */
} ora_options;

enum {
  ORA_TERM_NO_CONVERGENCE = 0,
  ORA_TERM_GRADIENT = 1,
  ORA_TERM_PARAMETER = 2,
  ORA_TERM_FUNCTION = 3,
  ORA_TERM_MIN_RADIUS = 4,
  ORA_TERM_FAILURE = 5
};

typedef struct ora_iter {
  int32_t iteration, step_is_valid, step_is_successful, linear_iters;
  double cost, cost_change, gradient_max_norm, step_norm, relative_decrease,
      radius, model_cost_change;
} ora_iter;

typedef struct ora_summary {
  int32_t termination, num_iterations, num_successful, num_unsuccessful;
  double initial_cost, final_cost;
  int64_t total_linear_iters;
  double seconds_total, seconds_linearize, seconds_linear_solve;
} ora_summary;

void ora_default_options(ora_options *o);

/* ---- SE3 (Sophus restatement) ---- */
void ora_se3_exp(const double delta6[6], double out7[7]);
void ora_se3_mul(const double a7[7], const double b7[7], double out7[7]);
void ora_se3_inverse(const double a7[7], double out7[7]);
void ora_se3_act(const double a7[7], const double p[3], double out[3]);
/* src/Map3D.cpp:76-97 (+ :44 when pose7 != NULL); -1 if a key point is outside the image */
int ora_backproject(int n, const float *uv2f, const float *depth_img, int width, int height, const double intr4[4],
                    const double *pose7, double *local3, double *world3);
void ora_se3_dx_this_mul_exp_x_at_0(const double a7[7], double J7x6[42]);

/* ---- cost functors with ambient (autodiff-equivalent) Jacobians ----
 * Any Jacobian pointer may be NULL. Row-major. */
void ora_reprojection(const double pose7[7], const double pt[3],
                      const double intr[4], const double uv[2], double weight,
                      double r[2], double Jpose2x7[14], double Jpt2x3[6],
                      double Jintr2x4[8]);
void ora_depth_prior(const double pose7[7], const double pt[3],
                     const double intr[4], double depth, double weight,
                     double r[1], double Jpose1x7[7], double Jpt1x3[3],
                     double Jintr1x4[4]);
void ora_intrinsics_prior(const double intr[4], const double prior[4],
                          double weight, double r[4], double J4x4[16]);

/* Huber loss as ceres::HuberLoss::Evaluate. rho[3]. */
void ora_huber(double a, double s, double rho[3]);

/* Full evaluation at the current state: robustified residuals and LOCAL
 * Jacobians (pose block already multiplied by the 7x6 plus-Jacobian), in the
 * canonical observation order. R = 2 + use_depth_prior rows per observation.
 *   r   [n_obs*R]       Jc [n_obs*R*6]   Jp [n_obs*R*3]
 *   Jk  [n_obs*2*4]     (reprojection rows only; zero for depth rows)
 *   g_c [n_cam*6] (zero for the fixed camera), g_p [n_pt*3], g_k[4]
 * Any output may be NULL. Returns 0, or -1 on a non-finite evaluation. */
int ora_evaluate(const ora_problem *p, const ora_options *o, double *cost,
                 double *r, double *Jc, double *Jp, double *Jk, double *g_c,
                 double *g_p, double *g_k);

/* Stable counting sort of pt_idx: point-major permutation + CSR row pointers.
 * perm [n_obs], pt_rowptr [n_pt+1], cam_rowptr [n_cam+1]. */
void ora_build_indices(const ora_problem *p, int32_t *perm, int32_t *pt_rowptr,
                       int32_t *cam_rowptr);

/* Levenberg-Marquardt solve restating ceres::Solve.
 * Mutates p->pose7, p->pt3, p->intr in place. trace may be NULL. */
int ora_solve(ora_problem *p, const ora_options *o, ora_summary *s,
              ora_iter *trace, int32_t trace_cap);

/* One reduced-system matvec y = S x at the current state with the given
 * radius (test hook for the implicit-Schur kernels and the 2-rank gloo test).
 * x, y: [n_cam*6] in UNSCALED coordinates, fixed camera rows = damping only.
 * If pt_begin<pt_end only points in [pt_begin,pt_end) contribute (a shard's
 * partial product, without the U/damping term when include_diag==0). */
int ora_schur_matvec(const ora_problem *p, const ora_options *o, double radius,
                     const double *x, double *y, int32_t pt_begin,
                     int32_t pt_end, int32_t include_diag);

#ifdef __cplusplus
}
#endif
#endif

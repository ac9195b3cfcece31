// K2/K3/K4: persistent, cluster-per-frame ECC registration.
//
// One thread-block cluster owns one frame and runs its complete coarse-to-fine alignment on the device:
// every pyramid level, every Gauss-Newton / Levenberg-Marquardt iteration, the optional translation-first
// pass and the correlation gate.  Per trial there is exactly ONE fused pass over the level (warp + residual
// [+ warped-image gradient] + steepest-descent images + all normal-equation sums), one cluster barrier, a
// DSMEM reduction in fixed order (deterministic) and a scalar solver step that every CTA of the cluster
// executes redundantly in double precision, so no broadcast is needed.
//
// Reference semantics reproduced here (core/proc/image_registration/):
//   c_ecch::align                               ecc2.cc:1133-1176
//   c_ecc_forward_additive::align               ecc2.cc:1247-1365
//   c_ecclm::align / compute_jac / compute_rhs  ecc2.cc:1444-1650
//   c_ecc_inverse_compositional::align          ecc2.cc:1693-1788
//   c_ecclm_inverse_compositional::align        ecc2.cc:1894-2086
//   ecc_differentiate                           ecc2.cc:142-169
//   compute_correlation                         ecc2.cc:65-137
//   c_image_transform::{scale_transfrom, eps, invert_and_compose, create_steepest_descent_images}
//                                               c_image_transform.cc / c_image_transform.h (per type)
//   c_frame_registration::register_frame flow   c_frame_registration.cc:797-872
#pragma once
#include "ssk_ecc.cuh"
#include <cooperative_groups.h>
#include <float.h>

namespace cg = cooperative_groups;

namespace ssk {

namespace {

#ifndef SSK_ECC_NT
#define SSK_ECC_NT 256
#endif
#ifndef SSK_ECC_MINB
#define SSK_ECC_MINB 2
#endif
constexpr int NT = SSK_ECC_NT;
constexpr int NW = NT / 32;
constexpr int NSMAX = 72;
constexpr int FT_W = 32, FT_H = NT / 32;     // stencil tile of the forward methods (one pixel per thread)
constexpr int FT_IW = FT_W + 4, FT_IH = FT_H + 4;

template <int TYPE> struct NParams;
template <> struct NParams<SSK_MOTION_TRANSLATION> { static constexpr int M = 2; };
template <> struct NParams<SSK_MOTION_EUCLIDEAN> { static constexpr int M = 3; };
template <> struct NParams<SSK_MOTION_SCALED_EUCLIDEAN> { static constexpr int M = 4; };
template <> struct NParams<SSK_MOTION_AFFINE> { static constexpr int M = 6; };
template <> struct NParams<SSK_MOTION_HOMOGRAPHY> { static constexpr int M = 8; };

__host__ __device__ inline int nparams_of(int type) {
  switch (type) {
    case SSK_MOTION_TRANSLATION: return 2;
    case SSK_MOTION_EUCLIDEAN: return 3;
    case SSK_MOTION_SCALED_EUCLIDEAN: return 4;
    case SSK_MOTION_AFFINE: return 6;
    default: return 8;
  }
}

// coefficients the steepest-descent images are evaluated with
struct JCoef { float c[8]; };

struct Shared {
  double part[2][NSMAX];      // this CTA's partial sums (double buffered; peers read them through DSMEM)
  double tot[NSMAX];          // cluster totals
  double wpart[NW][NSMAX];
  float gw[FT_IH][FT_IW + 1]; // warped-image tile of the forward methods
  unsigned long long tl_full[8], tl_empty[8];   // TMA-staged pass: transaction barrier / consumer barrier per stage
  int4 tl_meta[8][2];         // per stage: {window origin x, y, first tap-safe wx, count} {first tap-safe wy, count, -, -}
  // ---- solver state: identical in every CTA of the cluster ----
  ssk_transform t;            // transform being estimated (accepted parameters)
  ssk_transform tq;           // parameters of the next pass
  ssk_transform tmain;        // main transform parked during the translation-first pass
  MapCoef map;                // map coefficients of tq
  JCoef jc;                   // steepest-descent coefficients
  float Hp[64];               // reference-side / current normal matrix (float, like cv::Mat1f)
  float v[8];                 // projected error (float)
  float vtrial[8];
  float Htrial[64];
  float deltap[8];
  float fa_a, fa_c;           // forward-additive residual coefficients: rhs = fma(f, fa_a, g) - fa_c
  double err, newerr, lambda, dp, rmsold, eps;
  int num_it, recompute, brk, converged, failed, level_ok, cont;
  int total_iterations;
  double rho;
  // compute_correlation sums [n, Sf, Sg, Sf2, Sg2, Sfg] gathered by the level-0 passes of the IC-LM solver at the
  // parameters of the last pass (rho_try) and at the accepted parameters (rho_acc); rho_have: rho_acc belongs to S.t
  double rho_try[6], rho_acc[6];
  int rho_have;
};

struct Ctx {
  const EccConfig *cfg;
  const EccFrame *frame;
  Shared *S;
  int rank, csize, tid;
  int buf;
  unsigned tl_par, tl_used;   // TMA-staged pass: per stage, parity of the uses so far / stage used at all (uniform across the CTA)
};

// ------------------------------------------------------------------------------------------------
// scalar parameter algebra (thread 0 of every CTA)
// ------------------------------------------------------------------------------------------------
__device__ void xf_scale(ssk_transform &t, double f) {
  switch (t.motion_type) {
    case SSK_MOTION_TRANSLATION:                       // c_image_transform.cc:130-134
      t.params[0] = (float)(t.params[0] * f); t.params[1] = (float)(t.params[1] * f);
      break;
    case SSK_MOTION_EUCLIDEAN:
    case SSK_MOTION_SCALED_EUCLIDEAN:                  // c_image_transform.cc:502-507 (T and C)
      t.params[0] = (float)(t.params[0] * f); t.params[1] = (float)(t.params[1] * f);
      t.aux[0] = (float)(t.aux[0] * f); t.aux[1] = (float)(t.aux[1] * f);
      break;
    case SSK_MOTION_AFFINE:                            // c_image_transform.cc:911-915
      t.params[2] = (float)(t.params[2] * f); t.params[5] = (float)(t.params[5] * f);
      break;
    default:                                           // c_image_transform.cc:1186-1194
      t.params[2] = (float)(t.params[2] * f); t.params[5] = (float)(t.params[5] * f);
      t.params[6] = (float)(t.params[6] / f); t.params[7] = (float)(t.params[7] / f);
      break;
  }
}

__device__ void xf_get_translation(const ssk_transform &t, float &tx, float &ty) {
  switch (t.motion_type) {
    case SSK_MOTION_AFFINE: tx = t.params[2]; ty = t.params[5]; break;
    case SSK_MOTION_HOMOGRAPHY: tx = t.params[2] / t.aux[2]; ty = t.params[5] / t.aux[2]; break;
    default: tx = t.params[0]; ty = t.params[1]; break;
  }
}

__device__ void xf_set_translation(ssk_transform &t, float tx, float ty) {
  switch (t.motion_type) {
    case SSK_MOTION_AFFINE: t.params[2] = tx; t.params[5] = ty; break;
    case SSK_MOTION_HOMOGRAPHY: t.params[2] = tx * t.aux[2]; t.params[5] = ty * t.aux[2]; break;
    default: t.params[0] = tx; t.params[1] = ty; break;
  }
}

__device__ inline float sqf(float x) { return x * x; }

// c_image_transform::eps(dp, size)
__device__ double xf_eps(int type, const float *dp, int w, int h, float fixed_scale = 1.0f) {
  switch (type) {
    case SSK_MOTION_TRANSLATION:                       // c_image_transform.cc:136-139
      return sqrt((double)(dp[0] * dp[0] + dp[1] * dp[1]));
    case SSK_MOTION_EUCLIDEAN:
    case SSK_MOTION_SCALED_EUCLIDEAN: {                // c_image_transform.cc:509-523
      const float da = dp[2], ds = type == SSK_MOTION_SCALED_EUCLIDEAN ? dp[3] : fixed_scale;  // fixed scale: get_parameters returns _scale
      const float sa = (float)sin((double)da);
      return (double)sqrtf(sqf(dp[0]) + sqf(dp[1]) + sqf(w * sa) + sqf(h * sa) + sqf((float)max(w, h) * ds));
    }
    case SSK_MOTION_AFFINE:                            // c_image_transform.cc:917-924
      return (double)sqrtf(sqf(w * dp[0]) + sqf(h * dp[1]) + sqf(dp[2]) + sqf(w * dp[3]) + sqf(h * dp[4]) + sqf(dp[5]));
    default:                                           // c_image_transform.cc:1196-1205
      return (double)sqrtf(sqf(dp[2]) + sqf(dp[5]) + sqf(w * dp[0]) + sqf(h * dp[1]) + sqf(w * dp[3]) + sqf(h * dp[4]));
  }
}

// cv::invertAffineTransform on CV_32F data, bit-exact against cv2 4.13 (oracle/cvmodel.py::invert_affine_f32):
// determinant, its reciprocal and the 2x2 part in float; the translation from double products of the rounded
// 2x2 entries.
__device__ void invert_affine(const float *M, float *iM) {
  const float D = __fsub_rn(__fmul_rn(M[0], M[4]), __fmul_rn(M[1], M[3]));
  const float Di = D != 0.f ? __fdiv_rn(1.0f, D) : 0.f;
  const float A11 = __fmul_rn(M[4], Di), A22 = __fmul_rn(M[0], Di), A12 = __fmul_rn(-M[1], Di), A21 = __fmul_rn(-M[3], Di);
  iM[0] = A11; iM[1] = A12; iM[2] = (float)(-((double)A11 * (double)M[2] + (double)A12 * (double)M[5]));
  iM[3] = A21; iM[4] = A22; iM[5] = (float)(-((double)A21 * (double)M[2] + (double)A22 * (double)M[5]));
}

// 3x3 inverse: cofactors and determinant in double, result float (cv::invert closed form)
__device__ void invert3x3(const float *a, float *b) {
  const double a00 = a[0], a01 = a[1], a02 = a[2], a10 = a[3], a11 = a[4], a12 = a[5], a20 = a[6], a21 = a[7], a22 = a[8];
  double d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
  d = d != 0 ? 1. / d : 0;
  b[0] = (float)((a11 * a22 - a12 * a21) * d); b[1] = (float)((a02 * a21 - a01 * a22) * d); b[2] = (float)((a01 * a12 - a02 * a11) * d);
  b[3] = (float)((a12 * a20 - a10 * a22) * d); b[4] = (float)((a00 * a22 - a02 * a20) * d); b[5] = (float)((a02 * a10 - a00 * a12) * d);
  b[6] = (float)((a10 * a21 - a11 * a20) * d); b[7] = (float)((a01 * a20 - a00 * a21) * d); b[8] = (float)((a00 * a11 - a01 * a10) * d);
}

__device__ void euclid_matrix(float tX, float tY, float ang, float scl, float cX, float cY, float *M /*2x3*/) {
  // lambda at c_image_transform.cc:768-782
  const float sa = (float)sin((double)ang), ca = (float)cos((double)ang);
  const float sc = __fmul_rn(scl, ca), ss = __fmul_rn(scl, sa);
  M[0] = sc; M[1] = -ss; M[2] = __fadd_rn(__fsub_rn(tX, __fmul_rn(sc, cX)), __fmul_rn(ss, cY));
  M[3] = ss; M[4] = sc;  M[5] = __fsub_rn(__fsub_rn(tY, __fmul_rn(ss, cX)), __fmul_rn(sc, cY));
}

// newparams = transform->invert_and_compose(p, dp)
__device__ void xf_invert_and_compose(const ssk_transform &t, const float *dp, float *out) {
  switch (t.motion_type) {
    case SSK_MOTION_TRANSLATION:                       // c_image_transform.h:164-167
      out[0] = t.params[0] - dp[0]; out[1] = t.params[1] - dp[1];
      break;
    case SSK_MOTION_AFFINE: {                          // c_image_transform.h:312-318
      float a[6], s[6];
      invert_affine(t.params, a);
      for (int i = 0; i < 6; ++i) s[i] = a[i] + dp[i];
      invert_affine(s, out);
      break;
    }
    case SSK_MOTION_HOMOGRAPHY: {                      // c_image_transform.h:379-384
      float m[9], inv1[9], s[9], aii[9];
      for (int i = 0; i < 8; ++i) m[i] = t.params[i];
      m[8] = t.aux[2];
      invert3x3(m, inv1);
      for (int i = 0; i < 8; ++i) s[i] = inv1[i] + dp[i];
      s[8] = inv1[8];
      invert3x3(s, aii);
      const float k = 1.0f / aii[8];
      for (int i = 0; i < 8; ++i) out[i] = aii[i] * k;
      break;
    }
    default: {                                         // c_image_transform.cc:736-833
      const bool fix_scale = t.motion_type == SSK_MOTION_EUCLIDEAN;
      const float Tx = t.params[0], Ty = t.params[1], angle = t.params[2];
      const float scale = fix_scale ? t.aux[3] : t.params[3];
      const float Cx = t.aux[0], Cy = t.aux[1];
      const float scale_dp = fix_scale ? 1.0f : 1.0f + dp[3];
      float Mp[6], Mdp[6], Mi[6];
      euclid_matrix(Tx, Ty, angle, scale, Cx, Cy, Mp);
      euclid_matrix(dp[0], dp[1], dp[2], scale_dp, Cx, Cy, Mdp);
      invert_affine(Mdp, Mi);
      // M_res = Mp * Mdp_inv (3x3 float product, last row 0 0 1)
      // Matx33f product, s += a(i,k)*b(k,j) in order (third row of Mdp_inv is 0 0 1)
      const float m00 = __fadd_rn(__fmul_rn(Mp[0], Mi[0]), __fmul_rn(Mp[1], Mi[3]));
      const float m02 = __fadd_rn(__fadd_rn(__fmul_rn(Mp[0], Mi[2]), __fmul_rn(Mp[1], Mi[5])), Mp[2]);
      const float m10 = __fadd_rn(__fmul_rn(Mp[3], Mi[0]), __fmul_rn(Mp[4], Mi[3]));
      const float m12 = __fadd_rn(__fadd_rn(__fmul_rn(Mp[3], Mi[2]), __fmul_rn(Mp[4], Mi[5])), Mp[5]);
      const float rs = fix_scale ? scale : __fsqrt_rn(__fadd_rn(__fmul_rn(m00, m00), __fmul_rn(m10, m10)));
      const float ra = (float)atan2((double)m10, (double)m00);
      const float rca = (float)cos((double)ra), rsa = (float)sin((double)ra);
      out[0] = __fsub_rn(__fadd_rn(m02, __fmul_rn(__fmul_rn(rs, rca), Cx)), __fmul_rn(__fmul_rn(rs, rsa), Cy));
      out[1] = __fadd_rn(__fadd_rn(m12, __fmul_rn(__fmul_rn(rs, rsa), Cx)), __fmul_rn(__fmul_rn(rs, rca), Cy));
      out[2] = ra;
      if (!fix_scale) out[3] = rs;
      break;
    }
  }
}

// hal::Cholesky32f (CholImpl<float>, OpenCV modules/core/src/matrix_decomp.cpp), bit-exact emulation
// (oracle/cvmodel.py::chol_solve_f32 is checked against cv2.solve / cv2.invert): float storage, float
// products, double accumulators, reciprocal diagonal.  A (m x m) is overwritten, b (m x n) holds the solution.
__device__ bool cv_cholesky32f(float *A, int m, float *b, int n) {
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < i; ++j) {
      double s = A[i * m + j];
      for (int k = 0; k < j; ++k) s -= (double)__fmul_rn(A[i * m + k], A[j * m + k]);
      A[i * m + j] = (float)(s * (double)A[j * m + j]);
    }
    double s = A[i * m + i];
    for (int k = 0; k < i; ++k) { const double t = A[i * m + k]; s -= t * t; }
    if (s < (double)FLT_EPSILON) return false;
    A[i * m + i] = (float)(1. / sqrt(s));
  }
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = b[i * n + j];
      for (int k = 0; k < i; ++k) s -= (double)__fmul_rn(A[i * m + k], b[k * n + j]);
      b[i * n + j] = (float)(s * (double)A[i * m + i]);
    }
  for (int i = m - 1; i >= 0; --i)
    for (int j = 0; j < n; ++j) {
      double s = b[i * n + j];
      for (int k = m - 1; k > i; --k) s -= (double)__fmul_rn(A[k * m + i], b[k * n + j]);
      b[i * n + j] = (float)(s * (double)A[i * m + i]);
    }
  return true;
}

__device__ inline double det3d(const float *m) {
  return (double)m[0] * ((double)m[4] * m[8] - (double)m[5] * m[7]) - (double)m[1] * ((double)m[3] * m[8] - (double)m[5] * m[6]) +
         (double)m[2] * ((double)m[3] * m[7] - (double)m[4] * m[6]);
}

// cv::solve(H, b, x, DECOMP_CHOLESKY) on CV_32F data: closed forms in double for n <= 3 (bit-exact for n = 2;
// the n = 3 form is evaluated entirely in double here), hal::Cholesky32f otherwise.  x = 0 on failure.
__device__ bool cv_solve(int M, const float *H, const float *b, float *x) {
  if (M == 2) {
    double d = (double)H[0] * H[3] - (double)H[1] * H[2];
    if (d == 0.) { x[0] = x[1] = 0.f; return false; }
    d = 1. / d;
    x[0] = (float)(((double)b[0] * H[3] - (double)b[1] * H[1]) * d);
    x[1] = (float)(((double)b[1] * H[0] - (double)b[0] * H[2]) * d);
    return true;
  }
  if (M == 3) {
    double d = det3d(H);
    if (d == 0.) { x[0] = x[1] = x[2] = 0.f; return false; }
    d = 1. / d;
    const double S00 = H[0], S01 = H[1], S02 = H[2], S10 = H[3], S11 = H[4], S12 = H[5], S20 = H[6], S21 = H[7], S22 = H[8];
    const double b0 = b[0], b1 = b[1], b2 = b[2];
    x[0] = (float)(d * (b0 * (S11 * S22 - S12 * S21) - S01 * (b1 * S22 - S12 * b2) + S02 * (b1 * S21 - S11 * b2)));
    x[1] = (float)(d * (S00 * (b1 * S22 - S12 * b2) - b0 * (S10 * S22 - S12 * S20) + S02 * (S10 * b2 - b1 * S20)));
    x[2] = (float)(d * (S00 * (S11 * b2 - b1 * S21) - S01 * (S10 * b2 - b1 * S20) + b0 * (S10 * S21 - S11 * S20)));
    return true;
  }
  float A[64];
  for (int i = 0; i < M * M; ++i) A[i] = H[i];
  for (int i = 0; i < M; ++i) x[i] = b[i];
  if (!cv_cholesky32f(A, M, x, 1)) { for (int i = 0; i < M; ++i) x[i] = 0.f; return false; }
  return true;
}

// cv::invert(H, Hinv, DECOMP_CHOLESKY) on CV_32F data
__device__ bool cv_invert(int M, const float *H, float *Hi) {
  if (M == 2) {
    const double d = (double)H[0] * H[3] - (double)H[1] * H[2];
    if (d == 0.) return false;
    const float di = (float)(1. / d);   // the SIMD path multiplies in float
    Hi[0] = __fmul_rn(H[3], di); Hi[1] = -__fmul_rn(H[1], di); Hi[2] = -__fmul_rn(H[2], di); Hi[3] = __fmul_rn(H[0], di);
    return true;
  }
  if (M == 3) {
    double d = det3d(H);
    if (d == 0.) return false;
    d = 1. / d;
    const double S00 = H[0], S01 = H[1], S02 = H[2], S10 = H[3], S11 = H[4], S12 = H[5], S20 = H[6], S21 = H[7], S22 = H[8];
    Hi[0] = (float)((S11 * S22 - S12 * S21) * d); Hi[1] = (float)((S02 * S21 - S01 * S22) * d); Hi[2] = (float)((S01 * S12 - S02 * S11) * d);
    Hi[3] = (float)((S12 * S20 - S10 * S22) * d); Hi[4] = (float)((S00 * S22 - S02 * S20) * d); Hi[5] = (float)((S02 * S10 - S00 * S12) * d);
    Hi[6] = (float)((S10 * S21 - S11 * S20) * d); Hi[7] = (float)((S01 * S20 - S00 * S21) * d); Hi[8] = (float)((S00 * S11 - S01 * S10) * d);
    return true;
  }
  float A[64];
  for (int i = 0; i < M * M; ++i) { A[i] = H[i]; Hi[i] = (i / M == i % M) ? 1.f : 0.f; }
  return cv_cholesky32f(A, M, Hi, M);
}

__device__ void make_jcoef(const ssk_transform &t, JCoef &j) {
  for (int i = 0; i < 8; ++i) j.c[i] = 0.f;
  switch (t.motion_type) {
    case SSK_MOTION_EUCLIDEAN:
    case SSK_MOTION_SCALED_EUCLIDEAN:
      j.c[0] = t.motion_type == SSK_MOTION_SCALED_EUCLIDEAN ? t.params[3] : t.aux[3];
      j.c[1] = (float)cos((double)t.params[2]); j.c[2] = (float)sin((double)t.params[2]);
      j.c[3] = t.aux[0]; j.c[4] = t.aux[1];
      break;
    case SSK_MOTION_HOMOGRAPHY:
      for (int i = 0; i < 8; ++i) j.c[i] = t.params[i];
      break;
    default: break;
  }
}

// c_image_transform::create_steepest_descent_images, one pixel.  Explicit single-rounding operations in the
// reference's operand order (no FMA contraction), so that J is bit-identical to the oracle's float arithmetic.
template <int TYPE>
__device__ __forceinline__ void eval_J(const JCoef &jc, float x, float y, float gx, float gy, float *J) {
  if (TYPE == SSK_MOTION_TRANSLATION) {                // c_image_transform.cc:262-269
    J[0] = gx; J[1] = gy;
  } else if (TYPE == SSK_MOTION_AFFINE) {              // c_image_transform.cc:1054-1067
    J[0] = __fmul_rn(gx, x); J[1] = __fmul_rn(gx, y); J[2] = gx;
    J[3] = __fmul_rn(gy, x); J[4] = __fmul_rn(gy, y); J[5] = gy;
  } else if (TYPE == SSK_MOTION_HOMOGRAPHY) {          // c_image_transform.cc:1339-1363
    const float den = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(x, jc.c[6]), __fmul_rn(y, jc.c[7])), 1.f));
    const float hatX = __fmul_rn(-__fadd_rn(__fadd_rn(__fmul_rn(x, jc.c[0]), __fmul_rn(y, jc.c[1])), jc.c[2]), den);
    const float hatY = __fmul_rn(-__fadd_rn(__fadd_rn(__fmul_rn(x, jc.c[3]), __fmul_rn(y, jc.c[4])), jc.c[5]), den);
    const float ggx = __fmul_rn(gx, den), ggy = __fmul_rn(gy, den);
    const float gg = __fadd_rn(__fmul_rn(hatX, ggx), __fmul_rn(hatY, ggy));
    J[0] = __fmul_rn(ggx, x); J[1] = __fmul_rn(ggx, y); J[2] = ggx;
    J[3] = __fmul_rn(ggy, x); J[4] = __fmul_rn(ggy, y); J[5] = ggy;
    J[6] = __fmul_rn(gg, x); J[7] = __fmul_rn(gg, y);
  } else {                                             // c_image_transform.cc:636-665
    const float xx = __fsub_rn(x, jc.c[3]), yy = __fsub_rn(y, jc.c[4]);
    const float ca = jc.c[1], sa = jc.c[2];
    const float t1 = __fadd_rn(__fmul_rn(sa, xx), __fmul_rn(ca, yy));   // sa*xx + ca*yy
    const float t2 = __fsub_rn(__fmul_rn(ca, xx), __fmul_rn(sa, yy));   // ca*xx - sa*yy
    J[0] = gx; J[1] = gy;
    J[2] = __fmul_rn(jc.c[0], __fadd_rn(__fmul_rn(-gx, t1), __fmul_rn(gy, t2)));
    if (TYPE == SSK_MOTION_SCALED_EUCLIDEAN) J[3] = __fadd_rn(__fmul_rn(gx, t2), __fmul_rn(gy, t1));
  }
}

// ------------------------------------------------------------------------------------------------
// cluster-wide deterministic reduction of NS per-thread partial sums
// ------------------------------------------------------------------------------------------------
template <int NS, typename T>
__device__ void cluster_reduce(Ctx &c, const T (&acc)[NS]) {
  Shared &S = *c.S;
  const int lane = c.tid & 31, warp = c.tid >> 5;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const double v = warp_sum((double)acc[k]);
    if (lane == 0) S.wpart[warp][k] = v;
  }
  __syncthreads();
  if (c.tid < NS) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += S.wpart[w][c.tid];
    S.part[c.buf][c.tid] = s;
  }
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  if (c.tid < NS) {
    double s = 0;
    for (int r = 0; r < c.csize; ++r) s += *cluster.map_shared_rank(&S.part[c.buf][c.tid], r);
    S.tot[c.tid] = s;
  }
  __syncthreads();
  c.buf ^= 1;
}

__device__ __forceinline__ Img level_image(const Ctx &c, int lvl) {
  const EccLevel &L = c.cfg->lv[lvl];
  Img im;
  im.data = c.frame->pyr + L.cur_off;
  im.step = (int64_t)L.cols * 4; im.rows = L.rows; im.cols = L.cols; im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
  return im;
}

// thread 0 of rank 0, frame 0: append a trace record
__device__ void trace_rec(const Ctx &c, int lvl, int pass) {
  const EccConfig &cfg = *c.cfg;
  if (!cfg.trace || c.rank != 0 || blockIdx.x / c.csize != 0) return;
  const int i = *cfg.trace_count;
  if (i >= cfg.trace_capacity) return;
  *cfg.trace_count = i + 1;
  const Shared &S = *c.S;
  float *r = cfg.trace + (size_t)i * kTraceRec;
  r[0] = (float)lvl; r[1] = (float)pass; r[2] = (float)S.num_it; r[3] = (float)S.err; r[4] = (float)S.newerr;
  r[5] = (float)S.lambda; r[6] = (float)S.eps; r[7] = (float)S.tot[1];
  for (int k = 0; k < 8; ++k) { r[8 + k] = S.t.params[k]; r[16 + k] = S.tq.params[k]; r[24 + k] = S.deltap[k]; r[32 + k] = S.v[k]; }
}

__device__ void trace_hp(const Ctx &c, int lvl, int M) {
  const EccConfig &cfg = *c.cfg;
  if (!cfg.trace || c.rank != 0 || blockIdx.x / c.csize != 0) return;
  const int i = *cfg.trace_count;
  if (i >= cfg.trace_capacity) return;
  *cfg.trace_count = i + 1;
  float *r = cfg.trace + (size_t)i * kTraceRec;
  r[0] = (float)lvl; r[1] = 9.f; r[2] = (float)M; r[3] = 0.f;
  for (int k = 0; k < 36; ++k) r[4 + k] = k < M * M ? c.S->Hp[k] : 0.f;
}

// thread 0: publish the parameters of the next pass
__device__ void set_pass_params(Shared &S, const ssk_transform &q) {
  S.tq = q;
  S.map = make_mapcoef(q);
}

// ------------------------------------------------------------------------------------------------
// per-pixel plumbing of the passes
// ------------------------------------------------------------------------------------------------
// Strided row-major walk over a level (pixel i = start, start + stride, ...) that keeps (x, y) incrementally:
// no per-pixel integer division.
struct Walk {
  int i, x, y, dx, dy, stride, cols;
  __device__ __forceinline__ Walk(int start, int stride_, int cols_) : stride(stride_), cols(cols_) {
    i = start; y = start / cols_; x = start - y * cols_;
    dy = stride_ / cols_; dx = stride_ - dy * cols_;
  }
  __device__ __forceinline__ void next() {
    i += stride; x += dx; y += dy;
    if (x >= cols) { x -= cols; ++y; }
  }
};
// The same walk with the coordinates also kept as floats (exact: they are integers below 2^24), so that the per-pixel
// int-to-float conversions (XU pipe) become additions.
#ifndef SSK_ECC_FWALK
#define SSK_ECC_FWALK 1
#endif
struct WalkF : Walk {
  float fx, fy, fdx, fdy, fcols;
  __device__ __forceinline__ WalkF(int start, int stride_, int cols_) : Walk(start, stride_, cols_) {
    fx = (float)x; fy = (float)y; fdx = (float)dx; fdy = (float)dy; fcols = (float)cols_;
  }
  __device__ __forceinline__ void next() {
#if SSK_ECC_FWALK
    i += stride; x += dx; y += dy; fx += fdx; fy += fdy;
    if (x >= cols) { x -= cols; ++y; fx -= fcols; fy += 1.0f; }
#else
    Walk::next(); fx = (float)x; fy = (float)y;
#endif
  }
};

// Pixels of a level one CTA of the cluster visits in a pass.  SSK_ECC_BANDS (default): rank r owns the contiguous band
// [r * chunk, (r + 1) * chunk) and walks it NT pixels at a time, so that the source row a bilinear tap pair touches for
// output row y is still in L1 when row y + 1 needs it (the interleaved assignment sent consecutive rows to different
// SMs).  The order of the per-thread partial sums changes with it, their fixed-order double reduction does not.
// compute_correlation's sums gathered by the level-0 passes of the IC-LM solver: 0 = never (separate pass_rho), 1 = by every
// level-0 pass, 2 = by the trial passes only (the accepted trial is the one the solver ends on; a level that accepts no trial
// falls back to pass_rho)
#ifndef SSK_ECC_RHO_FUSION
#define SSK_ECC_RHO_FUSION 2
#endif
__device__ __forceinline__ bool getenv_no_rho_fusion() { return !SSK_ECC_RHO_FUSION; }
#ifndef SSK_ECC_BANDS
#define SSK_ECC_BANDS 1
#endif
struct Band { int start, stride, end; };
__device__ __forceinline__ Band pass_band(int rank, int csize, int tid, int n) {
  Band b;
#if SSK_ECC_BANDS
  const int chunk = (((n + csize - 1) / csize) + 31) & ~31;
  b.start = rank * chunk + tid; b.stride = NT; b.end = min(n, (rank + 1) * chunk);
#else
  b.start = rank * NT + tid; b.stride = csize * NT; b.end = n;
#endif
  return b;
}

// keeps a base pointer as one 64-bit register value (stops the compiler from re-adding its parts per access)
template <typename T> __device__ __forceinline__ const T *opaque_ptr(const T *p) { asm volatile("" : "+l"(p)); return p; }

// cvRound(v * 32).  Round-to-nearest-even through the 1.5 * 2^23 addend (FADD + IADD on the FP32 / integer pipes instead
// of F2I on the quarter-rate conversion unit): identical to __float2int_rn for |32 v| < 2^22; beyond that (131 072 px
// outside the image) both forms give coordinates every validity test rejects and every sampler clamps.
#ifndef SSK_ECC_FASTROUND
#define SSK_ECC_FASTROUND 1
#endif
__device__ __forceinline__ int cvround32(float v) {
#if SSK_ECC_FASTROUND
  return __float_as_int(__fadd_rn(__fmul_rn(v, 32.0f), 12582912.0f)) - 0x4B400000;
#else
  return __float2int_rn(__fmul_rn(v, 32.0f));
#endif
}
// (float)(s & 31) / 32 without an integer-to-float conversion: 1 + f / 32 assembled in the mantissa, minus one (exact)
__device__ __forceinline__ float frac32(int s) {
#if SSK_ECC_FASTROUND
  // ((s << 18) & 0x7c0000) | 0x3f800000 as one LOP3 (the constant 1.0f comes from a register the compiler cannot fold)
  unsigned one = 0x3F800000u, r;
  asm("" : "+r"(one));
  asm("lop3.b32 %0, %1, 0x7c0000, %2, 0xEA;" : "=r"(r) : "r"((unsigned)s << 18), "r"(one));
  return __fsub_rn(__uint_as_float(r), 1.0f);
#else
  return (float)(s & 31) * 0.03125f;
#endif
}
// cvRound(u) in [0, n) decided on the float itself (round half to even at both ends): u >= -0.5 and u <= hi with
// hi = n - 0.5 when n - 1 is even, the float below it otherwise
__device__ __forceinline__ float nearest_hi(int n) {
  const float h = (float)n - 0.5f;
  return (n & 1) ? h : __int_as_float(__float_as_int(h) - 1);
}

// map_xy with the map kind fixed at compile time (the passes are instantiated per transform type)
template <int TYPE> struct MapKind { static constexpr int MT = MAP_EUCLIDEAN; };
template <> struct MapKind<SSK_MOTION_TRANSLATION> { static constexpr int MT = MAP_TRANSLATION; };
template <> struct MapKind<SSK_MOTION_AFFINE> { static constexpr int MT = MAP_AFFINE; };
template <> struct MapKind<SSK_MOTION_HOMOGRAPHY> { static constexpr int MT = MAP_HOMOGRAPHY; };

template <int MT>
__device__ __forceinline__ void map_xy_t(const MapCoef &m, float x, float y, float &u, float &v) {
  if (MT == MAP_TRANSLATION) {
    u = __fadd_rn(x, m.c[0]);
    v = __fadd_rn(y, m.c[1]);
  } else if (MT == MAP_AFFINE) {
    u = __fadd_rn(__fadd_rn(__fmul_rn(m.c[0], x), __fmul_rn(m.c[1], y)), m.c[2]);
    v = __fadd_rn(__fadd_rn(__fmul_rn(m.c[3], x), __fmul_rn(m.c[4], y)), m.c[5]);
  } else {
    map_xy(m, x, y, u, v);
  }
}

// valid255_linear on pre-quantised coordinates (sx, sy = cvRound(32 u), cvRound(32 v))
__device__ __forceinline__ bool lin_valid(int sx, int sy, int cols, int rows) {
  const int ix = sx >> 5, iy = sy >> 5;
  return (unsigned)ix < (unsigned)cols && (unsigned)iy < (unsigned)rows &&
         ((sx & 31) == 0 || ix + 1 < cols) && ((sy & 31) == 0 || iy + 1 < rows);
}

// cv::remap of a CV_8UC1 mask with INTER_LINEAR and BORDER_CONSTANT 0 at the pre-quantised coordinates (sx, sy): the 15-bit
// fixed-point bilinear table of a 1/32-px fraction is exactly 32 (32 - fx | fx)(32 - fy | fy) (sum 32768, no correction),
// the result (sum + 2^14) >> 15 (FixedPtCast).  The reference compares it with 255 (forward-additive, ecc2.cc:1311),
// 250 (ecc_remap, ecc2.cc:205-216) or 254 (compute_correlation, ecc2.cc:117).
__device__ __forceinline__ int mask_lin_value(const uint8_t *__restrict__ m, int cols, int rows, int sx, int sy) {
  const int ix = sx >> 5, iy = sy >> 5, fx = sx & 31, fy = sy & 31;
  const bool x0 = (unsigned)ix < (unsigned)cols, x1 = (unsigned)(ix + 1) < (unsigned)cols;
  const bool y0 = (unsigned)iy < (unsigned)rows, y1 = (unsigned)(iy + 1) < (unsigned)rows;
  const int m00 = x0 && y0 ? (int)m[iy * cols + ix] : 0, m01 = x1 && y0 ? (int)m[iy * cols + ix + 1] : 0;
  const int m10 = x0 && y1 ? (int)m[(iy + 1) * cols + ix] : 0, m11 = x1 && y1 ? (int)m[(iy + 1) * cols + ix + 1] : 0;
  const int sum = (32 - fx) * (32 - fy) * m00 + fx * (32 - fy) * m01 + (32 - fx) * fy * m10 + fx * fy * m11;
  return (32 * sum + (1 << 14)) >> 15;
}

// cv::remap INTER_LINEAR of a dense CV_32FC1 level with BORDER_REPLICATE: clamped tap coordinates are the
// replicate border, the arithmetic is sample_linear's (bit-exact against cv2).  Unconditionally safe to call
// (every address is in bounds), which lets the passes run branch-free.  For pixels that passed lin_valid the
// out-of-range taps carry zero weight, so the result also equals the BORDER_CONSTANT sample.
__device__ __forceinline__ float lin_sample(const float *__restrict__ p, int cols, int rows, int sx, int sy) {
  const int ix = sx >> 5, iy = sy >> 5;
  const float tx = frac32(sx), ty = frac32(sy);
  const float wx0 = 1.0f - tx, wy0 = 1.0f - ty;
  const int x0 = min(max(ix, 0), cols - 1), x1 = min(max(ix + 1, 0), cols - 1);
  const int y0 = min(max(iy, 0), rows - 1), y1 = min(max(iy + 1, 0), rows - 1);
  // unsigned 32-bit offsets (a level has far fewer than 2^31 pixels): one wide multiply-add per tap address
  const unsigned o0 = (unsigned)(y0 * cols), o1 = (unsigned)(y1 * cols);
  const float s00 = __ldg(p + (o0 + (unsigned)x0)), s01 = __ldg(p + (o0 + (unsigned)x1)), s10 = __ldg(p + (o1 + (unsigned)x0)), s11 = __ldg(p + (o1 + (unsigned)x1));
  float out = __fadd_rn(__fmul_rn(s00, __fmul_rn(wy0, wx0)), __fmul_rn(s01, __fmul_rn(wy0, tx)));
  out = __fadd_rn(out, __fmul_rn(s10, __fmul_rn(ty, wx0)));
  out = __fadd_rn(out, __fmul_rn(s11, __fmul_rn(ty, tx)));
  return out;
}

// ------------------------------------------------------------------------------------------------
// pass of the inverse-compositional solvers: sums = [ |rhs|^2, #valid, J_i . rhs ]
//   lm_masks = true : c_ecclm_inverse_compositional::compute_rhs (ecc2.cc:1894-1917): mask by INTER_NEAREST remap of the
//                     inverted current mask with constant border 255 -> a pixel is bad iff its rounded source
//                     coordinate falls outside the current image (no user current mask)
//   lm_masks = false: ecc_remap (ecc2.cc:178-219): bilinear remap of the all-255 mask >= 250
// ------------------------------------------------------------------------------------------------
//   RHO = true (level 0 of the IC-LM solver when the registration checks the correlation afterwards): the same pass also
//   gathers compute_correlation's sums (ecc2.cc:65-137) for its map - the bilinear sample is the one the residual uses,
//   only the validity rule differs (remap(255) >= 254, i.e. lin_valid) - so that the separate correlation pass over
//   level 0 is not needed when the solver ends on parameters it has evaluated (it always does).
template <int TYPE, bool RHO>
__device__ void pass_ic(Ctx &c, int lvl, bool lm_masks) {
  constexpr int M = NParams<TYPE>::M;
  constexpr int NS = 2 + M + (RHO ? 6 : 0);
  constexpr int OR = 2 + M;            // offset of the correlation sums
  Shared &S = *c.S;
  const EccLevel &L = c.cfg->lv[lvl];
  const float *__restrict__ cur = opaque_ptr(c.frame->pyr + L.cur_off);
  const float *__restrict__ ref = L.ref, *__restrict__ gxp = L.gx, *__restrict__ gyp = L.gy;
  // c_ecc_inverse_compositional leaves the reference mask out of rhs / CMA (ecc2.cc:1752-1763: only the remapped
  // current mask); c_ecclm_inverse_compositional ORs it in (ecc2.cc:1901-1904)
  const uint8_t *__restrict__ rmask = lm_masks ? L.refmask : nullptr;
  const uint8_t *__restrict__ cm = c.frame->cmask ? c.frame->cmask + L.cur_off : nullptr;   // current mask of this level
  const int cols = L.cols, rows = L.rows;
  const MapCoef m = S.map;
  const JCoef jc = S.jc;
  double acc[NS];   // Mat::dot / norm accumulate in double
#pragma unroll
  for (int k = 0; k < NS; ++k) acc[k] = 0.0;
  int nvalid = 0, nrho = 0;
  const uint8_t *__restrict__ rho_mask = RHO ? L.refmask : nullptr;
  const int n = cols * rows;
  // Branch-free body (invalid pixels contribute an exact 0.0), so that several pixels per thread are in flight.
  const Band bd = pass_band(c.rank, c.csize, c.tid, n);
  WalkF w(bd.start, bd.stride, cols);
  const float hx = nearest_hi(cols), hy = nearest_hi(rows);
#ifndef SSK_ECC_UNROLL_RHO
#define SSK_ECC_UNROLL_RHO 3
#endif
  constexpr int kUnroll = RHO ? SSK_ECC_UNROLL_RHO : 3;   // pixels in flight per thread (measured: 2 -> 3 is -4 % on the pass, 4 is slower)
#pragma unroll kUnroll
  for (; w.i < bd.end; w.next()) {
    const float x = w.fx, y = w.fy;
    float u, v;
    map_xy_t<MapKind<TYPE>::MT>(m, x, y, u, v);
    const int sx = cvround32(u), sy = cvround32(v);
    bool ok;
#if SSK_ECC_FASTROUND
    if (lm_masks) ok = u >= -0.5f && u <= hx && v >= -0.5f && v <= hy;   // cvRound(u), cvRound(v) inside the image
#else
    if (lm_masks) ok = (unsigned)__float2int_rn(u) < (unsigned)cols && (unsigned)__float2int_rn(v) < (unsigned)rows;
#endif
    else ok = lin_valid(sx, sy, cols, rows);
    if (cm) {
      // IC-LM: INTER_NEAREST remap of the inverted current mask, border 255 (ecc2.cc:1877-1884); IC: ecc_remap's
      // bilinear mask remap >= 250 (ecc2.cc:205-216)
      if (lm_masks) ok = ok && cm[min(max(__float2int_rn(v), 0), rows - 1) * cols + min(max(__float2int_rn(u), 0), cols - 1)] != 0;
      else ok = mask_lin_value(cm, cols, rows, sx, sy) >= 250;
    }
    if (rmask) ok = ok && rmask[w.i] != 0;
    const float g = lin_sample(cur, cols, rows, sx, sy);
    const float fref = __ldg(ref + w.i);
    const float rhs = g - fref;
    if (RHO) {
      bool okr = cm ? mask_lin_value(cm, cols, rows, sx, sy) >= 254 : lin_valid(sx, sy, cols, rows);
      if (rho_mask) okr = okr && rho_mask[w.i] != 0;
      const double gd = okr ? (double)g : 0.0, fd = okr ? (double)fref : 0.0;
      nrho += okr ? 1 : 0;
      acc[OR + 1] += fd; acc[OR + 2] += gd; acc[OR + 3] += fd * fd; acc[OR + 4] += gd * gd; acc[OR + 5] += fd * gd;
    }
    float J[M];
    eval_J<TYPE>(jc, x, y, __ldg(gxp + w.i), __ldg(gyp + w.i), J);
    if (TYPE == SSK_MOTION_HOMOGRAPHY) {   // J may be non-finite at a vanishing denominator: keep the branch
      if (ok) {
        acc[0] += (double)rhs * (double)rhs;
        ++nvalid;
#pragma unroll
        for (int k = 0; k < M; ++k) acc[2 + k] += (double)J[k] * (double)rhs;
      }
    } else {
      const double r = ok ? (double)rhs : 0.0;
      acc[0] += r * r;
      nvalid += ok ? 1 : 0;
#pragma unroll
      for (int k = 0; k < M; ++k) acc[2 + k] += (double)J[k] * r;
    }
  }
  acc[1] = (double)nvalid;
  if (RHO) acc[OR] = (double)nrho;
  cluster_reduce<NS>(c, acc);
  if (RHO) {
    if (c.tid < 6) c.S->rho_try[c.tid] = c.S->tot[OR + c.tid];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged form of pass_ic (same sums, same per-pixel arithmetic) for the large levels of a pyramid.
// The level is cut into TL_W x TL_H tiles, dealt to the CTAs of the cluster as contiguous ranges of the row-major tile
// order.  Per tile the copy engine brings four boxes into one of NSTAGE shared-memory stages: the reference, gx and gy tiles
// (2-D tensor maps of the level) and the window of the current image that holds every bilinear tap of the tile (3-D tensor
// map over all frame slots of the level: x, y, slot).  The window origin follows the map at the tile corners (affine-like
// maps are monotone along both axes).  One warp per tile (round robin) plans and issues the copies two tiles ahead; every
// warp waits on the stage's transaction barrier, evaluates its 4 pixels of the tile from shared memory and arrives on
// the stage's "empty" barrier - no CTA barrier inside a pass, no global load in the pixel loop.
// A pixel whose 2 x 2 tap footprint leaves the window or the image (the copy engine fills zeros there, the reference
// semantics want the replicated border) takes lin_sample's gather instead: same arithmetic, rare.
// ------------------------------------------------------------------------------------------------
#ifndef SSK_ECC_TMA
#define SSK_ECC_TMA 1
#endif
constexpr int TL_W = 64, TL_H = 4 * (NT / TL_W);          // 4 pixels per thread: rows tq, tq + NT/TL_W, ...
constexpr int TL_WW = 76, TL_WH = TL_H + 6;               // window: 2 taps + 3 columns of origin alignment + drift of the map over a tile
#ifndef SSK_ECC_NSTAGE
#define SSK_ECC_NSTAGE 4
#endif
constexpr int TL_NSTAGE = SSK_ECC_NSTAGE;
constexpr unsigned TL_TILE_BYTES = TL_W * TL_H * 4;
constexpr unsigned TL_WIN_BYTES = TL_WW * TL_WH * 4;
constexpr unsigned TL_WIN_SLOT = (TL_WIN_BYTES + 127) & ~127u;
constexpr unsigned TL_STAGE_BYTES = TL_WIN_SLOT + 3 * TL_TILE_BYTES;
constexpr unsigned TL_DYN_SMEM = TL_NSTAGE * TL_STAGE_BYTES + 128;
static_assert(TL_W == 64 && (TL_WW * 4) % 16 == 0, "tile geometry");

__device__ __forceinline__ void ecc_mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void ecc_mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ecc_mbar_arrive(unsigned mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void ecc_mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SSK_ECC_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SSK_ECC_MBAR_DONE;\n"
      "bra SSK_ECC_MBAR_WAIT;\n"
      "SSK_ECC_MBAR_DONE:\n"
      "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ecc_tma_2d(unsigned dst, const void *tmap, int x, int y, unsigned mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tmap), "r"(x), "r"(y), "r"(mbar) : "memory");
}
__device__ __forceinline__ void ecc_tma_3d(unsigned dst, const void *tmap, int x, int y, int z, unsigned mbar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
}

// true when level lvl of this launch can take the TMA-staged pass
__device__ __forceinline__ bool level_uses_tma(const Ctx &c, int lvl) {
  return SSK_ECC_TMA && lvl < c.cfg->tma_levels && c.frame->cmask == nullptr;
}

template <int TYPE, bool RHO>
__device__ void pass_ic_tma(Ctx &c, int lvl, bool lm_masks) {
  constexpr int M = NParams<TYPE>::M;
  constexpr int NS = 2 + M + (RHO ? 6 : 0);
  constexpr int OR = 2 + M;
  constexpr int MT = MapKind<TYPE>::MT;
  constexpr int RQ = NT / TL_W;        // row step between the pixels of a thread
  Shared &S = *c.S;
  const EccLevel &L = c.cfg->lv[lvl];
  const float *__restrict__ cur = opaque_ptr(c.frame->pyr + L.cur_off);
  const uint8_t *__restrict__ rmask = lm_masks ? L.refmask : nullptr;
  const uint8_t *__restrict__ rho_mask = RHO ? L.refmask : nullptr;
  const int cols = L.cols, rows = L.rows;
  const MapCoef m = S.map;
  const JCoef jc = S.jc;
  double acc[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) acc[k] = 0.0;
  int nvalid = 0, nrho = 0;
  const float hx = nearest_hi(cols), hy = nearest_hi(rows);
  const int ntx = (cols + TL_W - 1) / TL_W, nty = (rows + TL_H - 1) / TL_H;
  const int per = (ntx * nty + c.csize - 1) / c.csize;
  const int t0 = min(ntx * nty, c.rank * per), t1 = min(ntx * nty, t0 + per);
  const int ntiles = t1 - t0;
  const int tx = c.tid & (TL_W - 1), tq = c.tid / TL_W;
  const int lane = c.tid & 31, warp = c.tid >> 5;
  extern __shared__ unsigned char ecc_dyn_smem[];
  const unsigned char *dynp = ecc_dyn_smem + ((128u - ((unsigned)__cvta_generic_to_shared(ecc_dyn_smem) & 127u)) & 127u);
  const unsigned dyn0 = (unsigned)__cvta_generic_to_shared(dynp);
  const unsigned full0 = (unsigned)__cvta_generic_to_shared(&S.tl_full[0]), empty0 = (unsigned)__cvta_generic_to_shared(&S.tl_empty[0]);
  const int slot = (int)((c.frame->pyr - c.cfg->pyr_base) / c.cfg->pyr_floats);
  const unsigned char *tm = c.cfg->tm[lvl][0];

  // The stage barriers are initialised once per kernel (k_ecc) and keep their phase across passes: use u of stage s
  // completes phase u of both barriers; tl_par / tl_used carry the parities into this pass.
  const unsigned par0 = c.tl_par, used0 = c.tl_used;

  // one warp: plan local tile j and ask the copy engine for its four boxes
  auto issue = [&](int j) {
    const int s = j % TL_NSTAGE;
    const int t = t0 + j;
    const int tyi = t / ntx, txi = t - tyi * ntx;
    const int x0 = txi * TL_W, y0 = tyi * TL_H;
    const int x1 = min(x0 + TL_W - 1, cols - 1), y1 = min(y0 + TL_H - 1, rows - 1);
    float u, v;
    map_xy_t<MT>(m, (float)((lane & 1) ? x1 : x0), (float)((lane & 2) ? y1 : y0), u, v);
    float umin = u, umax = u, vmin = v, vmax = v;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o)); umax = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, o));
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o)); vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if (lane == 0) {
      // taps of a pixel: columns ix, ix + 1 with floor(u) - 1 <= ix <= floor(u) + 1 (1/32-px rounding, one ulp of slack
      // for the corner bound), likewise rows
      const bool finite = umin > -1.0e8f && vmin > -1.0e8f && umax < 1.0e8f && vmax < 1.0e8f;
      const int xl = finite ? (int)floorf(umin) - 1 : 0, yl = finite ? (int)floorf(vmin) - 1 : 0;
      const int ox = xl & ~3;                  // the copy engine wants the box origin 16-byte aligned along the row
      const bool fits = MT != MAP_HOMOGRAPHY && finite && (int)floorf(umax) + 2 - ox < TL_WW && (int)floorf(vmax) + 2 - yl < TL_WH;
      // window coordinates (wx, wy) of a pixel's first tap for which all four taps are inside window and image
      const int wx_lo = max(0, -ox), wx_hi = min(TL_WW - 1, cols - 1 - ox);
      const int wy_lo = max(0, -yl), wy_hi = min(TL_WH - 1, rows - 1 - yl);
      const unsigned up = ((par0 >> s) ^ (unsigned)(j / TL_NSTAGE)) & 1u;     // parity of this use of the stage
      if (j >= TL_NSTAGE || ((used0 >> s) & 1u)) ecc_mbar_wait(empty0 + 8 * s, up ^ 1u);
      S.tl_meta[s][0] = make_int4(ox, yl, wx_lo, fits ? max(0, wx_hi - wx_lo) : 0);
      S.tl_meta[s][1] = make_int4(wy_lo, max(0, wy_hi - wy_lo), 0, 0);
      const unsigned st = dyn0 + s * TL_STAGE_BYTES, mb = full0 + 8 * s;
      ecc_mbar_expect_tx(mb, 3 * TL_TILE_BYTES + (fits ? TL_WIN_BYTES : 0u));
      ecc_tma_2d(st + TL_WIN_SLOT, tm + 128, x0, y0, mb);
      ecc_tma_2d(st + TL_WIN_SLOT + TL_TILE_BYTES, tm + 256, x0, y0, mb);
      ecc_tma_2d(st + TL_WIN_SLOT + 2 * TL_TILE_BYTES, tm + 384, x0, y0, mb);
      if (fits) ecc_tma_3d(st, tm, ox, yl, slot, mb);
    }
    __syncwarp();
  };
  constexpr int AHEAD = TL_NSTAGE - 2;    // tiles issued ahead of the one being evaluated
  for (int j = 0; j < min(AHEAD, ntiles); ++j)
    if (warp == j % NW) issue(j);

#pragma unroll 1
  for (int j = 0; j < ntiles; ++j) {
    if (j + AHEAD < ntiles && warp == (j + AHEAD) % NW) issue(j + AHEAD);
    const int s = j % TL_NSTAGE;
    const int t = t0 + j;
    const int tyi = t / ntx, txi = t - tyi * ntx;
    const int x0 = txi * TL_W, y0 = tyi * TL_H;
    ecc_mbar_wait(full0 + 8 * s, ((par0 >> s) ^ (unsigned)(j / TL_NSTAGE)) & 1u);
    const int4 ma = S.tl_meta[s][0], mb4 = S.tl_meta[s][1];
    const unsigned char *stg = dynp + s * TL_STAGE_BYTES;
    const float *win = reinterpret_cast<const float *>(stg);
    const float *sref = reinterpret_cast<const float *>(stg + TL_WIN_SLOT) + tq * TL_W + tx;
    const int xi = x0 + tx;
    const float x = (float)min(xi, cols - 1);
    const float ymax = (float)(rows - 1);
    const float yb = (float)(y0 + tq);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yi = y0 + tq + k * RQ;
      const bool inb = xi < cols && yi < rows;
      const float y = fminf(yb + (float)(k * RQ), ymax);
      float u, v;
      map_xy_t<MT>(m, x, y, u, v);
      const int sx = cvround32(u), sy = cvround32(v);
      bool ok;
      if (lm_masks) ok = u >= -0.5f && u <= hx && v >= -0.5f && v <= hy;   // cvRound(u), cvRound(v) inside the image
      else ok = lin_valid(sx, sy, cols, rows);
      if (rmask || rho_mask) {
        const unsigned pi = (unsigned)(min(yi, rows - 1) * cols) + (unsigned)min(xi, cols - 1);
        if (rmask) ok = ok && rmask[pi] != 0;
      }
      ok = ok && inb;
      const int wx = (sx >> 5) - ma.x, wy = (sy >> 5) - ma.y;
      float g;
      if ((unsigned)(wx - ma.z) < (unsigned)ma.w && (unsigned)(wy - mb4.x) < (unsigned)mb4.y) {
        const float txf = frac32(sx), tyf = frac32(sy);
        const float wx0 = 1.0f - txf, wy0 = 1.0f - tyf;
        const float *p0 = win + wy * TL_WW + wx;
        const float s00 = p0[0], s01 = p0[1], s10 = p0[TL_WW], s11 = p0[TL_WW + 1];
        g = __fadd_rn(__fmul_rn(s00, __fmul_rn(wy0, wx0)), __fmul_rn(s01, __fmul_rn(wy0, txf)));
        g = __fadd_rn(g, __fmul_rn(s10, __fmul_rn(tyf, wx0)));
        g = __fadd_rn(g, __fmul_rn(s11, __fmul_rn(tyf, txf)));
      } else {
        g = lin_sample(cur, cols, rows, sx, sy);
      }
      const float fref = sref[k * RQ * TL_W];
      const float gxv = sref[TL_W * TL_H + k * RQ * TL_W], gyv = sref[2 * TL_W * TL_H + k * RQ * TL_W];
      const float rhs = g - fref;
      if (RHO) {
        bool okr = lin_valid(sx, sy, cols, rows) && inb;
        if (rho_mask) okr = okr && rho_mask[(unsigned)(min(yi, rows - 1) * cols) + (unsigned)min(xi, cols - 1)] != 0;
        const double gd = okr ? (double)g : 0.0, fd = okr ? (double)fref : 0.0;
        nrho += okr ? 1 : 0;
        acc[OR + 1] += fd; acc[OR + 2] += gd; acc[OR + 3] += fd * fd; acc[OR + 4] += gd * gd; acc[OR + 5] += fd * gd;
      }
      float J[M];
      eval_J<TYPE>(jc, x, y, gxv, gyv, J);
      const double r = ok ? (double)rhs : 0.0;
      acc[0] += r * r;
      nvalid += ok ? 1 : 0;
#pragma unroll
      for (int q = 0; q < M; ++q) acc[2 + q] += (double)J[q] * r;
    }
    __syncwarp();
    if (lane == 0) ecc_mbar_arrive(empty0 + 8 * s);
  }
#pragma unroll
  for (int s = 0; s < TL_NSTAGE; ++s) {
    const int uses = ntiles > s ? (ntiles - s + TL_NSTAGE - 1) / TL_NSTAGE : 0;
    c.tl_par ^= (unsigned)(uses & 1) << s;
    c.tl_used |= (unsigned)(uses > 0) << s;
  }
  acc[1] = (double)nvalid;
  if (RHO) acc[OR] = (double)nrho;
  cluster_reduce<NS>(c, acc);
  if (RHO) {
    if (c.tid < 6) c.S->rho_try[c.tid] = c.S->tot[OR + c.tid];
    __syncthreads();
  }
}

template <int TYPE, bool RHO>
__device__ __forceinline__ void pass_ic_any(Ctx &c, int lvl, bool lm_masks) {
  if constexpr (TYPE != SSK_MOTION_HOMOGRAPHY) {
    if (level_uses_tma(c, lvl)) { pass_ic_tma<TYPE, RHO>(c, lvl, lm_masks); return; }
  }
  pass_ic<TYPE, RHO>(c, lvl, lm_masks);
}

// reference-side normal matrix Hp = J^T J (ecc_compute_hessian_matrix, ecc2.cc:295-322): sums = lower triangle
template <int TYPE>
__device__ void pass_hp(Ctx &c, int lvl) {
  constexpr int M = NParams<TYPE>::M;
  constexpr int NS = M * (M + 1) / 2;
  Shared &S = *c.S;
  const EccLevel &L = c.cfg->lv[lvl];
  const JCoef jc = S.jc;
  double acc[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) acc[k] = 0.0;
  const int n = L.cols * L.rows;
  const Band bd = pass_band(c.rank, c.csize, c.tid, n);
  for (Walk w(bd.start, bd.stride, L.cols); w.i < bd.end; w.next()) {
    const int i = w.i;
    float J[M];
    eval_J<TYPE>(jc, (float)w.x, (float)w.y, __ldg(L.gx + i), __ldg(L.gy + i), J);
    int q = 0;
#pragma unroll
    for (int a = 0; a < M; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) { acc[q] += (double)J[a] * (double)J[b]; ++q; }
  }
  cluster_reduce<NS>(c, acc);
}

// thread 0: unpack a lower triangle of sums into a symmetric float matrix (H stored as cv::Mat1f)
__device__ void unpack_H(int M, const double *tri, float *H) {
  int q = 0;
  for (int a = 0; a < M; ++a)
    for (int b = 0; b <= a; ++b) { H[a * M + b] = (float)tri[q]; H[b * M + a] = (float)tri[q]; ++q; }
}

// ------------------------------------------------------------------------------------------------
// passes of the forward methods (the warped-image gradient needs a stencil -> tiles with a 2-px halo in smem)
//   FA (c_ecc_forward_additive, ecc2.cc:1297-1347) runs two passes per iteration so that the residual is formed
//   exactly as the reference forms it:
//     pass_fa_stats : [ n, Sf, Sf2, Sg, Sg2 ] over wmask -> r = sg/sf, c = gm - r fm          (cv::meanStdDev)
//     pass_forward  : rhs = fma(f, (float)-r, g) - (float)c  (cv::scaleAdd, cv::subtract with a Scalar);
//                     sums = [ H (lower triangle), ep_i = J_i.rhs ]
//   LM (c_ecclm::compute_jac, ecc2.cc:1482-1525): sums = [ |rhs|^2, #valid, v_i = J_i.rhs, H (lower triangle) ]
//   All sums are accumulated in double like cv::Mat::dot / cv::norm.
// ------------------------------------------------------------------------------------------------
template <int DUMMY>
__device__ __forceinline__ float fa_sample(const Img &cur, int interp, float u, float v) {
  if (interp == SSK_INTER_NEAREST) return sample_nearest<SSK_32F>(cur, 0, u, v, SSK_BORDER_REPLICATE, 0.f);
  return lin_sample(static_cast<const float *>(cur.data), cur.cols, cur.rows, cvround32(u), cvround32(v));
}

__device__ void pass_fa_stats(Ctx &c, int lvl) {
  Shared &S = *c.S;
  const EccLevel &L = c.cfg->lv[lvl];
  const Img cur = level_image(c, lvl);
  const MapCoef m = S.map;
  const int interp = c.cfg->interp;
  double acc[5] = {0, 0, 0, 0, 0};
  const int n = L.cols * L.rows;
  const Band bd = pass_band(c.rank, c.csize, c.tid, n);
  const uint8_t *__restrict__ cm = c.frame->cmask ? c.frame->cmask + L.cur_off : nullptr;
  for (Walk w(bd.start, bd.stride, L.cols); w.i < bd.end; w.next()) {
    const int i = w.i;
    float u, v;
    map_xy(m, (float)w.x, (float)w.y, u, v);
    // wmask = remap(current_mask, INTER_LINEAR, BORDER_CONSTANT 0) >= 255 (ecc2.cc:1311-1312)
    bool ok = cm ? mask_lin_value(cm, L.cols, L.rows, cvround32(u), cvround32(v)) >= 255 : valid255_linear(u, v, L.cols, L.rows);
    if (ok && L.refmask) ok = L.refmask[i] != 0;
    if (!ok) continue;
    const double g = fa_sample<0>(cur, interp, u, v);
    const double f = __ldg(L.ref + i);
    acc[0] += 1.0; acc[1] += f; acc[2] += f * f; acc[3] += g; acc[4] += g * g;
  }
  cluster_reduce<5>(c, acc);
}

template <int TYPE, bool FA>
__device__ void pass_forward(Ctx &c, int lvl) {
  constexpr int M = NParams<TYPE>::M;
  constexpr int NH = M * (M + 1) / 2;
  constexpr int NS = FA ? NH + M : 2 + M + NH;
  constexpr int OH = FA ? 0 : 2 + M;     // offset of the H triangle
  constexpr int OV = FA ? NH : 2;        // offset of J.rhs
  Shared &S = *c.S;
  const EccLevel &L = c.cfg->lv[lvl];
  const Img cur = level_image(c, lvl);
  const MapCoef m = S.map;
  const JCoef jc = S.jc;
  const int interp = FA ? c.cfg->interp : SSK_INTER_LINEAR;
  const float fa_a = S.fa_a, fa_c = S.fa_c;
  // forward-additive: remapped mask >= 255 (ecc2.cc:1311-1312); LM: ecc_remap's >= 250 (ecc2.cc:205-216, 1444-1474)
  const uint8_t *__restrict__ cm = c.frame->cmask ? c.frame->cmask + L.cur_off : nullptr;
  const int cm_thr = FA ? 255 : 250;
  double acc[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) acc[k] = 0.0;
  const int ntx = (L.cols + FT_W - 1) / FT_W, nty = (L.rows + FT_H - 1) / FT_H;
  const int tx = c.tid & 31, ty = c.tid >> 5;
  for (int t = c.rank; t < ntx * nty; t += c.csize) {
    const int x0 = (t % ntx) * FT_W, y0 = (t / ntx) * FT_H;
    for (int k = c.tid; k < FT_IH * FT_IW; k += NT) {
      const int r = k / FT_IW, cc = k - r * FT_IW;
      const int gx_ = min(max(x0 - 2 + cc, 0), L.cols - 1), gy_ = min(max(y0 - 2 + r, 0), L.rows - 1);
      float u, v;
      map_xy(m, (float)gx_, (float)gy_, u, v);
      S.gw[r][cc] = fa_sample<0>(cur, interp, u, v);
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    if (x < L.cols && y < L.rows) {
      const int i = y * L.cols + x;
      float u, v;
      map_xy(m, (float)x, (float)y, u, v);
      bool ok = cm ? mask_lin_value(cm, L.cols, L.rows, cvround32(u), cvround32(v)) >= cm_thr : valid255_linear(u, v, L.cols, L.rows);
      if (ok && L.refmask) ok = L.refmask[i] != 0;
      if (ok) {
        const int r = ty + 2, cc = tx + 2;
        // ecc_differentiate (ecc2.cc:142-169) with OpenCV's filter-engine arithmetic:
        // gx = sepFilter2D(gw, d5 along x, s3 along y); gy = sepFilter2D(gw, s3 along x, d5 along y)
        const float k1 = 2.f / 3.f, k2 = -1.f / 12.f;
        float rd[3], rs[5];
#pragma unroll
        for (int d = -1; d <= 1; ++d)
          rd[d + 1] = __fmaf_rn(__fsub_rn(S.gw[r + d][cc + 2], S.gw[r + d][cc - 2]), k2,
                                __fmul_rn(__fsub_rn(S.gw[r + d][cc + 1], S.gw[r + d][cc - 1]), k1));
#pragma unroll
        for (int d = -2; d <= 2; ++d)
          rs[d + 2] = __fmaf_rn(__fadd_rn(S.gw[r + d][cc + 1], S.gw[r + d][cc - 1]), 0.25f, __fmul_rn(0.5f, S.gw[r + d][cc]));
        const float gxw = __fmaf_rn(__fadd_rn(rd[2], rd[0]), 0.25f, __fmul_rn(0.5f, rd[1]));
        const float gyw = __fmaf_rn(__fsub_rn(rs[4], rs[0]), k2, __fmul_rn(__fsub_rn(rs[3], rs[1]), k1));
        const float g = S.gw[r][cc], f = __ldg(L.ref + i);
        float J[M];
        eval_J<TYPE>(jc, (float)x, (float)y, gxw, gyw, J);
        float rhs;
        if (FA) {
          rhs = __fsub_rn(__fmaf_rn(f, fa_a, g), fa_c);
        } else {
          rhs = __fsub_rn(g, f);
          acc[0] += (double)rhs * (double)rhs;
          acc[1] += 1.0;
        }
#pragma unroll
        for (int k = 0; k < M; ++k) acc[OV + k] += (double)J[k] * (double)rhs;
        int q = OH;
#pragma unroll
        for (int a = 0; a < M; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) { acc[q] += (double)J[a] * (double)J[b]; ++q; }
      }
    }
    __syncthreads();
  }
  cluster_reduce<NS>(c, acc);
}

// compute_correlation (ecc2.cc:65-137): sums = [ n, Sf, Sg, Sf2, Sg2, Sfg ] over (remap(255) >= 254) & refmask,
// g = remap(current, INTER_LINEAR, BORDER_CONSTANT 0)
template <int MT>
__device__ void pass_rho(Ctx &c) {
  Shared &S = *c.S;
  const EccLevel &L = c.cfg->lv[0];
  const float *__restrict__ cur = opaque_ptr(c.frame->pyr + L.cur_off);
  const float *__restrict__ ref = L.ref;
  const uint8_t *__restrict__ rmask = L.refmask;
  const uint8_t *__restrict__ cm = c.frame->cmask;     // level 0 of the solver's current mask (c_ecch::current_mask())
  const int cols = L.cols, rows = L.rows;
  const MapCoef m = S.map;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  int nvalid = 0;
  const int n = cols * rows;
  const Band bd = pass_band(c.rank, c.csize, c.tid, n);
  WalkF w(bd.start, bd.stride, cols);
#pragma unroll 2
  for (; w.i < bd.end; w.next()) {
    float u, v;
    map_xy_t<MT>(m, w.fx, w.fy, u, v);
    const int sx = cvround32(u), sy = cvround32(v);
    bool ok = cm ? mask_lin_value(cm, cols, rows, sx, sy) >= 254 : lin_valid(sx, sy, cols, rows);
    if (rmask) ok = ok && rmask[w.i] != 0;
    // valid pixels have every non-zero-weight tap in bounds: the clamped sample equals BORDER_CONSTANT 0
    const double g = ok ? (double)lin_sample(cur, cols, rows, sx, sy) : 0.0;
    const double f = ok ? (double)__ldg(ref + w.i) : 0.0;
    nvalid += ok ? 1 : 0;
    acc[1] += f; acc[2] += g; acc[3] += f * f; acc[4] += g * g; acc[5] += f * g;
  }
  acc[0] = (double)nvalid;
  cluster_reduce<6>(c, acc);
}

__device__ double rho_from_sums(const double *t) {
  const double n = t[0];
  if (!(n > 0)) return 0.0;
  const double m1 = t[1] / n, m2 = t[2] / n;
  const double v1 = fmax(t[3] / n - m1 * m1, 0.0), v2 = fmax(t[4] / n - m2 * m2, 0.0);
  const double covar = t[5] / n - m1 * m2;
  return covar / (sqrt(v1) * sqrt(v2));
}

// ------------------------------------------------------------------------------------------------
// level solvers.  All threads run the control flow; scalar work is done by thread 0 of each CTA on the
// CTA's own copy of the state, which stays identical across the cluster.
// ------------------------------------------------------------------------------------------------
#define T0_BEGIN __syncthreads(); if (c.tid == 0) {
#define T0_END } __syncthreads();

// reference-side Hp and steepest-descent coefficients of this level for the current transform
template <int TYPE>
__device__ void prepare_ic_level(Ctx &c, int lvl, bool main_pass) {
  constexpr int M = NParams<TYPE>::M;
  Shared &S = *c.S;
  const EccConfig &cfg = *c.cfg;
  const EccHpCache *cache = main_pass ? cfg.hp_main : cfg.hp_trans;
  const int mode = main_pass ? cfg.hp_main_mode : 0;
  if (mode == 0) {
    T0_BEGIN
    for (int i = 0; i < M * M; ++i) S.Hp[i] = cache->Hp[lvl][i];
    ssk_transform jt = S.t;
    for (int i = 0; i < 8; ++i) jt.params[i] = cache->jp[lvl][i];
    for (int i = 0; i < 4; ++i) jt.aux[i] = cache->jp[lvl][8 + i];
    make_jcoef(jt, S.jc);
    trace_hp(c, lvl, M);
    T0_END
  } else {
    // jac is (re)built with the parameters the transform holds at this moment (ecc2.cc:1733-1738, 1985-1991)
    T0_BEGIN
    make_jcoef(S.t, S.jc);
    T0_END
    pass_hp<TYPE>(c, lvl);
    T0_BEGIN
    unpack_H(M, S.tot, S.Hp);
    if (mode == 2 && c.rank == 0) {
      EccHpCache *w = cfg.hp_main;
      for (int i = 0; i < M * M; ++i) w->Hp[lvl][i] = S.Hp[i];
      for (int i = 0; i < 8; ++i) w->jp[lvl][i] = S.t.params[i];
      for (int i = 0; i < 4; ++i) w->jp[lvl][8 + i] = S.t.aux[i];
      w->valid[lvl] = 1;
    }
    T0_END
  }
}

// c_ecclm_inverse_compositional::align (ecc2.cc:1926-2086)
template <int TYPE>
__device__ bool align_iclm(Ctx &c, int lvl, double max_eps, bool main_pass) {
  constexpr int M = NParams<TYPE>::M;
  Shared &S = *c.S;
  const EccConfig &cfg = *c.cfg;
  const EccLevel &L = cfg.lv[lvl];
  prepare_ic_level<TYPE>(c, lvl, main_pass);
  // the correlation gate that follows the alignment (c_frame_registration.cc:838, 861) reads level 0 at the final parameters
  const bool with_rho = lvl == 0 && (main_pass ? cfg.check_rho != 0 : true) && !getenv_no_rho_fusion();
  T0_BEGIN
  S.lambda = 0.001; S.dp = 0; S.recompute = 1; S.num_it = 0; S.eps = FLT_MAX; S.err = 0; S.newerr = 0;
  if (lvl == 0) S.rho_have = 0;
  T0_END
  while (S.num_it < cfg.max_iterations) {
    if (S.recompute) {
      T0_BEGIN set_pass_params(S, S.t); T0_END
      const bool rho_init = with_rho && SSK_ECC_RHO_FUSION == 1;
      if (rho_init) pass_ic_any<TYPE, true>(c, lvl, true); else pass_ic_any<TYPE, false>(c, lvl, true);
      T0_BEGIN
      const double CMA = S.tot[1], RMA = L.RMA;
      S.err = S.tot[0] * (RMA * RMA) / (CMA * CMA);
      for (int i = 0; i < M; ++i) S.v[i] = __fmul_rn((float)S.tot[2 + i], (float)(RMA / CMA));
      if (rho_init) { for (int i = 0; i < 6; ++i) S.rho_acc[i] = S.rho_try[i]; S.rho_have = 1; }
      T0_END
    }
    do {
      T0_BEGIN
      ++S.num_it;
      S.recompute = 1;
      float H[64], sdp[8], np[8];
      for (int i = 0; i < M * M; ++i) H[i] = S.Hp[i];
      for (int i = 0; i < M; ++i) H[i * M + i] = (float)((1 + S.lambda) * (double)S.Hp[i * M + i]);
      cv_solve(M, H, S.v, S.deltap);
      for (int i = 0; i < M; ++i) sdp[i] = (float)(cfg.update_step_scale * (double)S.deltap[i]);
      xf_invert_and_compose(S.t, sdp, np);
      ssk_transform q = S.t;
      for (int i = 0; i < M; ++i) q.params[i] = np[i];
      set_pass_params(S, q);
      T0_END
      if (with_rho) pass_ic_any<TYPE, true>(c, lvl, true); else pass_ic_any<TYPE, false>(c, lvl, true);
      T0_BEGIN
      const double CMA = S.tot[1], RMA = L.RMA;
      S.newerr = S.tot[0] * (RMA * RMA) / (CMA * CMA);
      for (int i = 0; i < M; ++i) S.vtrial[i] = __fmul_rn((float)S.tot[2 + i], (float)(RMA / CMA));
      S.dp = xf_eps(TYPE, S.deltap, L.cols, L.rows, S.t.aux[3]);
      S.brk = 0;
      if (S.dp < max_eps) {
        S.brk = 1;
      } else {
        // temp_d = -Hp * deltap + 2 v ; dS = deltap . temp_d (cv::gemm on float data, dot in double)
        double dS = 0;
        for (int i = 0; i < M; ++i) {
          double hd = 0;
          for (int k = 0; k < M; ++k) hd += (double)S.Hp[i * M + k] * (double)S.deltap[k];
          const float td = (float)(-hd + 2.0 * (double)S.v[i]);   // cv::gemm(Hp, deltap, -1, v, 2)
          dS += (double)S.deltap[i] * (double)td;
        }
        const double rho = (S.err - S.newerr) / (fabs(dS) > DBL_EPSILON ? dS : 1);
        if (rho > 0.25) { if (S.lambda > 1e-6) S.lambda = fmax(1e-6, S.lambda / 5); }
        else if (rho > 0.1) { }
        else if (S.lambda < 1) S.lambda = 1;
        else S.lambda *= 10;
        if (S.newerr < S.err) S.brk = 1;
      }
      S.eps = S.dp;
      trace_rec(c, lvl, 3);
      T0_END
    } while (!S.brk && S.num_it < cfg.max_iterations);
    T0_BEGIN
    if (S.newerr < S.err) {
      S.err = S.newerr;
      S.recompute = 0;
      for (int i = 0; i < M; ++i) { S.t.params[i] = S.tq.params[i]; S.v[i] = S.vtrial[i]; }
      if (with_rho) { for (int i = 0; i < 6; ++i) S.rho_acc[i] = S.rho_try[i]; S.rho_have = 1; }   // sums of the pass that evaluated tq
    }
    T0_END
    if (S.dp < max_eps) break;
  }
  T0_BEGIN S.eps = S.dp; S.total_iterations += S.num_it; T0_END
  return true;
}

// c_ecc_inverse_compositional::align (ecc2.cc:1693-1788)
template <int TYPE>
__device__ bool align_ic(Ctx &c, int lvl, double max_eps, bool main_pass) {
  constexpr int M = NParams<TYPE>::M;
  Shared &S = *c.S;
  const EccConfig &cfg = *c.cfg;
  const EccLevel &L = cfg.lv[lvl];
  prepare_ic_level<TYPE>(c, lvl, main_pass);
  T0_BEGIN S.num_it = 0; S.eps = FLT_MAX; S.rmsold = FLT_MAX; S.brk = 0; T0_END
  while (true) {
    T0_BEGIN
    S.cont = S.num_it < cfg.max_iterations;
    ++S.num_it;
    if (S.cont) set_pass_params(S, S.t);
    T0_END
    if (!S.cont) break;
    pass_ic_any<TYPE, false>(c, lvl, false);
    T0_BEGIN
    const double CMA = S.tot[1], RMA = L.RMA;
    const double rmsnew = S.tot[0] * (RMA * RMA) / (CMA * CMA);
    float vs[8];
    for (int i = 0; i < M; ++i) vs[i] = __fmul_rn((float)S.tot[2 + i], (float)(RMA / CMA));
    cv_solve(M, S.Hp, vs, S.deltap);
    S.brk = 0;
    if (rmsnew >= S.rmsold) {
      S.brk = 1;
    } else {
      float sdp[8], np[8];
      for (int i = 0; i < M; ++i) sdp[i] = (float)cfg.update_step_scale * S.deltap[i];
      xf_invert_and_compose(S.t, sdp, np);
      S.rmsold = rmsnew;
      for (int i = 0; i < M; ++i) S.t.params[i] = np[i];
      S.eps = xf_eps(TYPE, S.deltap, L.cols, L.rows, S.t.aux[3]);
      if (S.eps < max_eps) S.brk = 1;
    }
    S.err = rmsnew;
    trace_rec(c, lvl, 1);
    T0_END
    if (S.brk) break;
  }
  T0_BEGIN S.total_iterations += S.num_it; T0_END
  return true;
}

// c_ecc_forward_additive::align (ecc2.cc:1247-1365)
template <int TYPE>
__device__ bool align_fa(Ctx &c, int lvl, double max_eps) {
  constexpr int M = NParams<TYPE>::M;
  constexpr int NH = M * (M + 1) / 2;
  Shared &S = *c.S;
  const EccConfig &cfg = *c.cfg;
  const EccLevel &L = cfg.lv[lvl];
  if (max_eps <= 0) max_eps = 1e-3;
  T0_BEGIN S.num_it = 0; S.failed = 0; S.brk = 0; T0_END
  while (true) {
    T0_BEGIN
    S.cont = S.num_it < cfg.max_iterations;
    ++S.num_it;
    if (S.cont) { set_pass_params(S, S.t); make_jcoef(S.t, S.jc); }
    T0_END
    if (!S.cont) break;
    pass_fa_stats(c, lvl);
    T0_BEGIN
    {
      // cv::meanStdDev over wmask (ecc2.cc:1324-1326)
      const double *t = S.tot;
      const double n = t[0];
      const double fMean = t[1] / n, gMean = t[3] / n;
      const double fStd = sqrt(fmax(t[2] / n - fMean * fMean, 0.0)), gStd = sqrt(fmax(t[4] / n - gMean * gMean, 0.0));
      const double r = gStd / fStd;
      S.err = r;
      S.newerr = n;
      S.fa_a = (float)(-r);                      // cv::scaleAdd(f, -r, gw)
      S.fa_c = (float)(gMean - r * fMean);        // cv::subtract(rhs, Scalar)
    }
    T0_END
    pass_forward<TYPE, true>(c, lvl);
    T0_BEGIN
    const double *t = S.tot;
    const double n = S.newerr, r = S.err;
    float H[64], Hi[64], ep[8];
    unpack_H(M, t, H);
    for (int i = 0; i < M; ++i) ep[i] = (float)t[NH + i];
    S.brk = 0;
    if (!(n > 0) || !cv_invert(M, H, Hi)) {
      S.failed = 1;
      S.brk = 1;
    } else {
      // dp = -update_step_scale * (Hinv * ep): cv::gemm with alpha (double accumulation, one rounding)
      float dpv[8];
      for (int i = 0; i < M; ++i) {
        double acc = 0;
        for (int k = 0; k < M; ++k) acc += (double)Hi[i * M + k] * (double)ep[k];
        dpv[i] = (float)(acc * -cfg.update_step_scale);
        S.t.params[i] = S.t.params[i] + dpv[i];
      }
      S.eps = xf_eps(TYPE, dpv, L.cols, L.rows, S.t.aux[3]);
      if (S.eps < max_eps) S.brk = 1;
      for (int i = 0; i < M; ++i) { S.deltap[i] = dpv[i]; S.v[i] = ep[i]; }
    }
    S.err = r;
    trace_rec(c, lvl, 0);
    T0_END
    if (S.brk) break;
  }
  T0_BEGIN S.total_iterations += S.num_it; T0_END
  return !S.failed;
}

// c_ecclm::align (ecc2.cc:1528-1650)
template <int TYPE>
__device__ bool align_lm(Ctx &c, int lvl, double max_eps) {
  constexpr int M = NParams<TYPE>::M;
  Shared &S = *c.S;
  const EccConfig &cfg = *c.cfg;
  const EccLevel &L = cfg.lv[lvl];
  T0_BEGIN S.lambda = 0.1; S.num_it = 0; S.converged = 0; S.recompute = 1; T0_END
  while (S.num_it < cfg.max_iterations) {
    if (S.recompute) {
      T0_BEGIN set_pass_params(S, S.t); make_jcoef(S.t, S.jc); T0_END
      pass_forward<TYPE, false>(c, lvl);
      T0_BEGIN
      S.err = S.tot[0];
      for (int i = 0; i < M; ++i) S.v[i] = (float)S.tot[2 + i];
      unpack_H(M, S.tot + 2 + M, S.Hp);
      T0_END
    } else {
      // compute_jac(params, recompute_remap = false): reuse the accepted trial's image, rhs, H and v
      T0_BEGIN
      S.err = S.newerr;
      for (int i = 0; i < M; ++i) S.v[i] = S.vtrial[i];
      for (int i = 0; i < M * M; ++i) S.Hp[i] = S.Htrial[i];
      T0_END
    }
    if (S.err < 1) { T0_BEGIN S.converged = 1; T0_END break; }
    while (true) {
      T0_BEGIN
      S.cont = S.num_it < cfg.max_iterations;
      ++S.num_it;
      S.brk = 0;
      if (S.cont) {
        S.recompute = 1;
        float H[64];
        for (int i = 0; i < M * M; ++i) H[i] = S.Hp[i];
        for (int i = 0; i < M; ++i) H[i * M + i] = (float)((1 + S.lambda) * (double)S.Hp[i * M + i]);
        cv_solve(M, H, S.v, S.deltap);
        ssk_transform q = S.t;
        for (int i = 0; i < M; ++i) q.params[i] = __fadd_rn(__fmul_rn(S.deltap[i], (float)(-cfg.update_step_scale)), S.t.params[i]);   // cv::scaleAdd
        S.eps = xf_eps(TYPE, S.deltap, L.cols, L.rows, S.t.aux[3]);
        if (S.eps <= max_eps) {
          for (int i = 0; i < M; ++i) S.t.params[i] = q.params[i];
          S.converged = 1;
          S.brk = 1;
        } else {
          set_pass_params(S, q);
          make_jcoef(q, S.jc);
        }
      }
      T0_END
      if (!S.cont || S.brk) break;
      pass_forward<TYPE, false>(c, lvl);
      T0_BEGIN
      S.newerr = S.tot[0];
      S.brk = 0;
      if (S.newerr > S.err) {
        if (S.lambda > 1e6) S.brk = 1;
        else S.lambda *= 10.0;
      } else {
        for (int i = 0; i < M; ++i) { S.t.params[i] = S.tq.params[i]; S.vtrial[i] = (float)S.tot[2 + i]; }
        unpack_H(M, S.tot + 2 + M, S.Htrial);
        S.recompute = 0;
        const double diff = S.err - S.newerr;
        if (diff < S.err * cfg.max_epse) {
          S.converged = 1;
        } else {
          double dS = 0;
          for (int i = 0; i < M; ++i) {
            float td = 0.f;
            for (int k = 0; k < M; ++k) td += S.Hp[i * M + k] * S.deltap[k];
            td = -td + 2.f * S.v[i];
            dS += (double)S.deltap[i] * (double)td;
          }
          const double rho = fabs(dS) > (double)1e-9f ? diff / fabs(dS) : diff;
          if (rho > 0.25) S.lambda = fmax(1e-8, 0.2 * S.lambda);
          else if (rho < 0.1) S.lambda = S.lambda < 1.0 ? 1.0 : S.lambda * 10.0;
        }
        S.brk = 1;
      }
      T0_END
      if (S.brk) break;
    }
    if (S.converged) break;
  }
  T0_BEGIN S.total_iterations += S.num_it; T0_END
  return S.converged != 0;
}

template <int METHOD, int TYPE>
__device__ bool align_level(Ctx &c, int lvl, double max_eps, bool main_pass) {
  if (METHOD == SSK_ECC_FORWARD_ADDITIVE) return align_fa<TYPE>(c, lvl, max_eps);
  if (METHOD == SSK_ECC_INVERSE_COMPOSITIONAL) return align_ic<TYPE>(c, lvl, max_eps, main_pass);
  if (METHOD == SSK_ECC_INVERSE_COMPOSITIONAL_LM) return align_iclm<TYPE>(c, lvl, max_eps, main_pass);
  return align_lm<TYPE>(c, lvl, max_eps);
}

// c_ecch::align (ecc2.cc:1133-1176) for the transform currently held in S.t
template <int METHOD, int TYPE>
__device__ void ecch_align(Ctx &c, bool main_pass) {
  Shared &S = *c.S;
  const EccConfig &cfg = *c.cfg;
  int lvl = cfg.nlevels - 1;
  T0_BEGIN
  S.total_iterations = 0;
  if (lvl > 0) xf_scale(S.t, (double)cfg.lv[lvl].cols / (double)cfg.lv[0].cols);
  T0_END
  for (; lvl >= 0; --lvl) {
    double max_eps = cfg.epsx;
    for (int i = 0; i < lvl; ++i) max_eps *= 2;
    const bool ok = align_level<METHOD, TYPE>(c, lvl, max_eps, main_pass);
    T0_BEGIN
    if (ok && lvl > 0) xf_scale(S.t, (double)cfg.lv[lvl - 1].cols / (double)cfg.lv[lvl].cols);
    T0_END
  }
}

template <int MT>
__device__ double correlation(Ctx &c) {
  Shared &S = *c.S;
  if (S.rho_have) {       // level 0 was last evaluated at S.t by a pass that gathered these sums (align_iclm)
    T0_BEGIN S.rho = rho_from_sums(S.rho_acc); S.rho_have = 0; T0_END
    return S.rho;
  }
  T0_BEGIN set_pass_params(S, S.t); T0_END
  pass_rho<MT>(c);
  T0_BEGIN S.rho = rho_from_sums(S.tot); T0_END
  return S.rho;
}

template <int METHOD, int TYPE>
__global__ void __launch_bounds__(NT, SSK_ECC_MINB) k_ecc(const __grid_constant__ EccConfig cfg, EccFrame *frames) {
  __shared__ Shared S;
  cg::cluster_group cluster = cg::this_cluster();
  Ctx c;
  c.cfg = &cfg;
  c.csize = (int)cluster.num_blocks();
  c.rank = (int)cluster.block_rank();
  c.tid = threadIdx.x;
  c.buf = 0;
  c.tl_par = 0; c.tl_used = 0;
  c.S = &S;
  EccFrame *fr = frames + blockIdx.x / c.csize;
  c.frame = fr;
  if (c.tid == 0) {     // stage barriers of the TMA-staged passes (the T0 section below orders this before any use)
    for (int s = 0; s < TL_NSTAGE; ++s) {
      ecc_mbar_init((unsigned)__cvta_generic_to_shared(&S.tl_full[s]), 1);
      ecc_mbar_init((unsigned)__cvta_generic_to_shared(&S.tl_empty[s]), NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  T0_BEGIN
  S.t = fr->t;
  S.failed = 0; S.rho = -1; S.eps = FLT_MAX; S.total_iterations = 0; S.rho_have = 0;
  for (int i = 0; i < 8; ++i) S.jc.c[i] = 0.f;
  T0_END

  bool ok = true;
  // c_frame_registration.cc:815-850: translation-only estimate first
  if (TYPE != SSK_MOTION_TRANSLATION && cfg.translation_first) {
    T0_BEGIN
    float tx, ty;
    xf_get_translation(S.t, tx, ty);
    ssk_transform tt;
    tt.motion_type = SSK_MOTION_TRANSLATION; tt.nparams = 2;
    for (int i = 0; i < 8; ++i) tt.params[i] = 0.f;
    for (int i = 0; i < 4; ++i) tt.aux[i] = 0.f;
    tt.params[0] = tx; tt.params[1] = ty;
    S.tmain = S.t;
    S.t = tt;
    T0_END
    ecch_align<METHOD, SSK_MOTION_TRANSLATION>(c, false);
    const double rho = correlation<MAP_TRANSLATION>(c);
    if (rho < 0.75 * cfg.min_rho) ok = false;
    T0_BEGIN
    const float tx = S.t.params[0], ty = S.t.params[1];
    S.t = S.tmain;
    if (ok) xf_set_translation(S.t, tx, ty);
    T0_END
  }
  if (ok) {
    ecch_align<METHOD, TYPE>(c, true);
    if (cfg.check_rho) {
      const double rho = correlation<MapKind<TYPE>::MT>(c);
      if (rho < cfg.min_rho) ok = false;
    }
  }
  if (c.tid == 0 && c.rank == 0) {
    ssk_transform t = S.t;
    if (ok && cfg.final_scale != 1.0) xf_scale(t, cfg.final_scale);
    fr->t = t;
    fr->map = make_mapcoef(t);
    fr->rho = S.rho;
    fr->eps = S.eps;
    fr->num_iterations = S.total_iterations;
    fr->ok = ok ? 1 : 0;
    fr->failed = S.failed;
  }
  cluster.sync();   // no CTA may exit while a peer can still read its shared memory
}

inline int launch_clustered(const void *kernel, void **args, int nclusters, int cluster_size, cudaStream_t s, unsigned dyn_smem = 0) {
  if (dyn_smem) SSK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(nclusters * cluster_size);
  lc.blockDim = dim3(NT);
  lc.dynamicSmemBytes = dyn_smem;
  lc.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  SSK_CUDA(cudaLaunchKernelExC(&lc, kernel, args));
  count_launch();
  return SSK_OK;
}

template <int METHOD>
int launch_ecc_method(const EccConfig &cfg, EccFrame *frames, int nframes, int cluster_size, cudaStream_t s) {
  void *args[2] = {(void *)&cfg, (void *)&frames};
  const void *k;
  switch (cfg.motion_type) {
    case SSK_MOTION_TRANSLATION: k = (const void *)k_ecc<METHOD, SSK_MOTION_TRANSLATION>; break;
    case SSK_MOTION_EUCLIDEAN: k = (const void *)k_ecc<METHOD, SSK_MOTION_EUCLIDEAN>; break;
    case SSK_MOTION_SCALED_EUCLIDEAN: k = (const void *)k_ecc<METHOD, SSK_MOTION_SCALED_EUCLIDEAN>; break;
    case SSK_MOTION_AFFINE: k = (const void *)k_ecc<METHOD, SSK_MOTION_AFFINE>; break;
    case SSK_MOTION_HOMOGRAPHY: k = (const void *)k_ecc<METHOD, SSK_MOTION_HOMOGRAPHY>; break;
    default: set_error("ECC: unsupported motion type"); return SSK_ERR_INVALID;
  }
  const bool ic = METHOD == SSK_ECC_INVERSE_COMPOSITIONAL || METHOD == SSK_ECC_INVERSE_COMPOSITIONAL_LM;
  const unsigned dyn = ic && cfg.tma_levels > 0 && cfg.motion_type != SSK_MOTION_HOMOGRAPHY ? TL_DYN_SMEM : 0u;
  return launch_clustered(k, args, nframes, cluster_size, s, dyn);
}

}  // namespace

}  // namespace ssk

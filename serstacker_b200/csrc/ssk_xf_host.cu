// Host side of c_image_transform's parameter algebra: eps(dp, size) and invert_and_compose(p, dp) for the five motion types
// of the path (core/proc/image_registration/c_image_transform.cc:136-139, 509-523, 736-833, 917-924, 1196-1205;
// c_image_transform.h:164-167, 312-318, 379-384).  The solvers run the same algebra on the device inside the persistent ECC
// kernel (ssk_ecc_impl.cuh: xf_eps, xf_invert_and_compose - thread 0 of every CTA); these entry points give a host that still
// drives its own Gauss-Newton loop, or inspects a step, the reference's numbers without a device round trip.  Plain C++ (the
// host pass is compiled with -ffp-contract=off, so float products and sums round as written), checked bit for bit against
// oracle/transforms.py in tests/test_transform_algebra.py.
#include <cmath>
#include <algorithm>
#include "ssk_common.cuh"

namespace ssk {
namespace {

inline float sqf(float x) { return x * x; }

// cv::invertAffineTransform on CV_32F data as cv2 4.13 computes it (oracle/cvmodel.py::invert_affine_f32): determinant, its
// reciprocal and the 2x2 part in float; the translation from double products of the rounded 2x2 entries
void invert_affine(const float *M, float *iM) {
  const float D = M[0] * M[4] - M[1] * M[3];
  const float Di = D != 0.f ? 1.0f / D : 0.f;
  const float A11 = M[4] * Di, A22 = M[0] * Di, A12 = -M[1] * Di, A21 = -M[3] * Di;
  iM[0] = A11; iM[1] = A12; iM[2] = (float)(-((double)A11 * (double)M[2] + (double)A12 * (double)M[5]));
  iM[3] = A21; iM[4] = A22; iM[5] = (float)(-((double)A21 * (double)M[2] + (double)A22 * (double)M[5]));
}

// cv::invert of a 3x3 CV_32F matrix: cofactors and determinant in double, result float
void invert3x3(const float *a, float *b) {
  const double a00 = a[0], a01 = a[1], a02 = a[2], a10 = a[3], a11 = a[4], a12 = a[5], a20 = a[6], a21 = a[7], a22 = a[8];
  double d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
  d = d != 0 ? 1. / d : 0;
  b[0] = (float)((a11 * a22 - a12 * a21) * d); b[1] = (float)((a02 * a21 - a01 * a22) * d); b[2] = (float)((a01 * a12 - a02 * a11) * d);
  b[3] = (float)((a12 * a20 - a10 * a22) * d); b[4] = (float)((a00 * a22 - a02 * a20) * d); b[5] = (float)((a02 * a10 - a00 * a12) * d);
  b[6] = (float)((a10 * a21 - a11 * a20) * d); b[7] = (float)((a01 * a20 - a00 * a21) * d); b[8] = (float)((a00 * a11 - a01 * a10) * d);
}

// the 2x3 matrix of a (scaled) euclidean transform about the centre (cX, cY): lambda at c_image_transform.cc:768-782
void euclid_matrix(float tX, float tY, float ang, float scl, float cX, float cY, float *M) {
  const float sa = (float)std::sin((double)ang), ca = (float)std::cos((double)ang);
  const float sc = scl * ca, ss = scl * sa;
  M[0] = sc; M[1] = -ss; M[2] = (tX - sc * cX) + ss * cY;
  M[3] = ss; M[4] = sc;  M[5] = (tY - ss * cX) - sc * cY;
}

}  // namespace
}  // namespace ssk

using namespace ssk;

extern "C" {

int ssk_transform_eps(const ssk_transform *t, const float *dp, int ndp, int rows, int cols, double *eps) {
  SSK_REQUIRE(t && dp && eps, "ssk_transform_eps: null argument");
  SSK_REQUIRE(ndp == t->nparams, "ssk_transform_eps: dp must have the transform's parameter count");
  const int w = cols, h = rows;
  switch (t->motion_type) {
    case SSK_MOTION_TRANSLATION:
      *eps = std::sqrt((double)(dp[0] * dp[0] + dp[1] * dp[1]));
      return SSK_OK;
    case SSK_MOTION_EUCLIDEAN:
    case SSK_MOTION_SCALED_EUCLIDEAN: {
      // get_parameters(dp): with the scale fixed the "scale step" read back is the transform's own scale (c_image_transform.cc:343-368)
      const float da = dp[2], ds = t->motion_type == SSK_MOTION_SCALED_EUCLIDEAN ? dp[3] : t->aux[3];
      const float sa = (float)std::sin((double)da);
      *eps = (double)std::sqrt(sqf(dp[0]) + sqf(dp[1]) + sqf(w * sa) + sqf(h * sa) + sqf((float)std::max(w, h) * ds));
      return SSK_OK;
    }
    case SSK_MOTION_AFFINE:
      *eps = (double)std::sqrt(sqf(w * dp[0]) + sqf(h * dp[1]) + sqf(dp[2]) + sqf(w * dp[3]) + sqf(h * dp[4]) + sqf(dp[5]));
      return SSK_OK;
    case SSK_MOTION_HOMOGRAPHY:
      *eps = (double)std::sqrt(sqf(dp[2]) + sqf(dp[5]) + sqf(w * dp[0]) + sqf(h * dp[1]) + sqf(w * dp[3]) + sqf(h * dp[4]));
      return SSK_OK;
    default:
      set_error("unsupported motion type");
      return SSK_ERR_INVALID;
  }
}

int ssk_transform_invert_and_compose(const ssk_transform *t, const float *dp, int ndp, float *out) {
  SSK_REQUIRE(t && dp && out, "ssk_transform_invert_and_compose: null argument");
  SSK_REQUIRE(ndp == t->nparams, "ssk_transform_invert_and_compose: dp must have the transform's parameter count");
  switch (t->motion_type) {
    case SSK_MOTION_TRANSLATION:
      out[0] = t->params[0] - dp[0]; out[1] = t->params[1] - dp[1];
      return SSK_OK;
    case SSK_MOTION_AFFINE: {
      float a[6], s[6];
      invert_affine(t->params, a);
      for (int i = 0; i < 6; ++i) s[i] = a[i] + dp[i];
      invert_affine(s, out);
      return SSK_OK;
    }
    case SSK_MOTION_HOMOGRAPHY: {
      float m[9], inv1[9], s[9], aii[9];
      for (int i = 0; i < 8; ++i) m[i] = t->params[i];
      m[8] = t->aux[2];
      invert3x3(m, inv1);
      for (int i = 0; i < 8; ++i) s[i] = inv1[i] + dp[i];
      s[8] = inv1[8];
      invert3x3(s, aii);
      const float k = 1.0f / aii[8];
      for (int i = 0; i < 8; ++i) out[i] = aii[i] * k;
      return SSK_OK;
    }
    case SSK_MOTION_EUCLIDEAN:
    case SSK_MOTION_SCALED_EUCLIDEAN: {
      const bool fix_scale = t->motion_type == SSK_MOTION_EUCLIDEAN;
      const float Tx = t->params[0], Ty = t->params[1], angle = t->params[2];
      const float scale = fix_scale ? t->aux[3] : t->params[3];
      const float Cx = t->aux[0], Cy = t->aux[1];
      const float scale_dp = fix_scale ? 1.0f : 1.0f + dp[3];
      float Mp[6], Mdp[6], Mi[6];
      euclid_matrix(Tx, Ty, angle, scale, Cx, Cy, Mp);
      euclid_matrix(dp[0], dp[1], dp[2], scale_dp, Cx, Cy, Mdp);
      invert_affine(Mdp, Mi);
      // Matx33f product Mp * Mdp^-1, s += a(i,k) * b(k,j) in order; the third row of both is 0 0 1
      const float m00 = Mp[0] * Mi[0] + Mp[1] * Mi[3];
      const float m02 = (Mp[0] * Mi[2] + Mp[1] * Mi[5]) + Mp[2];
      const float m10 = Mp[3] * Mi[0] + Mp[4] * Mi[3];
      const float m12 = (Mp[3] * Mi[2] + Mp[4] * Mi[5]) + Mp[5];
      const float rs = fix_scale ? scale : std::sqrt(m00 * m00 + m10 * m10);
      const float ra = (float)std::atan2((double)m10, (double)m00);
      const float rca = (float)std::cos((double)ra), rsa = (float)std::sin((double)ra);
      out[0] = (m02 + (rs * rca) * Cx) - (rs * rsa) * Cy;
      out[1] = (m12 + (rs * rsa) * Cx) + (rs * rca) * Cy;
      out[2] = ra;
      if (!fix_scale) out[3] = rs;
      return SSK_OK;
    }
    default:
      set_error("unsupported motion type");
      return SSK_ERR_INVALID;
  }
}

// c_image_transform::remap(params, rpts, cpts): reference points -> current-frame points (c_image_transform.cc:232-249,
// 557-585, 1019-1033, 1294-1306); interleaved (x, y) floats.  The homography multiplies by the reciprocal of w, where
// create_remap divides (c_image_transform.cc:1207-1223) - restated as written.
int ssk_transform_remap_points(const ssk_transform *t, const float *rpts_xy, int n, float *cpts_xy) {
  SSK_REQUIRE(t && (n == 0 || (rpts_xy && cpts_xy)) && n >= 0, "ssk_transform_remap_points: bad argument");
  const float *p = t->params;
  switch (t->motion_type) {
    case SSK_MOTION_TRANSLATION:
      for (int i = 0; i < n; ++i) { cpts_xy[2 * i] = rpts_xy[2 * i] + p[0]; cpts_xy[2 * i + 1] = rpts_xy[2 * i + 1] + p[1]; }
      return SSK_OK;
    case SSK_MOTION_EUCLIDEAN:
    case SSK_MOTION_SCALED_EUCLIDEAN: {
      const float Tx = p[0], Ty = p[1], angle = p[2], scale = t->motion_type == SSK_MOTION_EUCLIDEAN ? t->aux[3] : p[3];
      const float Cx = t->aux[0], Cy = t->aux[1];
      const float sa = std::sin(angle), ca = std::cos(angle);          // float overloads, as the reference calls them here
      for (int i = 0; i < n; ++i) {
        const float xx = rpts_xy[2 * i] - Cx, yy = rpts_xy[2 * i + 1] - Cy;
        cpts_xy[2 * i] = scale * (ca * xx - sa * yy) + Tx;
        cpts_xy[2 * i + 1] = scale * (sa * xx + ca * yy) + Ty;
      }
      return SSK_OK;
    }
    case SSK_MOTION_AFFINE:
      for (int i = 0; i < n; ++i) {
        const float x = rpts_xy[2 * i], y = rpts_xy[2 * i + 1];
        cpts_xy[2 * i] = (p[0] * x + p[1] * y) + p[2];
        cpts_xy[2 * i + 1] = (p[3] * x + p[4] * y) + p[5];
      }
      return SSK_OK;
    case SSK_MOTION_HOMOGRAPHY:
      for (int i = 0; i < n; ++i) {
        const float x = rpts_xy[2 * i], y = rpts_xy[2 * i + 1];
        const float w = 1.f / ((p[6] * x + p[7] * y) + t->aux[2]);
        cpts_xy[2 * i] = ((p[0] * x + p[1] * y) + p[2]) * w;
        cpts_xy[2 * i + 1] = ((p[3] * x + p[4] * y) + p[5]) * w;
      }
      return SSK_OK;
    default:
      set_error("unsupported motion type");
      return SSK_ERR_INVALID;
  }
}

}  // extern "C"

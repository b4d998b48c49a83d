// On-the-fly cost evaluation for the entropic OT sweeps (the N x M matrix is never formed).
//
// Everything is expressed in the scaled log2 domain: with k = log2(e)/eps the sweeps need k*C_ij, so the functors
// return kC directly (and, for the backward, d(kC)/d(owner point)).  Kinds:
//   GEODESIC       acos(<x^,y^>)^p          Point_Cloud_Resistration/losses/s2_wasserstein.py:112-123
//   SQEUCLID       sum_k |x_k-y_k|^p        s2_wasserstein.py:52-63 (line 62), losses/Sinkhorn.py:72-82
//   EUCLID         (sum_k |x_k-y_k|^p)^(1/p)  losses/Sinkhorn_fixed.py:79-89
//   ONE_MINUS_COS  (1-<x^,y^>)^p            losses/max_spherical_w_cos_with_regulation.py:745
//   n_power        C^N                      Comparison_.../losses/sinkhorn.py:165-176 (log_N_Sinkhorn)
//
// Four hand-tuned fast paths -- geodesic p=2 (the north-star configuration), squared-Euclidean p=2 (what train_W_COS.py:393
// instantiates), and their p=1 siblings: geodesic p=1 (the default p of Geodesic_distance_W, s2_wasserstein.py:74) and
// the L1 cost sum_k |x_k-y_k| (Cos_disimilarity_W(p=1), Optimize_hyperparameters/train_W1_COS.py:393; type_of_cost_norm
// 'L1'), the Euclidean norm |x-y|_2 (log_Sinkhorn_Distance_Loss of Sinkhorn_fixed.py:79-89, 'L2') and (1-cos)^2
// (max_spherical_w_cos_with_regulation.py:745) -- plus one generic path that covers every other (kind, p, n_power).
#pragma once
#include "common.cuh"
#include <math.h>

namespace shwd {

enum { FAST_GEO2 = 0, FAST_SQE2 = 1, GENERIC = 2, FAST_GEO1 = 3, FAST_SQE1 = 4, FAST_EUC2 = 5, FAST_OMC2 = 6 };
// "geo" family: the owner's gradient direction is the streamed point (cost is a function of the dot product);
// "sqe" family: it is a function of the difference o - s and the exponent is pot - k * e
__host__ __device__ constexpr bool fast_is_geo(int f) { return f == FAST_GEO2 || f == FAST_GEO1 || f == FAST_OMC2; }
__host__ __device__ constexpr bool fast_is_sqe(int f) { return f == FAST_SQE2 || f == FAST_SQE1 || f == FAST_EUC2; }

struct CostParams {
  int kind;
  float p, npow;
  float k;    // log2(e)/eps
  float sk;   // sqrt(k)
  float q[7]; // qs * Q_i, acos(a) ~= sqrt(1-a) * Q(a) on [0,1] (degree-6 minimax, rel err 1.0e-7);
              // qs = sqrt(k) on the geodesic fast path (so th*th = k*theta^2 directly), 1 on the generic path
  float hpi;  // qs * pi/2
  float gscale;  // constant factor deferred out of the backward inner loop (fast paths)
};

// acos(a)/sqrt(1-a) on [0,1], degree 6, fitted by tools/fit_acos.py (max rel err 1.03e-7 before rounding).
#define SHWD_ACOS_Q                                                                                              \
  { 1.570796132e+00f, -2.145847082e-01f, 8.874903619e-02f, -4.877404124e-02f, 2.684460580e-02f, -1.109675225e-02f, \
    2.279403852e-03f }

// th = s * acos(c) with all constants pre-multiplied by s (s = sqrt(k) in the fast path, 1 in the generic path).
// 11 FP32/ALU issue slots + 1 MUFU.  |w| keeps cos values a hair above 1 finite (the reference returns NaN there).
__device__ __forceinline__ float scaled_acos(const float (&q)[7], float hpi, float c) {
  float a = fabsf(c);
  float w = 1.f - a;
  float r = q[6];
  r = fmaf(r, a, q[5]);
  r = fmaf(r, a, q[4]);
  r = fmaf(r, a, q[3]);
  r = fmaf(r, a, q[2]);
  r = fmaf(r, a, q[1]);
  r = fmaf(r, a, q[0]);
  float sq = sqrt_approx(fabsf(w));  // MUFU.SQRT and MUFU.RSQ issue at the same rate on sm_100 (A/B measured)
  float h = fmaf(-sq, r, hpi);  // s * asin(|c|)
  float hs = copysignf(h, c);
  return hpi - hs;
}

__device__ __forceinline__ float dot3(float ox, float oy, float oz, float sx, float sy, float sz) {
#ifdef SHWD_EXACT_DOT
  // ((x0*y0 + x1*y1) + x2*y2) with separately rounded products, as torch's cosine_similarity
  return __fadd_rn(__fadd_rn(__fmul_rn(ox, sx), __fmul_rn(oy, sy)), __fmul_rn(oz, sz));
#else
  return fmaf(oz, sz, fmaf(oy, sy, ox * sx));
#endif
}

// Per-element cost evaluation.  `E` carries the one expensive intermediate; `m(e, pot)` is THE canonical
// float32 evaluation of (pot - k*C): every sweep (forward and backward) that needs the exponent of a given forward
// half-step calls it with that half-step's streamed potential, so the backward reproduces the forward's roundings
// bit for bit and its softmax factors stay normalised to ~1e-7 (the gradient has a C/eps-fold cancellation between
// the direct term and the adjoint terms, SURVEY.md B.6, which amplifies any inconsistency).
template <int FAST>
struct Cost {
  struct E {
    float a;  // FAST_GEO2 / FAST_GEO1: th = sqrt(k)*theta | FAST_OMC2: sqrt(k)*(1-cos) | FAST_SQE2: |x-y|^2 | FAST_SQE1: |x-y|_1
              // FAST_EUC2: |x-y|_2 | GENERIC: k*C
  };

  static __device__ __forceinline__ float m(const CostParams& cp, E e, float pot) {
    if (FAST == FAST_GEO2 || FAST == FAST_OMC2) return fmaf(-e.a, e.a, pot);
    if (FAST == FAST_GEO1) return fmaf(-cp.sk, e.a, pot);
    if (fast_is_sqe(FAST)) return fmaf(-cp.k, e.a, pot);
    return __fsub_rn(pot, e.a);
  }
  static __device__ __forceinline__ float kc(const CostParams& cp, E e) {
    if (FAST == FAST_GEO2 || FAST == FAST_OMC2) return __fmul_rn(e.a, e.a);
    if (FAST == FAST_GEO1) return __fmul_rn(cp.sk, e.a);
    if (fast_is_sqe(FAST)) return __fmul_rn(cp.k, e.a);
    return e.a;
  }

  // forward
  static __device__ __forceinline__ E eval(const CostParams& cp, float ox, float oy, float oz, float sx, float sy,
                                           float sz) {
    E e;
    if (FAST == FAST_OMC2) {
      e.a = fmaf(-cp.sk, dot3(ox, oy, oz, sx, sy, sz), cp.sk);
    } else if (fast_is_geo(FAST)) {
      e.a = scaled_acos(cp.q, cp.hpi, dot3(ox, oy, oz, sx, sy, sz));
    } else if (FAST == FAST_SQE2 || FAST == FAST_EUC2) {
      float dx = ox - sx, dy = oy - sy, dz = oz - sz;
      e.a = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      if (FAST == FAST_EUC2) e.a = sqrt_approx(e.a);
    } else if (FAST == FAST_SQE1) {
      e.a = (fabsf(ox - sx) + fabsf(oy - sy)) + fabsf(oz - sz);
    } else {
      float C;
      if (cp.kind == SHWD_COST_GEODESIC || cp.kind == SHWD_COST_ONE_MINUS_COS) {
        float c = dot3(ox, oy, oz, sx, sy, sz);
        float base = (cp.kind == SHWD_COST_GEODESIC) ? scaled_acos(cp.q, cp.hpi, c) : 1.f - c;
        C = (cp.p == 1.f) ? base : ((cp.p == 2.f) ? base * base : powf(fmaxf(base, 0.f), cp.p));
      } else {
        float dx = fabsf(ox - sx), dy = fabsf(oy - sy), dz = fabsf(oz - sz);
        float s;
        if (cp.p == 2.f)
          s = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        else if (cp.p == 1.f)
          s = dx + dy + dz;
        else
          s = powf(dx, cp.p) + powf(dy, cp.p) + powf(dz, cp.p);
        C = (cp.kind == SHWD_COST_EUCLID) ? ((cp.p == 2.f) ? sqrtf(s) : ((cp.p == 1.f) ? s : powf(s, 1.f / cp.p))) : s;
      }
      if (cp.npow != 1.f) C = powf(C, cp.npow);
      e.a = __fmul_rn(cp.k, C);
    }
    return e;
  }

  // backward: the same E (bit-identical to eval) plus d(kC)/d(owner point) = gscale * gs * (gx, gy, gz)
  static __device__ __forceinline__ E eval_grad(const CostParams& cp, float ox, float oy, float oz, float sx, float sy,
                                                float sz, float& gs, float& gx, float& gy, float& gz) {
    E e;
    if (FAST == FAST_GEO2) {
      // C = theta^2, dC/dcos = -2 theta / sqrt(1 - cos^2) (torch: acos' = -(1 - x*x).rsqrt()).
      // d(kC)/dx^ = (-2 sqrt(k)) * th * rsqrt(1-c^2) * y^ ; the constant is cp.gscale, applied once per owner.
      float c = dot3(ox, oy, oz, sx, sy, sz);
      e.a = scaled_acos(cp.q, cp.hpi, c);
      float rs = rsqrt_approx(fmaxf(fmaf(-c, c, 1.f), 1e-12f));
      gs = e.a * rs;
      gx = sx;
      gy = sy;
      gz = sz;
    } else if (FAST == FAST_OMC2) {
      // C = (1-c)^2, d(kC)/dx^ = -2k (1-c) y^ = (-2 sqrt(k)) * th * y^ ; the constant is cp.gscale
      e.a = fmaf(-cp.sk, dot3(ox, oy, oz, sx, sy, sz), cp.sk);
      gs = e.a;
      gx = sx;
      gy = sy;
      gz = sz;
    } else if (FAST == FAST_EUC2) {
      // C = |d|, d(kC)/d(owner) = k d / |d| (0 at d = 0: the sub-gradient the generic path returns as well); gscale = k
      float dx = ox - sx, dy = oy - sy, dz = oz - sz;
      float sq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      e.a = sqrt_approx(sq);
      gs = (sq > 0.f) ? rsqrt_approx(sq) : 0.f;
      gx = dx;
      gy = dy;
      gz = dz;
    } else if (FAST == FAST_GEO1) {
      // C = theta, d(kC)/dx^ = -k rsqrt(1-c^2) y^ = (-sqrt(k)) * [sqrt(k) rs] ... the constant -k is cp.gscale
      float c = dot3(ox, oy, oz, sx, sy, sz);
      e.a = scaled_acos(cp.q, cp.hpi, c);
      gs = rsqrt_approx(fmaxf(fmaf(-c, c, 1.f), 1e-12f));
      gx = sx;
      gy = sy;
      gz = sz;
    } else if (FAST == FAST_SQE2) {
      float dx = ox - sx, dy = oy - sy, dz = oz - sz;  // gscale = 2k
      gs = 1.f;
      gx = dx;
      gy = dy;
      gz = dz;
      e.a = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    } else if (FAST == FAST_SQE1) {
      float dx = ox - sx, dy = oy - sy, dz = oz - sz;  // gscale = k; d|d|/dd = sign(d), 0 at 0 (torch.abs backward)
      gs = 1.f;
      gx = (dx > 0.f) - (dx < 0.f);
      gy = (dy > 0.f) - (dy < 0.f);
      gz = (dz > 0.f) - (dz < 0.f);
      e.a = (fabsf(dx) + fabsf(dy)) + fabsf(dz);
    } else {
      float C, dCx, dCy, dCz;  // gscale = k
      if (cp.kind == SHWD_COST_GEODESIC || cp.kind == SHWD_COST_ONE_MINUS_COS) {
        float c = dot3(ox, oy, oz, sx, sy, sz);
        float base, dbase;  // base(c), d base / d c
        if (cp.kind == SHWD_COST_GEODESIC) {
          base = scaled_acos(cp.q, cp.hpi, c);
          dbase = -rsqrtf(fmaxf(fmaf(-c, c, 1.f), 1e-12f));
        } else {
          base = 1.f - c;
          dbase = -1.f;
        }
        float dC;
        if (cp.p == 1.f) {
          C = base;
          dC = dbase;
        } else if (cp.p == 2.f) {
          C = base * base;
          dC = 2.f * base * dbase;
        } else {
          base = fmaxf(base, 0.f);
          C = powf(base, cp.p);
          dC = cp.p * powf(base, cp.p - 1.f) * dbase;
        }
        dCx = dC * sx;
        dCy = dC * sy;
        dCz = dC * sz;
      } else {
        float ex = ox - sx, ey = oy - sy, ez = oz - sz;
        float dx = fabsf(ex), dy = fabsf(ey), dz = fabsf(ez);
        float s, tx, ty, tz;  // s = sum |d|^p, t = d s / d owner
        if (cp.p == 2.f) {
          s = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          tx = 2.f * ex;
          ty = 2.f * ey;
          tz = 2.f * ez;
        } else if (cp.p == 1.f) {
          s = dx + dy + dz;
          tx = (ex > 0.f) - (ex < 0.f);
          ty = (ey > 0.f) - (ey < 0.f);
          tz = (ez > 0.f) - (ez < 0.f);
        } else {
          s = powf(dx, cp.p) + powf(dy, cp.p) + powf(dz, cp.p);
          tx = copysignf(cp.p * powf(dx, cp.p - 1.f), ex);
          ty = copysignf(cp.p * powf(dy, cp.p - 1.f), ey);
          tz = copysignf(cp.p * powf(dz, cp.p - 1.f), ez);
        }
        if (cp.kind == SHWD_COST_EUCLID && cp.p != 1.f) {
          C = (cp.p == 2.f) ? sqrtf(s) : powf(s, 1.f / cp.p);
          // d s^(1/p) = C / (p s) ds   (s == 0: torch's pow backward gives inf*0 = nan; we return the sub-gradient 0)
          float f = (s > 0.f) ? C / (cp.p * s) : 0.f;
          tx *= f;
          ty *= f;
          tz *= f;
        } else {
          C = s;
        }
        dCx = tx;
        dCy = ty;
        dCz = tz;
      }
      if (cp.npow != 1.f) {
        float f = cp.npow * powf(C, cp.npow - 1.f);
        C = powf(C, cp.npow);
        dCx *= f;
        dCy *= f;
        dCz *= f;
      }
      gs = 1.f;
      gx = dCx;
      gy = dCy;
      gz = dCz;
      e.a = __fmul_rn(cp.k, C);
    }
    return e;
  }
};

// Host side: which specialised kernel a (kind, p, n_power) runs on, and its constants.
inline int pick_fast(int kind, float p, float npow) {
  if (kind == SHWD_COST_GEODESIC && p == 2.f && npow == 1.f) return FAST_GEO2;
  if (kind == SHWD_COST_SQEUCLID && p == 2.f && npow == 1.f) return FAST_SQE2;
  if (kind == SHWD_COST_GEODESIC && p == 1.f && npow == 1.f) return FAST_GEO1;
  if ((kind == SHWD_COST_SQEUCLID || kind == SHWD_COST_EUCLID) && p == 1.f && npow == 1.f) return FAST_SQE1;  // (s)^(1/1) = s
  if (kind == SHWD_COST_EUCLID && p == 2.f && npow == 1.f) return FAST_EUC2;
  if (kind == SHWD_COST_ONE_MINUS_COS && p == 2.f && npow == 1.f) return FAST_OMC2;
  return GENERIC;
}

inline CostParams make_cost(int kind, float p, float npow, float eps, int fast) {
  CostParams cp;
  cp.kind = kind;
  cp.p = p;
  cp.npow = npow;
  const double k = 1.4426950408889634 / (double)eps;
  cp.k = (float)k;
  cp.sk = (float)sqrt(k);
  const float q[7] = SHWD_ACOS_Q;
  const double qs = (fast == FAST_GEO2 || fast == FAST_GEO1) ? sqrt(k) : 1.0;
  for (int i = 0; i < 7; ++i) cp.q[i] = (float)(qs * (double)q[i]);
  cp.hpi = (float)(qs * 1.5707963267948966);
  cp.gscale = (fast == FAST_GEO2 || fast == FAST_OMC2) ? (float)(-2.0 * sqrt(k))
              : (fast == FAST_GEO1) ? (float)(-k) : (fast == FAST_SQE2) ? (float)(2.0 * k) : (float)k;
  return cp;
}

// k == 1 exactly: the functors then return the plain float32 cost C (no scaling rounding) -- the exact-assignment kernel
// must order assignments by the very numbers the reference's cost matrix holds.
inline CostParams make_cost_unit(int kind, float p, float npow) {
  CostParams cp;
  cp.kind = kind;
  cp.p = p;
  cp.npow = npow;
  cp.k = 1.f;
  cp.sk = 1.f;
  const float q[7] = SHWD_ACOS_Q;
  for (int i = 0; i < 7; ++i) cp.q[i] = q[i];
  cp.hpi = 1.5707963267948966f;
  cp.gscale = 1.f;
  return cp;
}
}  // namespace shwd

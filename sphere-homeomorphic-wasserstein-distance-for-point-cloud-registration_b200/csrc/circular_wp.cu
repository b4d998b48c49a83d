// Circular Wasserstein W_p^p, p != 1, on sorted circle coordinates: the Delon-Salomon-Sobolevski bisection on the
// rotation theta, every round of it and the final cost + gradients in ONE launch, one CTA per slice.
//
// Replaces   binary_search_circle   Point_Cloud_Resistration/losses/max_spherical_sliced_w.py:117-207
//            dCost                  max_spherical_sliced_w.py:25-65
//            Cost                   max_spherical_sliced_w.py:68-113
//            roll_by_gather         max_spherical_sliced_w.py:9-22
// (the reference runs ~26 rounds of torch ops with a host sync per round; each round materialises rolled copies of the
// value / CDF arrays, two searchsorted results and, for Cost, a sort of the concatenated CDFs).
//
// Nothing is materialised here.  With uniform weights the CDFs are the same for every slice
// (u_cdf = cumsum(1/n), v_cdf = cumsum(1/m), accumulated in double and rounded per element like torch's CPU cumsum --
// evaluated in closed form in registers, see ucdf_at),
// the "shifted + rolled" arrays of the reference are pure index arithmetic on them:
//     frac = theta - floor(theta);  c_j = v_cdf_j - frac;  neg_j = c_j < 0   (a prefix: v_cdf is increasing)
//     j0 = #neg (0 if all are negative -- argmin over an all-inf row)        (the roll)
//     r_cdf[t] = c_{(t+j0)%m} (+1 if negative),  r_val[t] = v_{(t+j0)%m} + floor(theta) (+1 if negative),
//     r_val[m] = r_val[0] + 1
// and both searchsorted calls and the sort of cat(u_cdf, r_cdf) reduce to local probes around a closed-form guess,
// because both CDFs are (nearly) arithmetic progressions.  The merged CDF axis of Cost is walked block-wise:
//   * pass U: thread k owns the axis entries whose u-index is k (axis values in (u_cdf[k-1], u_cdf[k]]) -- cost and
//     d cost / d u_sorted[k] with no cross-thread reduction;
//   * pass V: thread t owns the entries whose v-index is t (axis values in (r_cdf[t-1], r_cdf[t]]) -- d cost / d v.
// Sums are reduced in a fixed order (bit-reproducible run to run).
//
// Reference quirks kept: the arc bookkeeping of Cost (v_0 + 1 appended, index clips at n-1 / m), theta detached in
// the final Cost (:207), the secant point only where |dCp(tm) - dCm(tp)| > 1e-3 (:196-197), `done` re-evaluated from
// dCp * dCm <= 0 every round.  The batch-global test "re-wrap the negative CDF entries only if the call holds both signs"
// (:41-42, :82-83) differs from "always re-wrap" only when EVERY entry of EVERY slice of the call is negative in the same
// evaluation, i.e. frac > v_cdf[m-1] in all slices at once.  v_cdf[m-1] = fl(m * fl(1/m)) >= 1 - 2^-24 and frac <= 1, so
// that needs frac == 1.0 exactly: theta in [-2^-25, 0) (the rounded theta - floor(theta) is then 1.0) for a cloud size m
// whose last CDF entry rounds below 1.  A call of ONE slice decides that locally and is reproduced (Shift::keep_neg: the
// CDF stays negative and un-rolled, exactly the reference's arithmetic); in a call of several slices every slice would
// have to sit in that 3e-8 window in the same round, and the kernel re-wraps like the reference does whenever at least one
// slice has a non-negative entry.
// The rounds of different slices are independent in the reference as well: every unfinished slice halves the same
// dyadic bracket each round, so all of them reach the stopping width in the same round.
#include "common.cuh"

namespace shwd {

constexpr int CW_THREADS = 256;
constexpr int CW_WARPS = CW_THREADS / 32;
constexpr int CW_MAX_ROUNDS = 96;  // the bracket [-1,1] reaches 1 ulp after ~25 halvings; a safety net, never hit

struct Circle {
  const float* u;     // sorted u values (n)            -- shared memory
  const float* v;     // sorted v values (m)
  int n, m;
  bool lone;     // the call holds a single slice (the batch-global "both signs present" test is then local, see make_shift)
  float p;
  float wu, wv;  // float32(1/n), float32(1/m): torch.full((len,), 1/len, dtype=float32)
  const float* ucdf;  // W (caller-supplied weights) only: the slice's CDF tables cumsum(weights[sorter]) -- shared memory
  const float* vcdf;
};

// u_cdf[i] = cumsum(full(1/n))[i] the way torch's CPU kernel forms it -- a double accumulator, every prefix rounded to
// float32 -- WITHOUT a table: the weight is a float32 (24 significant bits) and at most 2^15 copies are added, so every
// double partial sum is exact ((i+1) * w needs < 40 bits) and its float32 rounding is the single rounding of the exact
// product, i.e. __fmul_rn(float(i+1), w).  Two ALU instructions instead of a shared-memory load on the hottest path of the
// bisection (dcost was bound by shared-memory wavefronts: ~12 loads per CDF entry, 7 of them from the two CDF tables).
// W: non-uniform weights (u_weights / v_weights of the reference, :156-170) -- the CDFs are tables, every closed-form guess
// below becomes a binary search; the structure of the passes is unchanged (they only need the CDFs to be non-decreasing).
template <bool W>
__device__ __forceinline__ float ucdf_at(const Circle& c, int i) { return W ? c.ucdf[i] : __fmul_rn(__int2float_rn(i + 1), c.wu); }
template <bool W>
__device__ __forceinline__ float vcdf_at(const Circle& c, int j) { return W ? c.vcdf[j] : __fmul_rn(__int2float_rn(j + 1), c.wv); }

struct Shift {
  float fl, flp1, frac, r0;
  int j0;
  bool keep_neg;  // every entry negative in a call of ONE slice: the reference leaves the CDF un-wrapped (:41-42, :82-83)
  bool allneg;
};

template <bool P2>
__device__ __forceinline__ float powp(float d, float p) {
  return P2 ? d * d : powf(fabsf(d), p);
}
// d/dd |d|^p
template <bool P2>
__device__ __forceinline__ float dpowp(float d, float p) {
  if (P2) return 2.f * d;
  const float a = fabsf(d);
  if (a == 0.f) return 0.f;  // autograd: pow'(0) * sign(0) = 0
  return copysignf(p * powf(a, p - 1.f), d);
}

template <bool W>
__device__ __forceinline__ float r_cdf(const Circle& c, const Shift& s, int t) {
  int j = t + s.j0;
  if (j >= c.m) j -= c.m;
  const float x = __fsub_rn(vcdf_at<W>(c, j), s.frac);
  return (x < 0.f && !s.keep_neg) ? __fadd_rn(x, 1.f) : x;
}
template <bool W>
__device__ __forceinline__ float r_val_in(const Circle& c, const Shift& s, int t) {
  int j = t + s.j0;
  if (j >= c.m) j -= c.m;
  const float x = __fsub_rn(vcdf_at<W>(c, j), s.frac);
  return __fadd_rn(c.v[j], x < 0.f ? s.flp1 : s.fl);
}
// t in [0, m]: r_val[m] = r_val[0] + 1
template <bool W>
__device__ __forceinline__ float r_val(const Circle& c, const Shift& s, int t) {
  return t >= c.m ? __fadd_rn(r_val_in<W>(c, s, 0), 1.f) : r_val_in<W>(c, s, t);
}

template <bool W>
__device__ __forceinline__ Shift make_shift(const Circle& c, float theta) {
  Shift s;
  s.fl = floorf(theta);
  s.flp1 = __fadd_rn(s.fl, 1.f);
  s.frac = __fsub_rn(theta, s.fl);
  // first j with v_cdf_j - frac >= 0 (v_cdf is increasing): closed-form guess, then a local walk (W: bisection)
  int lo;
  if (W) {
    lo = 0;
    for (int hi = c.m; lo < hi;) {
      const int mid = (lo + hi) >> 1;
      if (__fsub_rn(c.vcdf[mid], s.frac) < 0.f) lo = mid + 1; else hi = mid;
    }
  } else {
    lo = min(max(__float2int_rd(s.frac * (float)c.m), 0), c.m);
    while (lo < c.m && __fsub_rn(vcdf_at<W>(c, lo), s.frac) < 0.f) ++lo;
    while (lo > 0 && !(__fsub_rn(vcdf_at<W>(c, lo - 1), s.frac) < 0.f)) --lo;
  }
  s.j0 = (lo == c.m) ? 0 : lo;
  s.allneg = (lo == c.m);
  s.keep_neg = s.allneg && c.lone;
  s.r0 = 0.f;
  s.r0 = r_cdf<W>(c, s, 0);
  return s;
}

// #{i : u_cdf[i] < x}  (torch.searchsorted(u_cdf, x), left), in [0, n] -- two probes, no loop.  u_cdf[i] = fl((i+1) fl(1/n))
// = (i+1)/n (1 + e), |e| < 2^-22, so entry i is surely below x when i+1 <= xn - 2^-7 and surely not when i+1 >= xn + 2^-7
// (xn <= 2^15); fl(x * n) is within 2^-9 of xn, hence with g = floor(fl(x * n)) the count is g - 1, g or g + 1: entries
// g - 1 and g (0-based) decide.
template <bool W>
__device__ __forceinline__ int u_count_lt(const Circle& c, float x) {
  if (W) {
    int lo = 0;
    for (int hi = c.n; lo < hi;) {
      const int mid = (lo + hi) >> 1;
      if (c.ucdf[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  const int g = __float2int_rd(x * (float)c.n);
  const int b = min(max(g, 1) - 1, c.n);
  return b + (int)(b < c.n && ucdf_at<W>(c, b) < x) + (int)(b + 1 < c.n && ucdf_at<W>(c, b + 1) < x);
}
// #{i : u_cdf^+[i] <= x} for u_cdf^+ = cat(u_cdf, u_cdf[0] + 1)  (searchsorted(..., right=True)), given iu = u_count_lt(x):
// u_cdf is strictly increasing (steps of 1/n >> its rounding), so at most the entry at iu equals x.
template <bool W>
__device__ __forceinline__ int u_count_le(const Circle& c, float x, int iu, float ucdf_wrap) {
  int ium = iu + (int)(iu < c.n && ucdf_at<W>(c, iu) <= x);
  if (W) {  // equal CDF entries (zero weights) are possible: walk over all of them
    while (ium > iu && ium < c.n && c.ucdf[ium] <= x) ++ium;
  }
  if (ium == c.n && ucdf_wrap <= x) ++ium;
  return ium;
}
// #{t : r_cdf[t] < x}, in [0, m]
template <bool W>
__device__ __forceinline__ int r_count_lt(const Circle& c, const Shift& s, float x) {
  if (W) {
    int lo = 0;
    for (int hi = c.m; lo < hi;) {
      const int mid = (lo + hi) >> 1;
      if (r_cdf<W>(c, s, mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  int t = min(max(__float2int_rd((x - s.r0) * (float)c.m) + 1, 0), c.m);
  while (t < c.m && r_cdf<W>(c, s, t) < x) ++t;
  while (t > 0 && r_cdf<W>(c, s, t - 1) >= x) --t;
  return t;
}

// Fixed-order block sum of two values; result broadcast to every thread.
template <int T>
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* wtot) {
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) wtot[threadIdx.x >> 5] = make_float2(a, b);
  __syncthreads();
  float2 t = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < T / 32; ++w) {
    t.x += wtot[w].x;
    t.y += wtot[w].y;
  }
  return t;
}

// What a thread remembers of one dCost evaluation (kept for the two ends of the bisection bracket).
//
// dCost(theta) depends on theta only through (floor(theta), j0, which CDF entries are negative) and, per rolled entry t,
// through the two search results iu_t = #{u_cdf < r_cdf[t]} and ium_t = #{u_cdf^+ <= r_cdf[t]}: the summands themselves are
// built from values (u, v + integer), not from frac.  C(theta) is piecewise linear and the bisection spends its second half
// inside one or two linear pieces, re-evaluating the very same sum (n = m = 4096: pieces are 2^-12 wide, the bracket ends at
// 2^-24).  For theta between the bracket ends and sharing (floor, j0, all-negative) with an end E, every r_cdf[t] is a
// monotone function of frac (a rounded subtraction, then possibly a rounded +1), so iu_t and ium_t are monotone in theta:
// iu_t(theta) <= iu_t(tm), >= iu_t(tp), and therefore   sum_t iu_t(theta) == sum_t iu_t(E)  <=>  iu_t(theta) == iu_t(E) for
// every t  (the same for ium).  A thread whose two sums over ITS entries equal the end's has exactly the end's summands in
// the end's order: its partial sums are the end's, bit for bit, and it skips the powers, the value loads and the clipping.
// The searches alone are ~1/4 of an entry's instructions.  Whether a round tries the searches first is decided from the share
// of warps that could have skipped in the previous round (clouds of different sizes have their kinks spread out, and a warp
// only skips if none of its 32 x m/T entries has one inside the bracket).
// Equal cloud sizes that are a power of two never get there: their kinks sit on the dyadic grid the midpoints walk, so the
// bisection stops on "dCp * dCm <= 0" at round log2(n) (cfg3: round 12 of 4096 points) -- those calls run MEMO = false, which
// keeps the kernel at 40 registers and six CTAs per SM.
struct DcMemo {
  float fl;
  int j0;       // with the all-negative flag in bit 30; -1: nothing remembered
  int siu, sium;
  float dcp, dcm;  // this thread's partial sums
};

template <int T>
__device__ __forceinline__ float2 block_sum2_vote(float a, float b, bool warp_skipped, float2* wtot, int* wskip, int& n_skipped) {
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    wtot[threadIdx.x >> 5] = make_float2(a, b);
    wskip[threadIdx.x >> 5] = warp_skipped ? 1 : 0;
  }
  __syncthreads();
  float2 t = make_float2(0.f, 0.f);
  int k = 0;
#pragma unroll
  for (int w = 0; w < T / 32; ++w) {
    t.x += wtot[w].x;
    t.y += wtot[w].y;
    k += wskip[w];
  }
  n_skipped = k;
  return t;
}

// dCost :25-65 -> (dCp, dCm): right / left derivative of the cost in theta.  lo / hi: the evaluations at the bracket ends
// (tm / tp; tm <= theta <= tp); out: this evaluation.  searches_first: see DcMemo (uniform over the CTA; updated).
template <bool P2, int T, bool MEMO, bool W>
__device__ float2 dcost(const Circle& c, float theta, float2* wtot, int* wskip, const DcMemo& lo, const DcMemo& hi, DcMemo& out,
                        bool& searches_first) {
  const Shift s = make_shift<W>(c, theta);
  const float u_wrap = __fadd_rn(c.u[0], 1.f), ucdf_wrap = __fadd_rn(ucdf_at<W>(c, 0), 1.f);
  const int key = s.j0 | (s.allneg ? (1 << 30) : 0);
  const bool near_lo = MEMO && lo.fl == s.fl && lo.j0 == key;
  const bool near_hi = MEMO && hi.fl == s.fl && hi.j0 == key;
  float dcp = 0.f, dcm = 0.f;
  int siu = 0, sium = 0;
  bool same = false;  // this thread's summands are those of a bracket end
  if (MEMO && searches_first && (near_lo || near_hi)) {
    for (int t = threadIdx.x; t < c.m; t += T) {
      const float x = r_cdf<W>(c, s, t);
      const int iu = u_count_lt<W>(c, x);
      const int ium = u_count_le<W>(c, x, iu, ucdf_wrap);
      siu += iu;
      sium += ium;
    }
    if (near_lo && siu == lo.siu && sium == lo.sium) {
      same = true; dcp = lo.dcp; dcm = lo.dcm;
    } else if (near_hi && siu == hi.siu && sium == hi.sium) {
      same = true; dcp = hi.dcp; dcm = hi.dcm;
    }
  }
  if (!same) {
    siu = 0;
    sium = 0;
    for (int t = threadIdx.x; t < c.m; t += T) {
      const float x = r_cdf<W>(c, s, t);
      const int iu = u_count_lt<W>(c, x);                       // searchsorted(u_cdf, x)
      const float ui = c.u[min(iu, c.n - 1)];
      const int ium = u_count_le<W>(c, x, iu, ucdf_wrap);       // searchsorted(cat(u_cdf, u_cdf_0 + 1), x, right=True)
      const float uim = (ium < c.n) ? c.u[ium] : u_wrap;     // index clipped at n -> u_0 + 1
      const float v0 = r_val_in<W>(c, s, t), v1 = r_val<W>(c, s, t + 1);
      dcp += __fsub_rn(powp<P2>(__fsub_rn(ui, v1), c.p), powp<P2>(__fsub_rn(ui, v0), c.p));
      dcm += __fsub_rn(powp<P2>(__fsub_rn(uim, v1), c.p), powp<P2>(__fsub_rn(uim, v0), c.p));
      if (MEMO) {
        siu += iu;
        sium += ium;
      }
    }
    // (would the searches alone have sufficed?  feeds the next round's decision)
    same = (near_lo && siu == lo.siu && sium == lo.sium) || (near_hi && siu == hi.siu && sium == hi.sium);
  }
  if (!MEMO) return block_sum2<T>(dcp, dcm, wtot);
  int n_skipped;
  const float2 tot = block_sum2_vote<T>(dcp, dcm, __all_sync(0xffffffffu, same), wtot, wskip, n_skipped);
  searches_first = 2 * n_skipped >= T / 32;
  out.fl = s.fl; out.j0 = key; out.siu = siu; out.sium = sium; out.dcp = dcp; out.dcm = dcm;
  return tot;
}

// Cost :68-113, pass U.  Returns the cost (broadcast); gu (nullable, global) receives d cost / d u_sorted.
// gcu (W only, nullable): d cost / d u_cdf[k].  The cost is sum_a delta_a h_a over the merged axis (delta_a = a - previous
// entry, h_a = |u_icdf - v_icdf|^p at a), so d cost / d a = h_a - h_next(a): the entry's own h minus that of the entry
// that follows it on the axis (the searches are not differentiated, like torch.searchsorted).
template <bool P2, int T, bool W>
__device__ float cost_pass_u(const Circle& c, float theta, float* __restrict__ gu, const int32_t* __restrict__ pu, float2* wtot,
                             float* __restrict__ gcu = nullptr) {
  const Shift s = make_shift<W>(c, theta);
  float acc = 0.f;
  for (int k = threadIdx.x; k < c.n; k += T) {
    float prev = (k > 0) ? ucdf_at<W>(c, k - 1) : 0.f;
    const float xk = ucdf_at<W>(c, k);
    const float U = c.u[k];
    int t = (k > 0) ? r_count_lt<W>(c, s, prev) : 0;
    float g = 0.f;
    // v entries below u_cdf[k]: v-index t, u-index k
    while (t < c.m) {
      const float rt = r_cdf<W>(c, s, t);
      if (rt >= xk) break;
      const float delta = __fsub_rn(rt, prev), d = __fsub_rn(U, r_val_in<W>(c, s, t));
      acc = fmaf(delta, powp<P2>(d, c.p), acc);
      g = fmaf(delta, dpowp<P2>(d, c.p), g);
      prev = rt;
      ++t;
    }
    {  // the u entry itself: v-index = #{r < u_cdf[k]} = t (clipped at m)
      const float delta = __fsub_rn(xk, prev), d = __fsub_rn(U, r_val<W>(c, s, t));
      const float h = powp<P2>(d, c.p);
      acc = fmaf(delta, h, acc);
      g = fmaf(delta, dpowp<P2>(d, c.p), g);
      prev = xk;
      if (W && gcu) {
        // the entry after u_cdf[k]: the next rolled v entry r[t] (>= u_cdf[k]) if it lies below u_cdf[k+1] (equal values: the u
        // entry sorts first), else the u entry k + 1, else nothing.  Equal CDF entries (zero weights): searchsorted returns
        // the FIRST index of the group, so every member of it reads the group's first value (kk).
        int kk = k;
        while (kk > 0 && ucdf_at<W>(c, kk - 1) == xk) --kk;
        const float h0 = (kk == k) ? h : powp<P2>(__fsub_rn(c.u[kk], r_val<W>(c, s, t)), c.p);
        float hn = 0.f;
        const bool last = (k == c.n - 1);
        const float rt = (t < c.m) ? r_cdf<W>(c, s, t) : 0.f;
        if (t < c.m && (last || rt < ucdf_at<W>(c, k + 1))) {
          const int iu = min((rt > xk) ? k + 1 : kk, c.n - 1);  // #{u_cdf < r[t]}
          int tt = t;                                             // #{r < r[t]}
          while (tt > 0 && r_cdf<W>(c, s, tt - 1) == rt) --tt;
          hn = powp<P2>(__fsub_rn(c.u[iu], r_val_in<W>(c, s, tt)), c.p);
        } else if (!last) {
          hn = powp<P2>(__fsub_rn(c.u[(ucdf_at<W>(c, k + 1) == xk) ? kk : k + 1], r_val<W>(c, s, t)), c.p);
        }
        gcu[k] = h0 - hn;
      }
    }
    if (k == c.n - 1) {  // axis entries above u_cdf[n-1]: u-index clipped to n-1
      while (t < c.m) {
        const float rt = r_cdf<W>(c, s, t);
        const float delta = __fsub_rn(rt, prev), d = __fsub_rn(U, r_val_in<W>(c, s, t));
        acc = fmaf(delta, powp<P2>(d, c.p), acc);
        g = fmaf(delta, dpowp<P2>(d, c.p), g);
        prev = rt;
        ++t;
      }
    }
    if (gu) gu[pu ? __ldg(pu + k) : k] = g;  // pu: straight to the unsorted key position
  }
  return block_sum2<T>(acc, 0.f, wtot).x;
}

// Cost, pass V: gv (global, indexed by sorted v position) receives d cost / d v_sorted; gcv (W only, nullable)
// d cost / d v_cdf[j] (see cost_pass_u; d v_cdf_theta / d v_cdf = 1, theta is detached).
template <bool P2, int T, bool W>
__device__ void cost_pass_v(const Circle& c, float theta, float* __restrict__ gv, const int32_t* __restrict__ pv,
                            float* __restrict__ gcv = nullptr) {
  const Shift s = make_shift<W>(c, theta);
  for (int t = threadIdx.x; t < c.m; t += T) {
    float prev = (t > 0) ? r_cdf<W>(c, s, t - 1) : 0.f;
    const float xt = r_cdf<W>(c, s, t);
    const float V = r_val_in<W>(c, s, t);
    int i = 0;
    if (t > 0) {  // first u entry above r_cdf[t-1]
      i = u_count_lt<W>(c, prev);
      while (i < c.n && ucdf_at<W>(c, i) <= prev) ++i;
    }
    float g = 0.f;
    while (i < c.n && ucdf_at<W>(c, i) < xt) {  // u entries inside (r[t-1], r[t]): v-index t
      const float a = ucdf_at<W>(c, i);
      g = fmaf(__fsub_rn(a, prev), dpowp<P2>(__fsub_rn(c.u[i], V), c.p), g);
      prev = a;
      ++i;
    }
    // the v entry itself: u-index = #{u_cdf < r[t]} = i (clipped at n-1)
    g = fmaf(__fsub_rn(xt, prev), dpowp<P2>(__fsub_rn(c.u[min(i, c.n - 1)], V), c.p), g);
    if (t == 0) {
      // the wrap-around partner r_val[m] = r_val[0] + 1 serves the u entries above r_cdf[m-1]
      const float Vw = __fadd_rn(V, 1.f);
      float pw = r_cdf<W>(c, s, c.m - 1);
      int iw = u_count_lt<W>(c, pw);
      while (iw < c.n && ucdf_at<W>(c, iw) <= pw) ++iw;
      for (; iw < c.n; ++iw) {
        const float a = ucdf_at<W>(c, iw);
        g = fmaf(__fsub_rn(a, pw), dpowp<P2>(__fsub_rn(c.u[iw], Vw), c.p), g);
        pw = a;
      }
    }
    int j = t + s.j0;
    if (j >= c.m) j -= c.m;
    gv[pv ? __ldg(pv + j) : j] = -g;
    if (W && gcv) {
      // own entry: u-index #{u_cdf < r[t]} = i, v-index #{r < r[t]} = the first member of r[t]'s group of equal entries (tt)
      int tt = t;
      while (tt > 0 && r_cdf<W>(c, s, tt - 1) == xt) --tt;
      const float h = powp<P2>(__fsub_rn(c.u[min(i, c.n - 1)], r_val_in<W>(c, s, tt)), c.p);
      // the entry after r[t]: the first u entry above it (index i2) if it does not exceed r[t+1] (equal values: u first), else
      // the v entry t + 1, else nothing
      int i2 = i;
      while (i2 < c.n && ucdf_at<W>(c, i2) <= xt) ++i2;
      const bool lastv = (t == c.m - 1);
      const float rn = lastv ? 0.f : r_cdf<W>(c, s, t + 1);
      float hn = 0.f;
      if (i2 < c.n && (lastv || ucdf_at<W>(c, i2) <= rn)) {
        hn = powp<P2>(__fsub_rn(c.u[i2], r_val<W>(c, s, t + 1)), c.p);  // v-index #{r < u_cdf[i2]} = t + 1 (m: the wrap entry)
      } else if (!lastv) {
        const bool tie = (rn == xt);
        hn = powp<P2>(__fsub_rn(c.u[min(tie ? i : i2, c.n - 1)], r_val_in<W>(c, s, tie ? tt : t + 1)), c.p);
      }
      gcv[j] = h - hn;
    }
  }
}

// ---- equal power-of-two cloud sizes, rotation on the 1/n grid -------------------------------------------------------------
// n == m == 2^k with uniform weights: u_cdf[i] = v_cdf[i] = (i+1) / n EXACTLY (1/n is a power of two), and the bisection's
// midpoints are dyadic: as long as theta is a multiple of 1/n -- every round up to round k, where those calls end on a kink (see
// DcMemo) -- the shifted CDF r_cdf[t] is an exact multiple of 1/n as well, q = frac * n being an integer:
//     q >= 1:  r_cdf[t] = t / n          q == 0:  r_cdf[t] = (t + 1) / n          (t = 0 .. m-1, after the roll by j0 = max(q-1, 0))
// so every search has a closed form:  #{u_cdf < r_cdf[t]} = max(t-1, 0) | t,  #{u_cdf^+ <= r_cdf[t]} = t | t+1,
// #{r_cdf < u_cdf[k]} = k+1 | k, and the merged axis of Cost consists of coincident (u, v) pairs: the entries the walks of the
// generic passes visit besides the owner's own carry delta == 0 exactly (fmaf(0, h, acc) == acc) and are skipped.  The
// summands, their operands and their order per thread are those of the generic code, so the results are the same bits
// (shwd_circular_wp_set_dyadic(0) switches the shortcut off: the tests compare the two).  ~25 instead of ~100 instructions
// per CDF entry and round.
// between: the off-grid form of equal sizes (see between_safe below) -- #{u_cdf < r_cdf[t]} = #{u_cdf^+ <= r_cdf[t]} = t.
template <bool P2, int T>
__device__ float2 dcost_dyadic(const Circle& c, float theta, float2* wtot, bool between = false, DcMemo* memo = nullptr) {
  const Shift s = make_shift<false>(c, theta);
  int siu = 0, sium = 0;
  const bool q0 = (s.frac == 0.f);
  const float u_wrap = __fadd_rn(c.u[0], 1.f);
  const float v_wrap = __fadd_rn(__fadd_rn(c.v[s.j0], s.fl), 1.f);  // r_val[m] = r_val[0] + 1 (entry j0 is never negative)
  float dcp = 0.f, dcm = 0.f;
  for (int t = threadIdx.x; t < c.m; t += T) {
    int j = t + s.j0;
    const bool wr = j >= c.m;  // wrapped <=> its shifted CDF entry was negative
    if (wr) j -= c.m;
    const float v0 = __fadd_rn(c.v[j], wr ? s.flp1 : s.fl);
    float v1 = v_wrap;
    if (t + 1 < c.m) {
      int j1 = t + 1 + s.j0;
      const bool wr1 = j1 >= c.m;
      if (wr1) j1 -= c.m;
      v1 = __fadd_rn(c.v[j1], wr1 ? s.flp1 : s.fl);
    }
    const int iu = (q0 || between) ? t : max(t - 1, 0), ium = (q0 && !between) ? t + 1 : t;
    const float ui = c.u[min(iu, c.n - 1)];
    const float uim = (ium < c.n) ? c.u[ium] : u_wrap;
    dcp += __fsub_rn(powp<P2>(__fsub_rn(ui, v1), c.p), powp<P2>(__fsub_rn(ui, v0), c.p));
    dcm += __fsub_rn(powp<P2>(__fsub_rn(uim, v1), c.p), powp<P2>(__fsub_rn(uim, v0), c.p));
    siu += iu;
    sium += ium;
  }
  if (memo) {  // what the generic evaluation would remember of this rotation (see DcMemo): the two search results are the counts
    memo->fl = s.fl;
    memo->j0 = s.j0 | (s.allneg ? (1 << 30) : 0);
    memo->siu = siu;
    memo->sium = sium;
    memo->dcp = dcp;
    memo->dcm = dcm;
  }
  return block_sum2<T>(dcp, dcm, wtot);
}

template <bool P2, int T>
__device__ float cost_pass_u_dyadic(const Circle& c, float theta, float* __restrict__ gu, const int32_t* __restrict__ pu, float2* wtot) {
  const Shift s = make_shift<false>(c, theta);
  const bool q0 = (s.frac == 0.f);
  float acc = 0.f;
  for (int k = threadIdx.x; k < c.n; k += T) {
    const float prev = (k > 0) ? ucdf_at<false>(c, k - 1) : 0.f;
    const float xk = ucdf_at<false>(c, k);
    const int t = q0 ? k : k + 1;  // #{r_cdf < u_cdf[k]}; m: the wrap entry
    const float delta = __fsub_rn(xk, prev), d = __fsub_rn(c.u[k], r_val<false>(c, s, t));
    acc = fmaf(delta, powp<P2>(d, c.p), acc);
    if (gu) gu[pu ? __ldg(pu + k) : k] = fmaf(delta, dpowp<P2>(d, c.p), 0.f);
  }
  return block_sum2<T>(acc, 0.f, wtot).x;
}

template <bool P2, int T>
__device__ void cost_pass_v_dyadic(const Circle& c, float theta, float* __restrict__ gv, const int32_t* __restrict__ pv) {
  const Shift s = make_shift<false>(c, theta);
  const bool q0 = (s.frac == 0.f);
  for (int t = threadIdx.x; t < c.m; t += T) {
    const float prev = (t > 0) ? r_cdf<false>(c, s, t - 1) : 0.f;
    const float xt = r_cdf<false>(c, s, t);
    const float V = r_val_in<false>(c, s, t);
    const int i = q0 ? t : max(t - 1, 0);  // #{u_cdf < r_cdf[t]}
    float g = fmaf(__fsub_rn(xt, prev), dpowp<P2>(__fsub_rn(c.u[min(i, c.n - 1)], V), c.p), 0.f);
    if (t == 0 && !q0) {
      // the wrap-around partner r_val[m] = r_val[0] + 1 serves the one u entry above r_cdf[m-1] = (m-1)/n: u_cdf[n-1] = 1
      const float pw = r_cdf<false>(c, s, c.m - 1);
      g = fmaf(__fsub_rn(ucdf_at<false>(c, c.n - 1), pw), dpowp<P2>(__fsub_rn(c.u[c.n - 1], __fadd_rn(V, 1.f)), c.p), g);
    }
    int j = t + s.j0;
    if (j >= c.m) j -= c.m;
    gv[pv ? __ldg(pv + j) : j] = -g;
  }
}

// ---- equal cloud sizes of ANY length, rotation safely OFF the 1/n grid ---------------------------------------------------------
// n == m, uniform weights: u_cdf[i] = v_cdf[i] = fl((i+1) w), w = fl(1/n).  With F = frac / w and phi = F - floor(F), the shifted
// entry is r_cdf[t] = (t + 1 - phi) / n up to rounding: comparing it with a u_cdf entry compares an integer with F, perturbed by
// at most n (2^-22 + 2^-25) index units (the roundings of the two CDF entries, of the subtraction and of the re-wrap).  So when
// phi and 1 - phi exceed n 2^-20 -- four times that bound -- every r_cdf[t] lies STRICTLY between u_cdf[t-1] and u_cdf[t]:
//     j0 = floor(F),   #{u_cdf < r_cdf[t]} = #{u_cdf^+ <= r_cdf[t]} = t,   #{r_cdf < u_cdf[k]} = k + 1,
// the merged axis of Cost is r_0 u_0 r_1 u_1 ... r_{n-1} u_{n-1}, and the searches of the generic code are known without being
// run -- same summands, same order, same bits (the A/B test switches this off with shwd_circular_wp_set_dyadic(0)).  This is
// what the rounds of a bisection look like once the bracket is narrower than 1/n, and almost every earlier round of sizes that
// are not a power of two; rotations within the margin of the grid (the first rounds; n = 1000: multiples of 1/8) run the
// generic searches.
__device__ __forceinline__ bool between_safe(const Circle& c, const Shift& s) {
  if (s.allneg) return false;
  const double F = (double)s.frac / (double)c.wu;
  const double fl = floor(F), phi = F - fl;
  const double margin = (double)c.n * 9.5367431640625e-07;  // n 2^-20
  return phi > margin && 1.0 - phi > margin && (int)fl == s.j0;
}

template <bool P2, int T>
__device__ float cost_pass_u_between(const Circle& c, float theta, float* __restrict__ gu, const int32_t* __restrict__ pu, float2* wtot) {
  const Shift s = make_shift<false>(c, theta);
  float acc = 0.f;
  for (int k = threadIdx.x; k < c.n; k += T) {
    const float prev = (k > 0) ? ucdf_at<false>(c, k - 1) : 0.f;
    const float xk = ucdf_at<false>(c, k);
    const float U = c.u[k];
    const float rk = r_cdf<false>(c, s, k);  // the one v entry inside (u_cdf[k-1], u_cdf[k]): v-index k, u-index k
    const float d1 = __fsub_rn(U, r_val_in<false>(c, s, k)), e1 = __fsub_rn(rk, prev);
    acc = fmaf(e1, powp<P2>(d1, c.p), acc);
    float g = fmaf(e1, dpowp<P2>(d1, c.p), 0.f);
    const float d2 = __fsub_rn(U, r_val<false>(c, s, k + 1)), e2 = __fsub_rn(xk, rk);  // the u entry: v-index k + 1 (m: the wrap entry)
    acc = fmaf(e2, powp<P2>(d2, c.p), acc);
    g = fmaf(e2, dpowp<P2>(d2, c.p), g);
    if (gu) gu[pu ? __ldg(pu + k) : k] = g;
  }
  return block_sum2<T>(acc, 0.f, wtot).x;
}

template <bool P2, int T>
__device__ void cost_pass_v_between(const Circle& c, float theta, float* __restrict__ gv, const int32_t* __restrict__ pv) {
  const Shift s = make_shift<false>(c, theta);
  for (int t = threadIdx.x; t < c.m; t += T) {
    float prev = (t > 0) ? r_cdf<false>(c, s, t - 1) : 0.f;
    const float xt = r_cdf<false>(c, s, t);
    const float V = r_val_in<false>(c, s, t);
    float g = 0.f;
    if (t > 0) {  // the one u entry inside (r[t-1], r[t]): u_cdf[t-1], v-index t
      const float a = ucdf_at<false>(c, t - 1);
      g = fmaf(__fsub_rn(a, prev), dpowp<P2>(__fsub_rn(c.u[t - 1], V), c.p), g);
      prev = a;
    }
    g = fmaf(__fsub_rn(xt, prev), dpowp<P2>(__fsub_rn(c.u[min(t, c.n - 1)], V), c.p), g);  // the v entry: u-index t
    if (t == 0) {  // the wrap-around partner r_val[m] = r_val[0] + 1 serves the one u entry above r_cdf[m-1]: u_cdf[n-1]
      const float pw = r_cdf<false>(c, s, c.m - 1);
      g = fmaf(__fsub_rn(ucdf_at<false>(c, c.n - 1), pw), dpowp<P2>(__fsub_rn(c.u[c.n - 1], __fadd_rn(V, 1.f)), c.p), g);
    }
    int j = t + s.j0;
    if (j >= c.m) j -= c.m;
    gv[pv ? __ldg(pv + j) : j] = -g;
  }
}

// theta on the 1/n grid (n a power of two: the product is exact)
__device__ __forceinline__ bool on_grid(float theta, int n) {
  const float fr = __fsub_rn(theta, floorf(theta));  // (a rotation in [-2^-25, 0) rounds to 1.0: not a grid point)
  const float fq = fr * (float)n;
  return fr < 1.f && fq == floorf(fq);
}

template <bool P2, int T, bool MEMO, bool W = false>
__global__ void __launch_bounds__(T) circular_wp_kernel(const float* __restrict__ us, const float* __restrict__ vs,
                                                                 const int32_t* __restrict__ pu, const int32_t* __restrict__ pv, int n,
                                                                 int m, float p, float tm0, float tp0, float tol,
                                                                 float* __restrict__ w_out,
                                                                 float* __restrict__ gus, float* __restrict__ gvs,
                                                                 float* __restrict__ theta_out,
                                                                 const float* __restrict__ ucdfs = nullptr,
                                                                 const float* __restrict__ vcdfs = nullptr,
                                                                 float* __restrict__ gcus = nullptr, float* __restrict__ gcvs = nullptr,
                                                                 int dyadic = 0) {
  extern __shared__ float cw_smem[];
  __shared__ float2 wtot[T / 32];
  __shared__ int wskip[T / 32];
  float* su = cw_smem;
  float* sv = su + n;
  const size_t sl = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += T) {
    su[i] = __ldg(us + sl * n + i);
  }
  for (int j = threadIdx.x; j < m; j += T) {
    sv[j] = __ldg(vs + sl * m + j);
  }
  float* scu = sv + m;
  float* scv = scu + n;
  if (W) {
    for (int i = threadIdx.x; i < n; i += T) scu[i] = __ldg(ucdfs + sl * n + i);
    for (int j = threadIdx.x; j < m; j += T) scv[j] = __ldg(vcdfs + sl * m + j);
  }
  __syncthreads();
  Circle c = {su, sv, n, m, gridDim.x == 1, p, (float)(1.0 / (double)n), (float)(1.0 / (double)m), W ? scu : nullptr,
              W ? scv : nullptr};

  // binary_search_circle :172-205 (every quantity is uniform over the CTA: the sums are broadcast)
  float tm = tm0, tp = tp0, tc = (tm0 + tp0) * 0.5f;
  DcMemo at_tm, at_tp, at_tc;  // dCost at the bracket ends (invalid until an end has been a midpoint) and at tc
  at_tm.j0 = -1;
  at_tp.j0 = -1;
  bool searches_first = false;
  // closed-form searches (uniform weights, n == m; `dyadic` bit 0: equal power-of-two sizes while the rotation sits on the 1/n
  // grid, bit 1: any equal sizes while it is safely off the grid): mode 1 / 2, else 0 = the generic searches
  const bool dy = !W && (dyadic & 1) != 0, bt = !W && (dyadic & 2) != 0;
  auto mode_of = [&](float th) -> int {
    if (dy && on_grid(th, n)) return 1;
    if (bt && between_safe(c, make_shift<false>(c, th))) return 2;
    return 0;
  };
  auto eval = [&](float th, DcMemo& out) -> float2 {
    const int md = mode_of(th);
    if (md != 0) {
      // (fills the memo of the generic rounds; a generic round that follows tries the searches first: the bracket is narrow by then)
      if (MEMO && md == 2) searches_first = true;
      return dcost_dyadic<P2, T>(c, th, wtot, md == 2, MEMO ? &out : nullptr);
    }
    return dcost<P2, T, MEMO, W>(c, th, wtot, wskip, at_tm, at_tp, out, searches_first);
  };
  auto cost_at = [&](float th) -> float {
    const int md = mode_of(th);
    if (md == 1) return cost_pass_u_dyadic<P2, T>(c, th, nullptr, nullptr, wtot);
    if (md == 2) return cost_pass_u_between<P2, T>(c, th, nullptr, nullptr, wtot);
    return cost_pass_u<P2, T, W>(c, th, nullptr, nullptr, wtot);
  };
  for (int round = 0; round < CW_MAX_ROUNDS; ++round) {
    const float2 dc = eval(tc, at_tc);
    if (dc.x * dc.y <= 0.f) break;  // done: the optimum is the kink at tc
    if (__fsub_rn(tp, tm) < tol) {
      DcMemo unused;
      const float2 dtp = eval(tp, unused);
      const float2 dtm = eval(tm, unused);
      const float ctm = cost_at(tm);
      const float ctp = cost_at(tp);
      const float den = __fsub_rn(dtm.x, dtp.y);  // dCptm - dCmtp
      if (fabsf(den) > 0.001f)
        tc = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(ctp, ctm), __fmul_rn(tm, dtm.x)), __fmul_rn(tp, dtp.y)), den);
      break;
    }
    if (dc.x < 0.f) { tm = tc; if (MEMO) at_tm = at_tc; } else { tp = tc; if (MEMO) at_tp = at_tc; }
    tc = __fmul_rn(__fadd_rn(tm, tp), 0.5f);
  }
  float w;
  const int mdf = mode_of(tc);
  if (mdf == 1) {
    w = cost_pass_u_dyadic<P2, T>(c, tc, gus ? gus + sl * n : nullptr, pu ? pu + sl * n : nullptr, wtot);
    if (gvs) cost_pass_v_dyadic<P2, T>(c, tc, gvs + sl * m, pv ? pv + sl * m : nullptr);
  } else if (mdf == 2) {
    w = cost_pass_u_between<P2, T>(c, tc, gus ? gus + sl * n : nullptr, pu ? pu + sl * n : nullptr, wtot);
    if (gvs) cost_pass_v_between<P2, T>(c, tc, gvs + sl * m, pv ? pv + sl * m : nullptr);
  } else {
    w = cost_pass_u<P2, T, W>(c, tc, gus ? gus + sl * n : nullptr, pu ? pu + sl * n : nullptr, wtot,
                              (W && gcus) ? gcus + sl * n : nullptr);
    if (gvs) cost_pass_v<P2, T, W>(c, tc, gvs + sl * m, pv ? pv + sl * m : nullptr, (W && gcvs) ? gcvs + sl * m : nullptr);
  }
  if (threadIdx.x == 0) {
    w_out[sl] = w;
    if (theta_out) theta_out[sl] = tc;
  }
}

}  // namespace shwd

using namespace shwd;

extern "C" size_t shwd_circular_wp_workspace_bytes(int S, int n, int m) {
  (void)S;
  (void)n;
  (void)m;
  return 0;  // the uniform CDFs are evaluated in registers (ucdf_at / vcdf_at); kept for ABI stability
}

static int g_wp_dyadic = 1;  // shwd_circular_wp_set_dyadic
extern "C" int shwd_circular_wp_set_dyadic(int on) {
  g_wp_dyadic = on ? 1 : 0;
  return SHWD_OK;
}

static int circular_wp_dispatch(const float* us, const float* vs, const int32_t* pu, const int32_t* pv, int S, int n, int m, float p,
                                float tm, float tp, float tol, float* w, float* gus, float* gvs, float* theta, void* workspace,
                                size_t workspace_bytes, void* stream, const float* ucdf = nullptr, const float* vcdf = nullptr,
                                float* gcu = nullptr, float* gcv = nullptr) {
  if (!us || !vs || !w || S < 0 || n <= 0 || m <= 0 || !(p > 0.f) || !(tm < tp)) return SHWD_ERR_INVALID_ARGUMENT;
  if (S == 0) return SHWD_OK;
  (void)workspace;
  (void)workspace_bytes;
  const bool weighted = ucdf != nullptr;
  const size_t smem = ((size_t)n + m) * sizeof(float) * (weighted ? 2 : 1);
  if (n > 32768 || m > 32768) return SHWD_ERR_UNSUPPORTED;  // exactness argument of ucdf_at: at most 2^15 weights per CDF
  if (smem > 220 * 1024) return SHWD_ERR_UNSUPPORTED;        // n + m <= 56320 per slice
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // Rows beyond 14336 merged entries leave room for at most three 256-thread CTAs per SM (one at cfg4's 32768): those
  // slices run 1024 threads per CTA instead (cfg4 SSW p=2: see DESIGN.md 4.4); shorter rows keep the 256-thread CTAs
  // (and their summation order).
  const bool big = smem > 56 * 1024;
  // equal power-of-two sizes: the bisection ends on a kink at round log2(n), nothing to remember (see DcMemo)
  const bool memo = !(n == m && (n & (n - 1)) == 0) && !weighted;
  // closed-form searches: bit 0 on the 1/n grid (equal power-of-two sizes), bit 1 safely off it (any equal sizes)
  const int wp_shortcuts = (!weighted && g_wp_dyadic && n == m) ? (((n & (n - 1)) == 0 ? 1 : 0) | 2) : 0;
#define SHWD_LAUNCH_WP_M(P2, T, MEMO)                                                                                              \
  do {                                                                                                                             \
    if (smem > 32 * 1024) /* static + dynamic beyond 48 KB needs the opt-in (static is < 16 KB here) */                            \
      SHWD_CUDA_CHECK(                                                                                                             \
          cudaFuncSetAttribute(circular_wp_kernel<P2, T, MEMO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    circular_wp_kernel<P2, T, MEMO><<<S, T, smem, s>>>(us, vs, pu, pv, n, m, p, tm, tp, tol, w, gus, gvs, theta, nullptr, nullptr, \
                                                       nullptr, nullptr, wp_shortcuts);                                            \
  } while (0)
#define SHWD_LAUNCH_WP_W(P2, T)                                                                                                    \
  do {                                                                                                                             \
    if (smem > 32 * 1024)                                                                                                          \
      SHWD_CUDA_CHECK(cudaFuncSetAttribute(circular_wp_kernel<P2, T, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                           (int)smem));                                                                           \
    circular_wp_kernel<P2, T, false, true><<<S, T, smem, s>>>(us, vs, pu, pv, n, m, p, tm, tp, tol, w, gus, gvs, theta, ucdf,     \
                                                               vcdf, gcu, gcv);                                                   \
  } while (0)
#define SHWD_LAUNCH_WP(P2, T)                                                                                                      \
  do {                                                                                                                             \
    if (weighted) SHWD_LAUNCH_WP_W(P2, T); else if (memo) SHWD_LAUNCH_WP_M(P2, T, true); else SHWD_LAUNCH_WP_M(P2, T, false);      \
  } while (0)
  if (p == 2.f) {
    if (big) SHWD_LAUNCH_WP(true, 1024); else SHWD_LAUNCH_WP(true, CW_THREADS);
  } else {
    if (big) SHWD_LAUNCH_WP(false, 1024); else SHWD_LAUNCH_WP(false, CW_THREADS);
  }
#undef SHWD_LAUNCH_WP
#undef SHWD_LAUNCH_WP_W
#undef SHWD_LAUNCH_WP_M
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_circular_wp(const float* us, const float* vs, int S, int n, int m, float p, float tm, float tp, float tol,
                                float* w, float* gus, float* gvs, float* theta, void* workspace, size_t workspace_bytes,
                                void* stream) {
  return circular_wp_dispatch(us, vs, nullptr, nullptr, S, n, m, p, tm, tp, tol, w, gus, gvs, theta, workspace, workspace_bytes,
                              stream);
}

extern "C" int shwd_circular_wp_scatter(const float* us, const float* vs, const int32_t* perm_u, const int32_t* perm_v, int S, int n,
                                        int m, float p, float tm, float tp, float tol, float* w, float* gku, float* gkv,
                                        float* theta, void* workspace, size_t workspace_bytes, void* stream) {
  if (!perm_u || !perm_v || !gku || !gkv) return SHWD_ERR_INVALID_ARGUMENT;
  return circular_wp_dispatch(us, vs, perm_u, perm_v, S, n, m, p, tm, tp, tol, w, gku, gkv, theta, workspace, workspace_bytes,
                              stream);
}

// binary_search_circle with u_weights / v_weights (max_spherical_sliced_w.py:156-170): ucdf (S,n) / vcdf (S,m) are the per-slice
// CDF tables cumsum(weights[..., sorter], -1) the reference forms; gcu / gcv (nullable) receive d W / d ucdf, d W / d vcdf (the
// reference's autograd reaches the weights through the merged CDF axis of the final Cost, :93-95).  n + m <= 28160.
extern "C" int shwd_circular_wp_weighted(const float* us, const float* vs, const float* ucdf, const float* vcdf, int S, int n, int m,
                                         float p, float tm, float tp, float tol, float* w, float* gus, float* gvs, float* gcu,
                                         float* gcv, float* theta, void* stream) {
  if (!ucdf || !vcdf) return SHWD_ERR_INVALID_ARGUMENT;
  if ((gcu || gcv) && !(gus && gvs)) return SHWD_ERR_INVALID_ARGUMENT;  // the CDF gradients ride on the gradient passes
  return circular_wp_dispatch(us, vs, nullptr, nullptr, S, n, m, p, tm, tp, tol, w, gus, gvs, theta, nullptr, 0, stream, ucdf, vcdf,
                              gcu, gcv);
}

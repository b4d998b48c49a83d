// Exact optimal assignment between two equally sized clouds on an on-the-fly cost: the GPU replacement
// for the per-pair exact solve the reference's W_COS path really runs,
//     ot.emd2(a_i, b_i, C_i)        Point_Cloud_Resistration/losses/s2_wasserstein.py:39-50, 99-110
//                                   Comparison_Wasserstein_with_Chamfer_distance/main_rotation.py:63-79
// (POT's C++ network simplex on a float64 copy of C, one pair at a time on the CPU, with a GPU->CPU->GPU round trip).
// For uniform weights and n == m the LP optimum is attained at a permutation, so emd2 = (1/n) * min-assignment cost.
// POT is a third-party dependency that is not vendored (version un-pinned): this solves the same LP by a different exact
// method, Bertsekas' forward auction with epsilon-scaling, in 64-bit fixed point on the float32 cost values (every cost is
// scaled by a power of two S with Cmax * S in [2^45, 2^46) and converted once -- exact for every cost >= Cmax * 2^-23, to
// 2^-46 Cmax below that -- so prices, bids and comparisons are integer adds and compares: the first version ran them in
// float64, whose dependent-op latency was what a bid and a scan spent their time on):
//   * an unassigned person i bids for its best object j1 = argmin_j w_ij, w_ij = C_ij + price_j, the price
//     price_j1 + (second best w - best w) + eps; the previous owner becomes unassigned; a phase ends when everybody is
//     assigned;
//   * eps starts at Cmax/100 and is divided by 5 per phase down to Cmax * 2^-40, prices are kept between phases.  At the
//     end the assignment is within n*eps of optimal -- far below one float32 ulp of the cost sum -- i.e. it is the
//     optimum unless two assignments tie to ~1e-9 relative.
//
// What the first version of this kernel measured (one CTA per pair, every bidder scanning all N objects every round): a
// 1024-point pair needs ~9 k rounds, nearly all of them "price wars" with one or two bidders, ~5 us each (five barriers, a
// full scan) -> 46 ms per pair, latency-bound.  This version removes the scan and the barriers from those bids:
//   * CANDIDATE LISTS.  Prices only rise (within and across phases), so w_ij only grows.  A full scan of person i stores
//     the K objects with the smallest w (index + float32 cost, in shared memory) and Tw_i = the smallest w it did NOT
//     store.  Later, as long as the second-smallest CURRENT w among the K listed objects is <= Tw_i, the list provably
//     contains the person's best and second-best object overall, and a bid needs K shared-memory loads instead of N
//     cost evaluations.  Lists are selected by a window [w1, max(w1 + delta_i, w2)] with a per-person adaptive delta_i.
//   * GAUSS-SEIDEL BIDS IN ONE WARP.  Warp 0 pops unassigned persons from a ring buffer and serves them from their lists
//     with warp-wide integer min-reductions (redux.sync on the two halves of the 64-bit w) -- no
//     __syncthreads, ~0.15 us per bid.  A person whose list is exhausted is deferred to a rescan list R.
//   * BATCHED RESCANS + ONE JACOBI ROUND.  When the queue is empty all 16 warps rescan the persons in R in parallel (one
//     warp each: best / second best, new list, new Tw) against the same price vector, and those persons bid at once
//     (a Jacobi round: one atomicMax per bid on (bid bits | slot), the winner takes the object).  Phases also start with
//     one such parallel round served from the lists.  Large-eps phases (every bid moves a price by more than a list's
//     window) therefore run as parallel rounds, small-eps price wars as barrier-free list bids.
// Any unassigned person may bid in any order and any bidder may win an object (only eps-complementary slackness of the
// assigned pairs and monotone prices are needed), so the result is exact in the same sense as before; everything is
// deterministic (fixed orders, the only atomics are max on keys that embed the slot index).
// Output: sigma (B, N) int32 with sigma[i] = object of person i.  The loss value and its gradient are assembled by the
// caller from the n matched pairs (the plan has n non-zeros, d emd2 / dC = plan -- what POT attaches for autograd).
#include "common.cuh"
#include "cost.cuh"
#include <limits.h>

namespace shwd {

constexpr int AU_THREADS = 512;
constexpr int AU_WARPS = AU_THREADS / 32;
#ifndef SHWD_AU_EPS_FACTOR
#define SHWD_AU_EPS_FACTOR 0.2
#endif
constexpr double AU_EPS_FACTOR = SHWD_AU_EPS_FACTOR;  // epsilon is multiplied by this between phases
constexpr long long AU_MAX_BIDS = 64ll << 20;  // safety net (a solve needs ~50 n bids in practice)
constexpr int AU_RC = 512;        // rescanned persons per Jacobi round (their slot index rides in the low 9 bits of the bid key)
constexpr int AU_MAX_N = 2048;    // person index rides in the low 11 bits of the phase-start bid key
constexpr int AU_PP = AU_MAX_N / AU_THREADS;  // persons per thread in the phase-start round
#ifndef SHWD_AU_EPS0
#define SHWD_AU_EPS0 0.01  // first epsilon as a fraction of Cmax (A/B on the training shape: 0.25 22 ms, 0.05 18, 0.01 16, 0.005 20)
#endif
#ifndef SHWD_AU_LIST_EPS
#define SHWD_AU_LIST_EPS 2.5e-3  // candidate lists are built once eps <= this fraction of Cmax
#endif
#ifndef SHWD_AU_DELTA0
#define SHWD_AU_DELTA0 8.0
#endif
// FAST value of the dense variant: the cost is READ from a caller-supplied (B, N, N) matrix instead of being evaluated from
// points -- the drop-in for ot.emd2(a, b, M) called on an explicit cost matrix (main_rotation.py:63-79 POT_loss; notebooks).
constexpr int AU_DENSE = 1000;
#ifdef SHWD_AU_PROFILE  // diagnostics build (tools/build_variant.sh): cycle counts per stage land in price_out[0..7]
#define AU_PROF_T(v) const long long v = clock64()
#define AU_PROF_ADD(acc, t0) acc += clock64() - t0
#else
#define AU_PROF_T(v)
#define AU_PROF_ADD(acc, t0)
#endif

template <int FAST>
__device__ __forceinline__ float au_cost(const CostParams& cp, const float4 o, const float4* sY, const float* crow, int j) {
  if constexpr (FAST == AU_DENSE) {
    return __ldg(crow + j);
  } else {
    typedef Cost<FAST> CF;
    const float4 t = sY[j];
    return CF::kc(cp, CF::eval(cp, o.x, o.y, o.z, t.x, t.y, t.z));
  }
}

typedef long long i64;
constexpr i64 AU_INF = 0x7FFFFFFFFFFFFFFFll;   // "no entry"
constexpr i64 AU_NONE = -AU_INF - 1;            // Tw of a person without a list
struct AuFix {  // float32 cost -> fixed point: one multiply by a power of two (exact) and one round-to-nearest conversion
  float scale;
  __device__ __forceinline__ i64 operator()(float c) const { return __float2ll_rn(c * scale); }
};
struct AuBest {  // smallest and second-smallest w of a scan, the object and plain cost of the smallest
  i64 w1, w2;
  int j1;
  float c1;
};
__device__ __forceinline__ AuBest au_merge(const AuBest& a, const AuBest& b) {
  AuBest r;
  const bool takea = (a.w1 < b.w1) || (a.w1 == b.w1 && a.j1 <= b.j1);
  if (takea) {
    r.w1 = a.w1;
    r.j1 = a.j1;
    r.c1 = a.c1;
    r.w2 = min(a.w2, b.w1);
  } else {
    r.w1 = b.w1;
    r.j1 = b.j1;
    r.c1 = b.c1;
    r.w2 = min(b.w2, a.w1);
  }
  return r;
}

struct AuShared {
  float4* sY;                  // N   packed points of the object cloud (unused by the dense variant)
  i64* price;                  // N
  unsigned long long* bidval;  // N   highest (bid bits | slot) per object in a Jacobi round (0: none)
  i64* Tw;                     // N   smallest w the person's last scan did NOT list (AU_NONE: no list yet)
  float* delta;                // N   the person's list window
  int* owner;                  // N   person holding object j (-1: free)
  int* queue;                  // N   ring buffer of unassigned persons with a (possibly) usable list
  int* R;                      // N   persons whose list is exhausted: to be rescanned
  unsigned short* lidx;        // N*K listed objects (0xFFFF: empty slot)
  float* lcost;                // N*K their plain costs
  i64* rbid;                   // AU_RC  bid of a rescanned person
  int* rj;                     // AU_RC  its object
};

// ---- full scans of person i by one warp against the current prices (four objects per lane in flight, branch-free) ----
__device__ __forceinline__ AuBest au_warp_merge(AuBest best) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    AuBest other;
    other.w1 = __shfl_xor_sync(0xffffffffu, best.w1, s);
    other.w2 = __shfl_xor_sync(0xffffffffu, best.w2, s);
    other.j1 = __shfl_xor_sync(0xffffffffu, best.j1, s);
    other.c1 = __shfl_xor_sync(0xffffffffu, best.c1, s);
    best = au_merge(best, other);
  }
  if (best.j1 == INT_MAX) best.j1 = 0;  // (cannot happen for finite costs; keeps indices in range)
  return best;
}
__device__ __forceinline__ void au_top2_update(AuBest& best, i64 w, int j, float c) {
  const bool lt = w < best.w1;
  best.w2 = min(best.w2, lt ? best.w1 : w);
  best.j1 = lt ? j : best.j1;
  best.c1 = lt ? c : best.c1;
  best.w1 = lt ? w : best.w1;
}
// best and second best only (the large-eps phases: every bid moves a price by more than a list window, lists are useless)
template <int FAST>
__device__ __forceinline__ AuBest au_scan_top2(const CostParams& cp, const AuShared& S, const float4 o, const float* crow,
                                               AuFix fix, int N, int lane) {
  // four independent (best, second best) accumulators: the update is a dependent compare / select chain
  AuBest acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = {AU_INF, AU_INF, INT_MAX, 0.f};
  for (int j0 = lane; j0 < N; j0 += 128) {
    float c[4];
    i64 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u;
      const bool ok = j < N;
      c[u] = ok ? au_cost<FAST>(cp, o, S.sY, crow, ok ? j : 0) : 0.f;
      p[u] = ok ? S.price[j] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u;
      au_top2_update(acc[u], j < N ? fix(c[u]) + p[u] : AU_INF, j, c[u]);
    }
  }
  return au_warp_merge(au_merge(au_merge(acc[0], acc[1]), au_merge(acc[2], acc[3])));
}
// One collecting pass: every object with w <= cut goes into the person's list in index order (at most K; j1k >= 0: that
// object is known to be the best, it is skipped here and owns slot 0), tmin = smallest w NOT stored, cnt = how many
// qualified; TRACK: also the exact best / second best over all objects (valid whenever cut >= the true second best).
template <int FAST, bool TRACK>
__device__ __forceinline__ void au_scan_range(const CostParams& cp, const AuShared& S, const float4 o, const float* crow,
                                              AuFix fix, int lo, int N, int K, unsigned short* li, float* lc, int lane,
                                              i64 cut, int j1k, int cnt0, int& cnt, i64& tmin, AuBest& best) {
  // objects lo <= j < N; entries go to li / lc (capacity K), the first one at position cnt0
  const unsigned lt = (1u << lane) - 1u;
  cnt = cnt0;
  tmin = AU_INF;
  for (int j0 = lo + lane; j0 < N + lane; j0 += 128) {  // (+ lane: every lane runs the same number of ballots)
    float c[4];
    i64 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u;
      const bool ok = j < N;
      c[u] = ok ? au_cost<FAST>(cp, o, S.sY, crow, ok ? j : 0) : 0.f;
      p[u] = ok ? S.price[j] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u;
      const i64 w = j < N ? fix(c[u]) + p[u] : AU_INF;
      if (TRACK) au_top2_update(best, w, j, c[u]);
      const bool other = j < N && j != j1k;
      const bool take = other && w <= cut;
      const unsigned b = __ballot_sync(0xffffffffu, take);
      const int pos = cnt + __popc(b & lt);
      const bool st = take && pos < K;
      if (st) {
        li[pos] = (unsigned short)j;
        lc[pos] = c[u];
      }
      tmin = (other && !st) ? min(tmin, w) : tmin;
      cnt += __popc(b);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) tmin = min(tmin, __shfl_xor_sync(0xffffffffu, tmin, s));
}
template <int FAST, bool TRACK>
__device__ __forceinline__ void au_scan_collect(const CostParams& cp, const AuShared& S, const float4 o, const float* crow,
                                                AuFix fix, int N, int K, int i, int lane, i64 cut, int j1k, int& cnt,
                                                i64& tmin, AuBest& best) {
  au_scan_range<FAST, TRACK>(cp, S, o, crow, fix, 0, N, K, S.lidx + (size_t)i * K, S.lcost + (size_t)i * K, lane, cut, j1k,
                             j1k >= 0 ? 1 : 0, cnt, tmin, best);
}
// A rescan that also rebuilds the candidate list.  guess: an upper bound of the person's true second-best w (the second
// best of its exhausted list), or non-finite when there is none: with a guess the list is cut at guess + delta_i and ONE
// pass finds the exact best / second best and the list; otherwise (or when more than K objects qualify) the cut comes from
// the exact values: max(w1 + delta_i, w2), then w2 itself.
template <int FAST>
__device__ __forceinline__ AuBest au_rescan(const CostParams& cp, const AuShared& S, const float4 o, const float* crow,
                                            AuFix fix, int N, int K, int i, int lane) {
  float dl = S.delta[i];
  const i64 guess = S.Tw[i];
  AuBest best = {AU_INF, AU_INF, INT_MAX, 0.f};
  int cnt = 0;
  i64 tmin = AU_INF;
  bool done = false;
  if (guess > AU_NONE && guess < AU_INF) {
    const i64 cut = guess + (i64)dl;
    au_scan_collect<FAST, true>(cp, S, o, crow, fix, N, K, i, lane, cut, -1, cnt, tmin, best);
    best = au_warp_merge(best);
    done = cnt <= K && best.w2 <= cut;  // (prices may have risen past the guess since it was taken)
    if (cnt > K) dl *= 0.125f;
  } else {
    best = au_scan_top2<FAST>(cp, S, o, crow, fix, N, lane);
  }
  unsigned short* li = S.lidx + (size_t)i * K;
  if (!done) {
    AuBest dummy = best;
    au_scan_collect<FAST, false>(cp, S, o, crow, fix, N, K, i, lane, max(best.w1 + (i64)dl, best.w2), best.j1, cnt, tmin, dummy);
    if (cnt > K) {
      dl *= 0.125f;
      au_scan_collect<FAST, false>(cp, S, o, crow, fix, N, K, i, lane, best.w2, best.j1, cnt, tmin, dummy);
    }
    if (lane == 0) {
      li[0] = (unsigned short)best.j1;
      S.lcost[(size_t)i * K] = best.c1;
    }
  }
  if (cnt <= K / 2) dl *= 2.f;
  for (int k = cnt + lane; k < K; k += 32) li[k] = 0xFFFFu;
  if (lane == 0) {
    S.Tw[i] = tmin;
    S.delta[i] = dl;
  }
  return best;
}

// The same rescan by G = 2, 4, 8 or 16 warps per person (a Jacobi round with few persons -- the price wars of the small-eps
// phases -- would otherwise leave most of the CTA idle behind one warp's 1024-object scan).  Warp `sub` of a group scans a
// contiguous N / G slice into its own staging row; the per-warp (count, smallest unstored w, best / second best) meet in
// shared memory, and every warp copies its staged entries to the list at the offset its predecessors' counts give (index
// order again).  Up to three attempts with the cuts of au_rescan; every attempt is one scan + one __syncthreads, and every
// warp of the CTA runs the same number of barriers.  build == false: best / second best only (one attempt, no list).
struct AuPart {
  AuBest best;
  i64 tmin;
  int cnt;
};
constexpr int AU_STAGE_K = 16;
template <int FAST>
__device__ __forceinline__ AuBest au_rescan_coop(const CostParams& cp, const AuShared& S, const float4* Xb, const float* Cb,
                                                 AuFix fix, int N, int K, int c0, int nC, int G, bool build,
                                                 AuPart (*part)[AU_WARPS], unsigned short (*stg_i)[AU_STAGE_K],
                                                 float (*stg_c)[AU_STAGE_K], int warp, int lane) {
  constexpr bool DENSE = FAST == AU_DENSE;
  const int r = warp / G, sub = warp % G;
  const bool active = r < nC;
  const int i = active ? S.R[c0 + r] : 0;
  const float4 o = (DENSE || !active) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(Xb + i);
  const float* crow = DENSE ? Cb + (size_t)i * N : nullptr;
  const int per = (N + G - 1) / G;
  const int lo = min(sub * per, N), hi = min(lo + per, N);
  float dl = active ? S.delta[i] : 0.f;
  const i64 guess = active ? S.Tw[i] : 0;
  const bool has_guess = build && guess > AU_NONE && guess < AU_INF;
  AuBest best = {AU_INF, AU_INF, INT_MAX, 0.f};
  bool done = !active;
  const int natt = build ? 3 : 1;
  for (int a = 0; a < natt; ++a) {
    i64 cut = AU_NONE;
    int j1k = -1;
    if (!done) {
      if (a == 0) {
        if (has_guess) cut = guess + (i64)dl;
      } else {
        cut = (a == 1) ? max(best.w1 + (i64)dl, best.w2) : best.w2;
        j1k = best.j1;
      }
      int cnt;
      i64 tmin;
      AuBest bp = {AU_INF, AU_INF, INT_MAX, 0.f};
      if (a == 0)
        au_scan_range<FAST, true>(cp, S, o, crow, fix, lo, hi, K, stg_i[warp], stg_c[warp], lane, cut, j1k, 0, cnt, tmin, bp);
      else
        au_scan_range<FAST, false>(cp, S, o, crow, fix, lo, hi, K, stg_i[warp], stg_c[warp], lane, cut, j1k, 0, cnt, tmin, bp);
      if (a == 0) {  // (au_warp_merge without the index fix-up: an empty slice keeps INT_MAX and loses every merge)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
          AuBest other;
          other.w1 = __shfl_xor_sync(0xffffffffu, bp.w1, s);
          other.w2 = __shfl_xor_sync(0xffffffffu, bp.w2, s);
          other.j1 = __shfl_xor_sync(0xffffffffu, bp.j1, s);
          other.c1 = __shfl_xor_sync(0xffffffffu, bp.c1, s);
          bp = au_merge(bp, other);
        }
      }
      if (lane == 0) {
        part[a][warp].best = bp;
        part[a][warp].tmin = tmin;
        part[a][warp].cnt = cnt;
      }
    }
    // (the barrier doubles as the vote "every group of the CTA is done": the remaining attempts are skipped by everybody)
    if (__syncthreads_and(done ? 1 : 0)) break;
    if (!done) {
      // the G partials of this group, one per lane: totals by warp reductions, the (best, second best) by a butterfly
      const int base = j1k >= 0 ? 1 : 0;
      AuPart q = {{AU_INF, AU_INF, INT_MAX, 0.f}, AU_INF, 0};
      if (lane < G) q = part[a][r * G + lane];
      int tot = q.cnt, before = lane < sub ? q.cnt : 0;
      i64 tminG = q.tmin;
#pragma unroll
      for (int sft = 8; sft > 0; sft >>= 1) {  // G <= 16 lanes hold data, the others the neutral element
        tot += __shfl_xor_sync(0xffffffffu, tot, sft);
        before += __shfl_xor_sync(0xffffffffu, before, sft);
        tminG = min(tminG, __shfl_xor_sync(0xffffffffu, tminG, sft));
      }
      tot += __shfl_xor_sync(0xffffffffu, tot, 16);
      before += __shfl_xor_sync(0xffffffffu, before, 16);
      tminG = min(tminG, __shfl_xor_sync(0xffffffffu, tminG, 16));
      tot += base;
      const int off = base + before;
      const int mine = __shfl_sync(0xffffffffu, q.cnt, sub);
      if (a == 0) {
        AuBest bb = q.best;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
          AuBest other;
          other.w1 = __shfl_xor_sync(0xffffffffu, bb.w1, sft);
          other.w2 = __shfl_xor_sync(0xffffffffu, bb.w2, sft);
          other.j1 = __shfl_xor_sync(0xffffffffu, bb.j1, sft);
          other.c1 = __shfl_xor_sync(0xffffffffu, bb.c1, sft);
          bb = au_merge(bb, other);
        }
        best = bb;
      }
      if (a == 0 && best.j1 == INT_MAX) best.j1 = 0;
      bool ok;
      if (a == 0) {
        ok = has_guess && tot <= K && best.w2 <= cut;  // (prices may have risen past the guess since it was taken)
        if (has_guess && tot > K) dl *= 0.125f;
      } else if (a == 1) {
        ok = tot <= K;
        if (!ok) dl *= 0.125f;
      } else {
        ok = true;
      }
      if (ok && build) {
        unsigned short* li = S.lidx + (size_t)i * K;
        float* lc = S.lcost + (size_t)i * K;
        for (int e = lane; e < min(mine, K); e += 32)
          if (off + e < K) {
            li[off + e] = stg_i[warp][e];
            lc[off + e] = stg_c[warp][e];
          }
        if (sub == 0) {
          for (int k = tot + lane; k < K; k += 32) li[k] = 0xFFFFu;
          if (lane == 0) {
            if (j1k >= 0) {
              li[0] = (unsigned short)best.j1;
              lc[0] = best.c1;
            }
            S.Tw[i] = tot > K ? min(tminG, best.w2) : tminG;  // (entries cut off by the capacity tie with the second best)
            S.delta[i] = tot <= K / 2 ? dl * 2.f : dl;
          }
        }
        done = true;
      }
    }
  }
  return best;
}

// Smallest and second-smallest of one 64-bit w per lane (AU_INF: no entry), by integer min-reductions on the two halves
// of the sign-flipped value (order-preserving as unsigned).  l1 = lowest lane holding the smallest.
__device__ __forceinline__ void au_top2(i64 w, int lane, i64& w1, i64& w2, int& l1) {
  const unsigned long long kb = (unsigned long long)w ^ 0x8000000000000000ull;
  unsigned hi = (unsigned)(kb >> 32), lo = (unsigned)kb;
  const unsigned m1h = __reduce_min_sync(0xffffffffu, hi);
  const unsigned m1l = __reduce_min_sync(0xffffffffu, hi == m1h ? lo : 0xffffffffu);
  l1 = __ffs(__ballot_sync(0xffffffffu, hi == m1h && lo == m1l)) - 1;
  if (lane == l1) {
    hi = 0xffffffffu;
    lo = 0xffffffffu;
  }
  const unsigned m2h = __reduce_min_sync(0xffffffffu, hi);
  const unsigned m2l = __reduce_min_sync(0xffffffffu, hi == m2h ? lo : 0xffffffffu);
  w1 = (i64)((((unsigned long long)m1h << 32) | m1l) ^ 0x8000000000000000ull);
  w2 = (i64)((((unsigned long long)m2h << 32) | m2l) ^ 0x8000000000000000ull);  // all ones = AU_INF: no second entry
}

// The same for TWO persons at once, one per 16-lane half of the warp (K <= 16): a shuffle butterfly over xor 8, 4, 2, 1 never
// leaves a half.  Every lane of a half ends with that half's (smallest, second smallest, lane-in-half of the smallest).
__device__ __forceinline__ void au_top2_half(i64 w, int hl, i64& w1, i64& w2, int& l1) {
  unsigned long long a1 = (unsigned long long)w ^ 0x8000000000000000ull, a2 = 0xffffffffffffffffull;
  int la = hl;
#pragma unroll
  for (int s = 8; s > 0; s >>= 1) {
    const unsigned long long b1 = __shfl_xor_sync(0xffffffffu, a1, s), b2 = __shfl_xor_sync(0xffffffffu, a2, s);
    const int lb = __shfl_xor_sync(0xffffffffu, la, s);
    const bool take = b1 < a1 || (b1 == a1 && lb < la);
    a2 = take ? min(a1, b2) : min(a2, b1);
    la = take ? lb : la;
    a1 = take ? b1 : a1;
  }
  w1 = (i64)(a1 ^ 0x8000000000000000ull);
  w2 = (i64)(a2 ^ 0x8000000000000000ull);
  l1 = la;
}

template <int FAST>
__global__ void __launch_bounds__(AU_THREADS, 1) auction_kernel(const float4* __restrict__ X, const float4* __restrict__ Y,
                                                             const float* __restrict__ Cd, int N, int K, CostParams cp,
                                                             int* __restrict__ sigma, double* __restrict__ price_out,
                                                             int* __restrict__ rounds_out, int* __restrict__ status) {
  constexpr bool DENSE = FAST == AU_DENSE;
  extern __shared__ float4 au_smem[];  // carved in decreasing alignment: float4, 8-byte, 4-byte, 2-byte arrays
  AuShared S;
  S.sY = au_smem;
  S.price = reinterpret_cast<i64*>(S.sY + (DENSE ? 0 : N));
  S.bidval = reinterpret_cast<unsigned long long*>(S.price + N);
  S.Tw = reinterpret_cast<i64*>(S.bidval + N);
  S.rbid = S.Tw + N;
  S.delta = reinterpret_cast<float*>(S.rbid + AU_RC);
  S.lcost = S.delta + N;
  S.owner = reinterpret_cast<int*>(S.lcost + (size_t)N * K);
  S.queue = S.owner + N;
  S.R = S.queue + N;
  S.rj = S.R + N;
  S.lidx = reinterpret_cast<unsigned short*>(S.rj + AU_RC);
  __shared__ double s_red[AU_WARPS], s_red2[AU_WARPS];
  __shared__ AuPart s_part[3][AU_WARPS];
  __shared__ unsigned short s_stg_i[AU_WARPS][AU_STAGE_K];
  __shared__ float s_stg_c[AU_WARPS][AU_STAGE_K];
  __shared__ int s_ctl[4];  // nR, queue head, queue count, failed
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* Xb = DENSE ? nullptr : X + (size_t)b * N;
  const float* Cb = DENSE ? Cd + (size_t)b * N * N : nullptr;  // row i of this pair's matrix: Cb + i * N
  for (int i = tid; i < N; i += AU_THREADS) {
    if (!DENSE) S.sY[i] = __ldg(Y + (size_t)b * N + i);
    S.price[i] = 0;
    S.bidval[i] = 0ull;
    S.Tw[i] = AU_NONE;
    S.owner[i] = -1;
  }
  __syncthreads();
  // the cost range Cmax - min(Cmin, 0) sets the epsilon schedule and the fixed-point scale (Cmin < 0 is possible for a
  // caller-supplied matrix only; signed 64-bit arithmetic needs no shift)
  float cm = 0.f, cn = 0.f;
  bool nonfinite = false;
  for (int i = warp; i < N; i += AU_WARPS) {
    const float4 o = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(Xb + i);
    for (int j = lane; j < N; j += 32) {
      const float cij = au_cost<FAST>(cp, o, S.sY, DENSE ? Cb + (size_t)i * N : nullptr, j);
      nonfinite = nonfinite || !(fabsf(cij) <= 3.0e38f);
      cm = fmaxf(cm, cij);
      cn = fminf(cn, cij);
    }
  }
  cm = warp_max(cm);
  cn = -warp_max(-cn);
  if (lane == 0) {
    s_red[warp] = (double)cm;
    s_red2[warp] = (double)cn;
  }
  if (__syncthreads_or(nonfinite)) {
    // a NaN / infinite cost (a diverged model upstream): no assignment is meaningful.  Report failure at once;
    // sigma = identity keeps every downstream gather in range.
    for (int i = tid; i < N; i += AU_THREADS) {
      sigma[(size_t)b * N + i] = i;
      if (price_out) price_out[(size_t)b * N + i] = 0.0;
    }
    if (tid == 0) {
      if (rounds_out) rounds_out[b] = 0;
      atomicExch(status, 1);
    }
    return;
  }
  double cmax = 0.0, cshift = 0.0;
  for (int w = 0; w < AU_WARPS; ++w) {
    cmax = fmax(cmax, s_red[w]);
    cshift = fmin(cshift, s_red2[w]);
  }
  cmax -= cshift;
  if (!(cmax > 0.0)) cmax = 1.0;  // all costs zero: any assignment is optimal; run one trivial phase
  int cexp;
  frexp(cmax, &cexp);  // cmax = m * 2^cexp, m in [0.5, 1)
  const double Sd = ldexp(1.0, min(46 - cexp, 126));  // cmax * S in [2^45, 2^46): prices may rise to 2^17 Cmax before 64 bits end
  AuFix fix;
  fix.scale = (float)Sd;
  const double unit = 1.0 / Sd;
  if (N == 1) {
    if (tid == 0) {
      sigma[(size_t)b] = 0;
      if (price_out) price_out[(size_t)b] = 0.0;
      if (rounds_out) rounds_out[b] = 0;
    }
    return;
  }
  for (int i = tid; i < N; i += AU_THREADS) S.delta[i] = (float)(SHWD_AU_DELTA0 * cmax * Sd / N);
  const i64 eps_final = max((i64)(cmax * Sd * 9.094947017729282e-13), (i64)1);  // 2^-40 Cmax: 32 .. 64 units
  const i64 eps_list = (i64)(cmax * Sd * SHWD_AU_LIST_EPS);
#ifdef SHWD_AU_PROFILE
  long long pf_hist_c[10] = {0}, pf_hist_n[10] = {0}, pf_hist_p[10] = {0};
  long long pf_q2 = 0;
  long long pf_top2 = 0, pf_neval = 0, pf_start = 0, pf_gs = 0, pf_rescan = 0, pf_apply = 0, pf_nresc = 0, pf_nlist = 0, pf_rounds = 0;
#endif
  long long bids = 0;  // (warp 0's count of list bids + every thread's view of the parallel rounds is not needed: info only)
  int failed = 0;
  for (i64 eps = (i64)(cmax * Sd * SHWD_AU_EPS0);; eps = max((i64)((double)eps * AU_EPS_FACTOR), eps_final)) {
    const bool lists = eps <= eps_list;  // larger eps: plain parallel rounds of full scans
    // ================= phase start: everybody unassigned, prices and lists kept; one parallel round from the lists
    __syncthreads();
    AU_PROF_T(t_start);
    for (int i = tid; i < N; i += AU_THREADS) S.owner[i] = -1;
    int pj[AU_PP];
    i64 pbid[AU_PP];
#pragma unroll
    for (int q = 0; q < AU_PP; ++q) {
      const int i = tid + q * AU_THREADS;
      pj[q] = -1;
      if (i < N) {
        const unsigned short* li = S.lidx + (size_t)i * K;
        const float* lc = S.lcost + (size_t)i * K;
        const i64 tw = S.Tw[i];
        if (tw > AU_NONE) {
          i64 w1 = AU_INF, w2 = AU_INF, p1 = 0;
          int j1 = -1;
          for (int k = 0; k < K; ++k) {
            const int j = li[k];
            if (j == 0xFFFF) break;
            const i64 p = S.price[j];
            const i64 w = fix(lc[k]) + p;
            if (w < w1) {
              w2 = w1;
              w1 = w;
              j1 = j;
              p1 = p;
            } else if (w < w2) {
              w2 = w;
            }
          }
          if (w2 <= tw) {
            pj[q] = j1;
            pbid[q] = p1 + (w2 - w1) + eps;
          } else {
            S.Tw[i] = w2;  // cut guess for the rescan
          }
        }
      }
    }
    __syncthreads();  // (owner reset above is complete; bids go in)
#pragma unroll
    for (int q = 0; q < AU_PP; ++q) {
      const int i = tid + q * AU_THREADS;
      if (pj[q] >= 0)
        atomicMax(S.bidval + pj[q], ((unsigned long long)pbid[q] & ~0x7FFull) | (unsigned long long)(AU_MAX_N - 1 - i));
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < AU_PP; ++q) {
      const int i = tid + q * AU_THREADS;
      if (i < N) {
        int flag = 2;  // no usable list: rescan
        if (pj[q] >= 0) {
          if ((int)(S.bidval[pj[q]] & 0x7FFull) == AU_MAX_N - 1 - i) {
            S.owner[pj[q]] = i;
            S.price[pj[q]] = pbid[q];
            flag = 0;
          } else {
            flag = 1;  // lost: bids again from its list
          }
        }
        S.R[i] = flag;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < AU_PP; ++q)
      if (pj[q] >= 0) S.bidval[pj[q]] = 0ull;
    int qhead = 0, qcount = 0, nR = 0;  // warp 0's registers
    if (warp == 0) {
      // ordered compaction of the flags (read from R in batches before R is overwritten below the read position)
      for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        const int flag = i < N ? S.R[i] : 0;
        const unsigned b1 = __ballot_sync(0xffffffffu, flag == 1), b2 = __ballot_sync(0xffffffffu, flag == 2);
        const unsigned lt = (1u << lane) - 1u;
        __syncwarp();
        if (flag == 1) S.queue[qcount + __popc(b1 & lt)] = i;
        if (flag == 2) S.R[nR + __popc(b2 & lt)] = i;
        qcount += __popc(b1);
        nR += __popc(b2);
        __syncwarp();
      }
    }
    AU_PROF_ADD(pf_start, t_start);
    // ================= the phase: list bids by warp 0, batched rescans + a Jacobi round by everybody
    for (;;) {
      AU_PROF_T(t_gs);
      if (warp == 0) {
        __syncwarp();
        while (qcount > 0) {
          if (lists && qcount >= 2) {
            // ---- two persons per pass, one per half-warp: A = queue head bids first (Gauss-Seidel order); B's evaluation is
            // taken only if A's bid did not touch an object of B's list (otherwise B stays at the head of the queue)
            const int half = lane >> 4, hl = lane & 15;
            const int q1 = (qhead + 1 == N) ? 0 : qhead + 1;
            const int i2 = S.queue[half ? q1 : qhead];
            int j2 = 0xFFFF;
            float c2 = 0.f;
            if (hl < K) {
              j2 = S.lidx[(size_t)i2 * K + hl];
              c2 = S.lcost[(size_t)i2 * K + hl];
            }
            i64 p2 = 0, wv = AU_INF;
            int own2 = -1;
            if (j2 != 0xFFFF) {
              p2 = S.price[j2];
              own2 = S.owner[j2];
              wv = fix(c2) + p2;
            }
            const i64 tw2 = S.Tw[i2];
            i64 h1, h2;
            int hb;
            au_top2_half(wv, hl, h1, h2, hb);
            const bool ok2 = h2 <= tw2;  // (uniform within a half)
            const bool okA = __shfl_sync(0xffffffffu, (int)ok2, 0) != 0, okB = __shfl_sync(0xffffffffu, (int)ok2, 16) != 0;
            const int lbA = __shfl_sync(0xffffffffu, hb, 0), lbB = 16 + __shfl_sync(0xffffffffu, hb, 16);
            const int jA = __shfl_sync(0xffffffffu, j2, lbA);
            const bool clash = okA && __ballot_sync(0xffffffffu, half == 1 && j2 == jA) != 0u;
            const int iA = __shfl_sync(0xffffffffu, i2, 0), iB = __shfl_sync(0xffffffffu, i2, 16);
#ifdef SHWD_AU_PROFILE
            pf_neval += clash ? 1 : 2;
#endif
            // A
            qhead = q1;
            --qcount;
            if (!okA) {
              if (lane == 0) {
                S.R[nR] = iA;
                S.Tw[iA] = h2;
              }
              ++nR;
            } else {
              if (lane == lbA) {
                S.price[j2] = p2 + (h2 - h1) + eps;
                S.owner[j2] = iA;
              }
              const int oldA = __shfl_sync(0xffffffffu, own2, lbA);
              if (oldA >= 0) {
                int qt = qhead + qcount;
                if (qt >= N) qt -= N;
                if (lane == 0) S.queue[qt] = oldA;
                ++qcount;
              }
              ++bids;
            }
            // B
            if (!clash) {
              qhead = (qhead + 1 == N) ? 0 : qhead + 1;
              --qcount;
              if (!okB) {
                if (lane == 16) {
                  S.R[nR] = iB;
                  S.Tw[iB] = h2;
                }
                ++nR;
              } else {
                if (lane == lbB) {
                  S.price[j2] = p2 + (h2 - h1) + eps;
                  S.owner[j2] = iB;
                }
                const int oldB = __shfl_sync(0xffffffffu, own2, lbB);
                if (oldB >= 0) {
                  int qt = qhead + qcount;
                  if (qt >= N) qt -= N;
                  if (lane == 0) S.queue[qt] = oldB;
                  ++qcount;
                }
                ++bids;
              }
            }
            __syncwarp();
            if (bids > AU_MAX_BIDS) {
              failed = 1;
              break;
            }
            continue;
          }
          const int i = S.queue[qhead];
          qhead = (qhead + 1 == N) ? 0 : qhead + 1;
          --qcount;
          if (!lists) {  // no lists in the large-eps phases: straight to the next parallel round
            if (lane == 0) S.R[nR] = i;
            ++nR;
            continue;
          }
          int j = 0xFFFF;
          float c = 0.f;
          if (lane < K) {
            j = S.lidx[(size_t)i * K + lane];
            c = S.lcost[(size_t)i * K + lane];
          }
          i64 p = 0, w = AU_INF;
          int own = -1;
          if (j != 0xFFFF) {
            p = S.price[j];
            own = S.owner[j];
            w = fix(c) + p;
          }
          const i64 tw = S.Tw[i];
          i64 w1, w2;
          int l1;
#ifdef SHWD_AU_PROFILE
          if (qcount >= 1) ++pf_q2;  // (another person was waiting in the queue when this one was popped)
#endif
          AU_PROF_T(t_t2);
          au_top2(w, lane, w1, w2, l1);
          AU_PROF_ADD(pf_top2, t_t2);
#ifdef SHWD_AU_PROFILE
          ++pf_neval;
#endif
          if (!(w2 <= tw)) {  // list exhausted: the best or second best may be an unlisted object
            if (lane == 0) {
              S.R[nR] = i;
              S.Tw[i] = w2;  // the rescan's cut guess: the true second best is <= the listed one (+inf: none)
            }
            ++nR;
            continue;
          }
          if (lane == l1) {
            S.price[j] = p + (w2 - w1) + eps;
            S.owner[j] = i;
          }
          const int old = __shfl_sync(0xffffffffu, own, l1);
          if (old >= 0) {
            int qt = qhead + qcount;
            if (qt >= N) qt -= N;
            if (lane == 0) S.queue[qt] = old;
            ++qcount;
          }
          __syncwarp();
#ifdef SHWD_AU_PROFILE
          ++pf_nlist;
#endif
          if (++bids > AU_MAX_BIDS) {
            failed = 1;
            break;
          }
        }
        if (lane == 0) {
          s_ctl[0] = nR;
          s_ctl[3] = failed;
        }
      }
      __syncthreads();
      AU_PROF_ADD(pf_gs, t_gs);
#ifdef SHWD_AU_PROFILE
      ++pf_rounds;
#endif
      const int nRall = s_ctl[0];
      failed = s_ctl[3];
      if (nRall == 0 || failed) break;
      for (int c0 = 0; c0 < nRall; c0 += AU_RC) {
        const int nC = min(AU_RC, nRall - c0);
        AU_PROF_T(t_rs);
        if (nC <= AU_WARPS / 2) {  // few persons: several warps share a person's scan
          int G = AU_WARPS;
          while (G * nC > AU_WARPS) G >>= 1;
          const AuBest best = au_rescan_coop<FAST>(cp, S, Xb, Cb, fix, N, K, c0, nC, G, lists, s_part, s_stg_i, s_stg_c, warp, lane);
          const int r = warp / G;
          if (r < nC && warp % G == 0 && lane == 0) {
            const i64 bid = S.price[best.j1] + (best.w2 - best.w1) + eps;
            S.rj[r] = best.j1;
            S.rbid[r] = bid;
            atomicMax(S.bidval + best.j1, ((unsigned long long)bid & ~0x1FFull) | (unsigned long long)(AU_RC - 1 - r));
          }
        } else {
        for (int r = warp; r < nC; r += AU_WARPS) {
            const int i = S.R[c0 + r];
            const float4 o = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(Xb + i);
            const float* crow = DENSE ? Cb + (size_t)i * N : nullptr;
            const AuBest best = lists ? au_rescan<FAST>(cp, S, o, crow, fix, N, K, i, lane)
                                      : au_scan_top2<FAST>(cp, S, o, crow, fix, N, lane);
            if (lane == 0) {
              const i64 bid = S.price[best.j1] + (best.w2 - best.w1) + eps;
              S.rj[r] = best.j1;
              S.rbid[r] = bid;
              atomicMax(S.bidval + best.j1, ((unsigned long long)bid & ~0x1FFull) | (unsigned long long)(AU_RC - 1 - r));
            }
          }
        }
        __syncthreads();
        AU_PROF_ADD(pf_rescan, t_rs);
#ifdef SHWD_AU_PROFILE
        {
          const int bk = (nC <= 8 ? 0 : nC <= 16 ? 1 : nC <= 64 ? 2 : nC < AU_RC ? 3 : 4) + (lists ? 5 : 0);
          pf_hist_c[bk] += clock64() - t_rs;
          pf_hist_n[bk] += 1;
          pf_hist_p[bk] += nC;
        }
#endif
        AU_PROF_T(t_ap);
        if (warp == 0) {
          for (int base = 0; base < nC; base += 32) {
            const int r = base + lane;
            const bool valid = r < nC;
            int i = -1, j = 0, old = -1;
            bool win = false;
            if (valid) {
              i = S.R[c0 + r];
              j = S.rj[r];
              win = (int)(S.bidval[j] & 0x1FFull) == AU_RC - 1 - r;
              if (win) {
                old = S.owner[j];
                S.owner[j] = i;
                S.price[j] = S.rbid[r];
              }
            }
            const bool app = valid && (!win || old >= 0);  // losers bid again from their fresh lists; displaced owners too
            const unsigned ba = __ballot_sync(0xffffffffu, app);
            if (app) {
              int qt = qhead + qcount + __popc(ba & ((1u << lane) - 1u));
              if (qt >= N) qt -= N;
              S.queue[qt] = win ? old : i;
            }
            qcount += __popc(ba);
          }
          __syncwarp();
          for (int r = lane; r < nC; r += 32) S.bidval[S.rj[r]] = 0ull;
          bids += nC;
        }
        __syncthreads();
        AU_PROF_ADD(pf_apply, t_ap);
#ifdef SHWD_AU_PROFILE
        pf_nresc += nC;
#endif
      }
      nR = 0;
    }
    if (failed || eps <= eps_final) break;
  }
  __syncthreads();
  if (failed) {
    for (int i = tid; i < N; i += AU_THREADS) sigma[(size_t)b * N + i] = i;  // (keep indices in range)
  } else {
    for (int j = tid; j < N; j += AU_THREADS) sigma[(size_t)b * N + S.owner[j]] = j;
  }
  for (int i = tid; i < N; i += AU_THREADS)
    if (price_out) price_out[(size_t)b * N + i] = (double)S.price[i] * unit;
  if (tid == 0) {
    if (rounds_out) rounds_out[b] = (int)(bids > INT_MAX ? INT_MAX : bids);
    if (failed) atomicExch(status, 1);
#ifdef SHWD_AU_PROFILE
    if (price_out && N >= 9) {
      double* po = price_out + (size_t)b * N;
      po[0] = (double)pf_start; po[1] = (double)pf_gs; po[2] = (double)pf_rescan; po[3] = (double)pf_apply;
      po[4] = (double)pf_nresc; po[5] = (double)pf_nlist; po[6] = (double)pf_rounds; po[7] = (double)pf_top2; po[8] = (double)pf_neval; po[9] = (double)pf_q2;
      if (N >= 40) for (int q = 0; q < 10; ++q) { po[10 + 3 * q] = (double)pf_hist_c[q]; po[11 + 3 * q] = (double)pf_hist_n[q]; po[12 + 3 * q] = (double)pf_hist_p[q]; }
    }
#endif
  }
}

}  // namespace shwd

using namespace shwd;

static size_t auction_smem(int N, int K, bool dense) {
  return (size_t)N * ((dense ? 0 : sizeof(float4)) + 3 * sizeof(double) + sizeof(float) + 3 * sizeof(int) +
                      (size_t)K * (sizeof(float) + sizeof(unsigned short))) +
         AU_RC * (sizeof(double) + sizeof(int)) + 16;
}
constexpr size_t AU_SMEM_BUDGET = 222 * 1024;  // + ~4.5 KB of static shared memory
// list width: 16 candidates per person where they fit next to the points, 8 for the largest clouds
static int auction_list_width(int N, bool dense) {
  if (auction_smem(N, 16, dense) <= AU_SMEM_BUDGET) return 16;
  return 8;
}

extern "C" int shwd_exact_assignment_max_points(void) {
  int n = 1;
  while (n < AU_MAX_N && auction_smem(n + 1, 8, false) <= AU_SMEM_BUDGET) ++n;
  return n;
}

extern "C" int shwd_exact_assignment(const float* x4, const float* y4, int B, int N, int cost_kind, float p, float n_power,
                                     int* sigma, double* prices, int* rounds, int* status, void* stream) {
  if (!x4 || !y4 || !sigma || !status || B < 0 || N <= 0 || cost_kind < 0 || cost_kind > 3 || !(p > 0.f))
    return SHWD_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(x4) & 15) || (reinterpret_cast<uintptr_t>(y4) & 15)) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  if (N > shwd_exact_assignment_max_points()) return SHWD_ERR_UNSUPPORTED;
  const int K = auction_list_width(N, false);
  const size_t smem = auction_smem(N, K, false);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SHWD_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), s));
  const int fast = pick_fast(cost_kind, p, n_power);
  const CostParams cp = make_cost_unit(cost_kind, p, n_power);  // k = 1: the functor returns the plain float32 cost
  const float4* X = reinterpret_cast<const float4*>(x4);
  const float4* Y = reinterpret_cast<const float4*>(y4);
#define SHWD_LAUNCH_AUCTION(F)                                                                                              \
  do {                                                                                                                      \
    if (smem > 32 * 1024) /* static + dynamic beyond 48 KB needs the opt-in */                                              \
      SHWD_CUDA_CHECK(cudaFuncSetAttribute(auction_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    auction_kernel<F><<<B, AU_THREADS, smem, s>>>(X, Y, nullptr, N, K, cp, sigma, prices, rounds, status);                  \
  } while (0)
  switch (fast) {
    case FAST_GEO2: SHWD_LAUNCH_AUCTION(FAST_GEO2); break;
    case FAST_SQE2: SHWD_LAUNCH_AUCTION(FAST_SQE2); break;
    default: SHWD_LAUNCH_AUCTION(GENERIC);
  }
#undef SHWD_LAUNCH_AUCTION
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// The same solve on an explicit cost matrix C (B, N, N) float32: what ot.emd2(a, b, M) receives at
// main_rotation.py:63-79 (POT_loss) and in the notebooks' W2 metric.  Uniform weights, square.
extern "C" int shwd_exact_assignment_dense(const float* C, int B, int N, int* sigma, double* prices, int* rounds, int* status,
                                           void* stream) {
  if (!C || !sigma || !status || B < 0 || N <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  if (N > shwd_exact_assignment_max_points()) return SHWD_ERR_UNSUPPORTED;
  const int K = auction_list_width(N, true);
  const size_t smem = auction_smem(N, K, true);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SHWD_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), s));
  const CostParams cp = make_cost_unit(SHWD_COST_SQEUCLID, 2.f, 1.f);  // unused by the dense functor
  if (smem > 32 * 1024)
    SHWD_CUDA_CHECK(cudaFuncSetAttribute(auction_kernel<AU_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auction_kernel<AU_DENSE><<<B, AU_THREADS, smem, s>>>(nullptr, nullptr, C, N, K, cp, sigma, prices, rounds, status);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// Exact optimal assignment between two equally sized clouds on an on-the-fly cost: the GPU replacement
// for the per-pair exact solve the reference's W_COS path really runs,
//     ot.emd2(a_i, b_i, C_i)        Point_Cloud_Resistration/losses/s2_wasserstein.py:39-50, 99-110
//                                   Comparison_Wasserstein_with_Chamfer_distance/main_rotation.py:63-79
// (POT's C++ network simplex on a float64 copy of C, one pair at a time on the CPU, with a GPU->CPU->GPU round trip).
// For uniform weights and n == m the LP optimum is attained at a permutation, so emd2 = (1/n) * min-assignment cost.
// POT is a third-party dependency that is not vendored (version un-pinned): this solves the same LP by a different exact
// method, Bertsekas' forward auction with epsilon-scaling, in float64 on the float32 cost values:
//   * an unassigned person i bids for its best object j1 = argmax_j (-C_ij - price_j) the price
//     price_j1 + (best - second best) + eps; every object goes to its highest bidder (ties: lowest person index), the
//     previous owner becomes unassigned;  a phase ends when everybody is assigned;
//   * eps starts at Cmax/4 and is divided by 5 per phase down to Cmax * 2^-40, prices are kept between phases.  At the
//     end the assignment is within n*eps of optimal -- far below one float32 ulp of the cost sum -- i.e. it is the
//     optimum unless two assignments tie to ~1e-9 relative.
// One CTA per pair; points, prices and the assignment live in shared memory; a warp serves one bidder at a time and
// scans the objects with the same cost functor the Sinkhorn sweeps use (the N x N matrix is never formed).  Everything
// is deterministic (no order-dependent atomics: bids are resolved by atomicMax on the price, then atomicMin on the index).
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
constexpr int AU_MAX_ROUNDS = 4000000;  // safety net (a phase needs O(n) rounds in practice)
// FAST value of the dense variant: the cost is READ from a caller-supplied (B, N, N) matrix instead of being evaluated from
// points -- the drop-in for ot.emd2(a, b, M) called on an explicit cost matrix (main_rotation.py:63-79 POT_loss; notebooks).
constexpr int AU_DENSE = 1000;

__device__ __forceinline__ unsigned long long au_key(double v) { return (unsigned long long)__double_as_longlong(v); }

struct AuBest {
  double v1, v2;
  int j1;
};
__device__ __forceinline__ AuBest au_merge(const AuBest& a, const AuBest& b) {
  AuBest r;
  const bool takea = (a.v1 > b.v1) || (a.v1 == b.v1 && a.j1 <= b.j1);
  if (takea) {
    r.v1 = a.v1;
    r.j1 = a.j1;
    r.v2 = fmax(a.v2, b.v1);
  } else {
    r.v1 = b.v1;
    r.j1 = b.j1;
    r.v2 = fmax(b.v2, a.v1);
  }
  return r;
}

// One warp's scan of the objects j0, j0 + step, ... for a bidder at `o`: best and second-best value -C - price.
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
template <int FAST>
__device__ __forceinline__ AuBest au_scan(const CostParams& cp, const float4 o, const float4* sY, const float* crow,
                                          const double* price, int N, int j0, int step) {
  AuBest best = {-INFINITY, -INFINITY, INT_MAX};
  for (int j = j0; j < N; j += step) {
    const double v = -(double)au_cost<FAST>(cp, o, sY, crow, j) - price[j];
    if (v > best.v1) {
      best.v2 = best.v1;
      best.v1 = v;
      best.j1 = j;
    } else if (v > best.v2) {
      best.v2 = v;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    AuBest other;
    other.v1 = __shfl_xor_sync(0xffffffffu, best.v1, s);
    other.v2 = __shfl_xor_sync(0xffffffffu, best.v2, s);
    other.j1 = __shfl_xor_sync(0xffffffffu, best.j1, s);
    best = au_merge(best, other);
  }
  return best;
}
__device__ __forceinline__ void au_bid(AuBest best, double eps, const double* price, int idx, int* lobj, double* lbid,
                                       unsigned long long* bidval) {
  if (best.j1 == INT_MAX) best.j1 = 0;  // every value NaN (non-finite input): keep the indices in range
  const double gap = (best.v2 == -INFINITY) ? 0.0 : best.v1 - best.v2;  // N == 1: no second best
  const double bid = price[best.j1] + gap + eps;
  lobj[idx] = best.j1;
  lbid[idx] = bid;
  atomicMax(bidval + best.j1, au_key(bid));  // prices are >= 0 and only rise: the bit pattern orders like the value
}

template <int FAST>
__global__ void __launch_bounds__(AU_THREADS) auction_kernel(const float4* __restrict__ X, const float4* __restrict__ Y,
                                                             const float* __restrict__ Cd, int N, CostParams cp,
                                                             int* __restrict__ sigma, double* __restrict__ price_out,
                                                             int* __restrict__ rounds_out, int* __restrict__ status) {
  constexpr bool DENSE = FAST == AU_DENSE;
  extern __shared__ float4 au_smem[];  // carved in decreasing alignment: float4, 8-byte, 4-byte arrays
  float4* sX = au_smem;                                               // N
  float4* sY = sX + N;                                                // N
  double* price = reinterpret_cast<double*>(sY + N);                  // N
  double* lbid = price + N;                                           // N   bid of list entry
  unsigned long long* bidval = reinterpret_cast<unsigned long long*>(lbid + N);  // N   highest bid per object (0: none)
  int* owner = reinterpret_cast<int*>(bidval + N);                        // N   person holding object j (-1: free)
  int* objof = owner + N;                                             // N   object of person i (-1: unassigned)
  int* bidder = objof + N;                                            // N   winning person per object this round
  int* list = bidder + N;                                             // N   unassigned persons
  int* lobj = list + N;                                               // N   best object of list entry
  int* list2 = lobj + N;                                              // N   next round's unassigned persons
  __shared__ int s_count;
  __shared__ AuBest s_part[AU_WARPS];
  __shared__ double s_red[AU_WARPS];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* Cb = DENSE ? Cd + (size_t)b * N * N : nullptr;  // row i of this pair's matrix: Cb + i * N
  for (int i = threadIdx.x; i < N; i += AU_THREADS) {
    sX[i] = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(X + (size_t)b * N + i);
    sY[i] = DENSE ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(Y + (size_t)b * N + i);
    price[i] = 0.0;
    bidval[i] = 0ull;
    bidder[i] = INT_MAX;
  }
  __syncthreads();
  // Cmax = max_ij kC_ij (sets the epsilon schedule)
  float cm = 0.f;
  bool nonfinite = false;
  for (int i = warp; i < N; i += AU_WARPS) {
    const float4 o = sX[i];
    for (int j = lane; j < N; j += 32) {
      const float cij = au_cost<FAST>(cp, o, sY, DENSE ? Cb + (size_t)i * N : nullptr, j);
      nonfinite = nonfinite || !(fabsf(cij) <= 3.0e38f);
      cm = fmaxf(cm, cij);
    }
  }
  cm = warp_max(cm);
  if (lane == 0) s_red[warp] = (double)cm;
  if (__syncthreads_or(nonfinite)) {
    // a NaN / infinite cost (a diverged model upstream): no assignment is meaningful and the bidding would run to
    // AU_MAX_ROUNDS.  Report failure at once; sigma = identity keeps every downstream gather in range.
    for (int i = threadIdx.x; i < N; i += AU_THREADS) {
      sigma[(size_t)b * N + i] = i;
      if (price_out) price_out[(size_t)b * N + i] = 0.0;
    }
    if (threadIdx.x == 0) {
      if (rounds_out) rounds_out[b] = 0;
      atomicExch(status, 1);
    }
    return;
  }
  double cmax = 0.0;
  for (int w = 0; w < AU_WARPS; ++w) cmax = fmax(cmax, s_red[w]);
  if (!(cmax > 0.0)) cmax = 1.0;  // all costs zero (or NaN): any assignment is optimal; run one trivial phase
  const double eps_final = cmax * 9.094947017729282e-13;  // 2^-40
  int rounds = 0;
  bool failed = false;
  int* cur = list;    // unassigned persons of this round
  int* nxt = list2;   // ... of the next round (built incrementally from the losers and the displaced owners)
  for (double eps = cmax * 0.25;; eps = fmax(eps * AU_EPS_FACTOR, eps_final)) {
    // ---- a phase: everybody unassigned, prices kept
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += AU_THREADS) {
      owner[i] = -1;
      objof[i] = -1;
      cur[i] = i;
    }
    int U = N;
    for (;;) {
      __syncthreads();
      if (U == 0) break;
      if (++rounds > AU_MAX_ROUNDS) {
        failed = true;
        break;
      }
      if (threadIdx.x == 0) s_count = 0;
      // ---- bidding.  Many bidders: one warp each.  Few bidders (the long tail of a phase): wpb warps share a bidder's
      // scan and their partial (best, second best) are merged through shared memory.
      int wpb = 1;
      while (wpb * 2 * U <= AU_WARPS) wpb *= 2;
      if (wpb == 1) {
        for (int idx = warp; idx < U; idx += AU_WARPS) {
          const AuBest best = au_scan<FAST>(cp, sX[cur[idx]], sY, DENSE ? Cb + (size_t)cur[idx] * N : nullptr, price, N, lane, 32);
          if (lane == 0) au_bid(best, eps, price, idx, lobj, lbid, bidval);
        }
      } else {
        const int idx = warp / wpb, sub = warp % wpb;
        if (idx < U) {
          const AuBest best = au_scan<FAST>(cp, sX[cur[idx]], sY, DENSE ? Cb + (size_t)cur[idx] * N : nullptr, price, N, sub * 32 + lane, 32 * wpb);
          if (lane == 0) s_part[warp] = best;
        }
        __syncthreads();
        if (threadIdx.x < U) {
          AuBest best = s_part[threadIdx.x * wpb];
          for (int q = 1; q < wpb; ++q) best = au_merge(best, s_part[threadIdx.x * wpb + q]);
          au_bid(best, eps, price, threadIdx.x, lobj, lbid, bidval);
        }
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < U; idx += AU_THREADS)
        if (au_key(lbid[idx]) == bidval[lobj[idx]]) atomicMin(bidder + lobj[idx], cur[idx]);
      __syncthreads();
      for (int idx = threadIdx.x; idx < U; idx += AU_THREADS) {
        const int j = lobj[idx], i = cur[idx];
        if (bidder[j] == i) {
          const int old = owner[j];
          if (old >= 0) {
            objof[old] = -1;
            nxt[atomicAdd(&s_count, 1)] = old;
          }
          owner[j] = i;
          objof[i] = j;
          price[j] = lbid[idx];
        } else {
          nxt[atomicAdd(&s_count, 1)] = i;
        }
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < U; idx += AU_THREADS) {
        bidval[lobj[idx]] = 0ull;
        bidder[lobj[idx]] = INT_MAX;
      }
      U = s_count;
      int* t = cur;
      cur = nxt;
      nxt = t;
    }
    if (failed || eps <= eps_final) break;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += AU_THREADS) {
    sigma[(size_t)b * N + i] = failed ? i : objof[i];  // (a failed solve leaves persons unassigned: keep indices in range)
    if (price_out) price_out[(size_t)b * N + i] = price[i];
  }
  if (threadIdx.x == 0) {
    if (rounds_out) rounds_out[b] = rounds;
    if (failed) atomicExch(status, 1);
  }
}

}  // namespace shwd

using namespace shwd;

static size_t auction_smem(int N) {
  return (size_t)N * (2 * sizeof(double) + sizeof(unsigned long long) + 2 * sizeof(float4) + 6 * sizeof(int));
}

extern "C" int shwd_exact_assignment_max_points(void) {
  int n = 1;
  while (auction_smem(n + 1) <= 220 * 1024) ++n;
  return n;
}

extern "C" int shwd_exact_assignment(const float* x4, const float* y4, int B, int N, int cost_kind, float p, float n_power,
                                     int* sigma, double* prices, int* rounds, int* status, void* stream) {
  if (!x4 || !y4 || !sigma || !status || B < 0 || N <= 0 || cost_kind < 0 || cost_kind > 3 || !(p > 0.f))
    return SHWD_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(x4) & 15) || (reinterpret_cast<uintptr_t>(y4) & 15)) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  const size_t smem = auction_smem(N);
  if (smem > 220 * 1024) return SHWD_ERR_UNSUPPORTED;
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
    auction_kernel<F><<<B, AU_THREADS, smem, s>>>(X, Y, nullptr, N, cp, sigma, prices, rounds, status);                     \
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

// The same solve on an explicit cost matrix C (B, N, N) float32 (costs >= 0, as every cost of this path is): what
// ot.emd2(a, b, M) receives at main_rotation.py:63-79 (POT_loss) and in the notebooks' W2 metric.  Uniform weights, square.
extern "C" int shwd_exact_assignment_dense(const float* C, int B, int N, int* sigma, double* prices, int* rounds, int* status,
                                           void* stream) {
  if (!C || !sigma || !status || B < 0 || N <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  const size_t smem = auction_smem(N);
  if (smem > 220 * 1024) return SHWD_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SHWD_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), s));
  const CostParams cp = make_cost_unit(SHWD_COST_SQEUCLID, 2.f, 1.f);  // unused by the dense functor
  if (smem > 32 * 1024)
    SHWD_CUDA_CHECK(cudaFuncSetAttribute(auction_kernel<AU_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auction_kernel<AU_DENSE><<<B, AU_THREADS, smem, s>>>(nullptr, nullptr, C, N, cp, sigma, prices, rounds, status);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

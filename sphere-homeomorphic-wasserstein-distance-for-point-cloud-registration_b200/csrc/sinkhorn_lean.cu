// Entropic OT for SMALL problems: the same sweeps as sinkhorn.cu (same inner loops, same roundings, same history
// format -- the two forward / backward kernels are interchangeable), but a different work mapping and a half-step
// skeleton cut down to its latency floor.
//
// Why: the flattened (pair, owner group) deal of sinkhorn.cu keeps 148 SMs busy at B = 32, N = 1024, where a half-step
// is ~24 us of arithmetic.  The reference's other call shapes -- B = 32 with N = 128...256 (train_RUNNER.py:124-127),
// single pairs of N ~ 1000 (Flow_ellipsoid.ipynb cell 8), and the 4 pairs per GPU left by an 8-way strong-scaled batch
//  -- are 0.1-2 us of arithmetic per half-step, and there the ~4 us of per-half-step bookkeeping of the
// general kernel (item descriptors, two-segment staging, 16-way merges, counter round trips) is all that is measured
// (profiles/r01h_sweep_cfg5_cfg3.md: 0.098 of the roofline at B = 1, N = 1024; 0.237 at B = 32, N = 256).
//
// Mapping: every pair gets Q dedicated CTAs (Q = #SMs / B, or whole pairs per CTA when B exceeds the SM count).  CTA c
// of a pair owns a fixed BAND of x points (row owners) and of y points (col owners) for the whole launch: the packed
// coordinates of both clouds, the owner records and the owners' running state (previous potential; in the backward the
// previous adjoint and the gradient accumulators) stay in shared memory / registers from the first half-step to the
// last.  A band is ng <= 8 groups of RW owners; RW < 32 (8 or 16 owners per warp row) lets a single pair spread over
// up to 128 CTAs: the 32 / RW lane groups of a warp stream different sub-slices and are combined by shuffles.  The 16
// warps are split over (group pair, streamed slice), so a warp makes ONE pass per half-step.
//
// A half-step is then: poll the streamed potentials (write-once, sentinel-initialised planes: the poll IS the load) ->
// barrier -> sweep -> barrier -> merge + publish.  Nothing else crosses the CTA boundary, no descriptor is rebuilt, no
// counter is touched; the backward uses level-indexed write-once adjoint planes the same way.  Cluster / DSMEM exchange
// was considered and not used: a DSMEM hop (~215 cycles) and cluster.sync (~380) are no faster than the L2 store ->
// poll hop (~2 x 250), a cluster caps a pair at 16 SMs, and B = 1, N = 1024 wants 128.
//
// The two FINAL sweeps of either direction (4 of 404 sweeps) run through the general kernel's code (fwd_tail /
// bwd_phase in sinkhorn_core.cuh) on the flattened deal; the hand-over is one counter per pair.
#define SHWD_PROF_SYM g_prof_lean
#include "sinkhorn_core.cuh"
#include "sinkhorn_lean.h"
#include <stdlib.h>

namespace shwd {

constexpr int LEAN_PART_ROWS = 32;  // (slice, group) rows of 32 partials: NS * 2 * NGP = 32 always

struct LeanView {
  float2 *X, *P, *A, *S;  // X: 3T records (X, Y, Z); P, A, S: T records each
  float4* own;            // (x, y, z, previous potential) per (group, lane), owners replicated over the sub-slices
  float4* own3;           // backward: (own_pot1, o2, oadj)
  float* ownadj;          // backward: the owners' latest adjoint
  const float4* cth;      // this thread's column of the cached-intermediate table of the type (see lean_fill_cache)
  int nc;                 // record pairs per lane the table holds
  int T, ng, NGP, r0, r1, n_str;
};

template <bool BWD>
__device__ __forceinline__ LeanView lean_view(float4* smem4, const LeanGeom& gm, int type, int c, int n_own, int n_str) {
  LeanView v;
  constexpr int W = BWD ? 6 : 4;
  float2* F = reinterpret_cast<float2*>(smem4);
  const int T0 = gm.T[0], T1 = gm.T[1];
  float2* base = F + (type ? W * T0 : 0);
  const int T = type ? T1 : T0;
  v.X = base;
  v.P = base + 3 * T;
  v.A = base + 4 * T;
  v.S = base + 5 * T;
  float4* after = reinterpret_cast<float4*>(F + W * (T0 + T1));
  float4* own = after + LEAN_PART_ROWS * 32;
  v.own = own + type * GMAX * 32;
  v.own3 = own + (2 + type) * GMAX * 32;
  v.ownadj = reinterpret_cast<float*>(own + 4 * GMAX * 32) + type * GMAX * 32;
  // cached intermediates: after the owners' adjoints (2 * GMAX * 32 floats), one float4 per (record pair, owner group) and
  // thread, thread-major inside an entry (conflict-free: lane i reads word 4 i of a 512-B line)
  const float4* cache = own + 4 * GMAX * 32 + (2 * GMAX * 32) / 4;
  v.cth = cache + (type ? (size_t)gm.nc[0] * gm.cr[0] * SK_THREADS : 0) + threadIdx.x;
  v.nc = gm.nc[type];
  v.T = T;
  v.ng = gm.R[type] >> gm.rw_shift;
  const int ngp = (v.ng + 1) >> 1;
  v.NGP = ngp <= 1 ? 1 : (ngp <= 2 ? 2 : 4);
  v.r0 = min(n_own, c * gm.R[type]);
  v.r1 = min(n_own, v.r0 + gm.R[type]);
  v.n_str = n_str;
  return v;
}
__device__ __forceinline__ float4* lean_part(float4* smem4, const LeanGeom& gm, bool bwd) {
  return reinterpret_cast<float4*>(reinterpret_cast<float2*>(smem4) + (bwd ? 6 : 4) * (gm.T[0] + gm.T[1]));
}

// packed position of streamed point q: record t = q mod T, half = q / T
__device__ __forceinline__ int packed_pos(int q, int T) { return q >= T ? 2 * (q - T) + 1 : 2 * q; }

// resident coordinates of the streamed cloud + owner records of the band, once per (pair, launch)
__device__ __forceinline__ void lean_stage_resident(const LeanView& v, const float4* str, const float4* own, int rw_shift) {
  float* X = reinterpret_cast<float*>(v.X);
  float* Y = X + 2 * v.T;
  float* Z = Y + 2 * v.T;
  for (int q = threadIdx.x; q < 2 * v.T; q += SK_THREADS) {
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < v.n_str) c = __ldg(str + q);
    const int o = packed_pos(q, v.T);
    X[o] = c.x;
    Y[o] = c.y;
    Z[o] = c.z;
  }
  const int RW = 1 << rw_shift;
  for (int idx = threadIdx.x; idx < v.ng * 32; idx += SK_THREADS) {
    const int o = v.r0 + ((idx >> 5) << rw_shift) + (idx & (RW - 1));
    float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o < v.r1) rec = __ldg(own + o);
    rec.w = 0.f;
    v.own[idx] = rec;
  }
}

// Poll `cnt` write-once values (sentinel-initialised by the host) and scatter them into a packed array.
// MUL: dst[o] = value * dst[o] (the backward's adjoint times the staged 2^res correction).
template <bool MUL>
__device__ __forceinline__ void lean_poll(const float* src, int cnt, int T, float* dst, int* status) {
  for (int q0 = threadIdx.x; q0 < cnt; q0 += 4 * SK_THREADS) {
    unsigned raw[4];
    long long t0 = 0;
    for (unsigned tries = 0;; ++tries) {
      bool all = true;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = q0 + u * SK_THREADS;
        if (q < cnt && (tries == 0 || raw[u] == SPIN_SENTINEL)) raw[u] = ld_relaxed_u32(src + q);
        if (q >= cnt) raw[u] = 0u;
        all = all && (raw[u] != SPIN_SENTINEL);
      }
      if (all) break;
      if ((tries & 255u) == 255u) {  // a lost producer ends the launch instead of hanging the GPU
        if (t0 == 0) t0 = clock64();
        if (*reinterpret_cast<volatile int*>(status) != 0 || clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
          atomicExch(status, 1);
          break;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * SK_THREADS;
      if (q < cnt) {
        const int o = packed_pos(q, T);
        const float val = __uint_as_float(raw[u]);
        dst[o] = MUL ? val * dst[o] : val;
      }
    }
  }
}

// What a thread does in every half-step of one type -- fixed for the whole pair, so computed once (the runtime divisions
// and the address arithmetic otherwise cost ~150 instructions per warp and half-step, on the critical path).
struct LeanThread {
  int tb, te;      // packed records of this lane's sub-slice; te < 0: the warp has no group pair
  int slot;        // float4 index of the lane's partial in `part`
  int own;         // float4 index of the lane's first owner record (relative to LeanView::own / own3)
  int two;         // the warp sweeps two owner groups
  int pp;          // merge: float4 index of this owner's slice-0 partial; < 0: not an owner thread
  int pp_stride;   // float4 stride between the slice partials of one owner
  int NS;
};
__device__ __forceinline__ LeanThread lean_thread(const LeanView& v, int rw_shift) {
  LeanThread t;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = warp / v.NGP, gpi = warp % v.NGP;
  const int g = 2 * gpi;
  t.NS = SK_WARPS / v.NGP;
  const int SLr = v.T / t.NS;             // records per slice
  const int SLs = SLr >> (5 - rw_shift);  // records per sub-slice (multiple of 4)
  t.tb = slice * SLr + (lane >> rw_shift) * SLs;
  t.te = g < v.ng ? t.tb + SLs : -1;
  t.two = g + 1 < v.ng;
  t.slot = ((slice * 2 * v.NGP + g) << 5) + lane;
  t.own = g * 32 + lane;
  const int og = threadIdx.x >> rw_shift, oin = threadIdx.x & ((1 << rw_shift) - 1);
  const bool owner = (int)threadIdx.x < (v.ng << rw_shift) && v.r0 + (int)threadIdx.x < v.r1;
  t.pp = owner ? (og << 5) + oin : -1;
  t.pp_stride = (2 * v.NGP) << 5;
  return t;
}

// The potential-independent part of an element -- dot product and acos (pk_e) -- is the same in all 2L+1 sweeps of a
// launch, and in this regime a lane's elements are few: fill a thread-private shared-memory table with them once per pair
// and let the sweeps read it back (compute_packed_geo2<..., CACHED>).  The sweeps are XU-bound (sqrt + ex2 per forward
// element, sqrt + rsqrt + 2 ex2 backward); a cached element needs no sqrt and no dot / acos arithmetic.  The table holds
// the first v.nc record pairs of every lane's sub-slice (all of them when shared memory allows -- LeanGeom::nc); the
// rest of the sub-slice is swept by the recomputing loop, which continues from the partial result of the cached part.
template <int PK>
__device__ __forceinline__ void lean_fill_cache(const CostParams& cp, const LeanView& v, const LeanThread& lt) {
  if (lt.te < 0 || v.nc == 0) return;
  const int R = lt.two ? 2 : 1;
  float4* dst = const_cast<float4*>(v.cth);
  const float2 *Xa = v.X, *Ya = v.X + v.T, *Za = v.X + 2 * v.T;
  for (int r = 0; r < R; ++r) {
    const float4 op = v.own[lt.own + 32 * r];
    for (int it = 0; it < v.nc && lt.tb + 2 * it < lt.te; ++it) {
      const int t = lt.tb + 2 * it;
      const float4 X = *reinterpret_cast<const float4*>(Xa + t), Y = *reinterpret_cast<const float4*>(Ya + t);
      const float4 Z = *reinterpret_cast<const float4*>(Za + t);
      const f2 th0 = pk_e<PK>(cp, op, mk2(X.x, X.y), mk2(Y.x, Y.y), mk2(Z.x, Z.y));
      const f2 th1 = pk_e<PK>(cp, op, mk2(X.z, X.w), mk2(Y.z, Y.w), mk2(Z.z, Z.w));
      dst[(it * R + r) * SK_THREADS] = make_float4(lo2(th0), hi2(th0), lo2(th1), hi2(th1));
    }
  }
}

// One warp's pass of a half-step: its (group pair, streamed slice), then the combination of the sub-slices of a warp row.
template <int PK, int MODE, bool OFF>
__device__ __forceinline__ void lean_compute(const CostParams& cp, float lconst, const LeanView& v, const LeanThread& lt, float4* part,
                                             int rw_shift) {
  if (lt.te < 0) return;
  const int lane = threadIdx.x & 31;
  const bool two = lt.two != 0;
  const int tb = lt.tb, te = lt.te;
  PackedSmem pv;
  pv.X = v.X;
  pv.Y = v.X + v.T;
  pv.Z = v.X + 2 * v.T;
  pv.P = v.P;
  pv.A = v.A;
  pv.S = v.S;
  float4* slot = part + lt.slot;
  const float4* own = v.own + lt.own;
  const float4* own3 = v.own3 + lt.own;
  // the backward reads the table only for the acos-based costs (the others' intermediates are by-products of the gradient)
  constexpr bool USE = MODE == MODE_LSE || PK == FAST_GEO2 || PK == FAST_GEO1;
  const int tc = USE ? min(te, tb + 2 * v.nc) : tb;  // [tb, tc): cached record pairs; [tc, te): recomputed
  if (tc > tb) {
    if (two)
      compute_packed_geo2<PK, MODE, 2, OFF, true>(cp, lconst, pv, tb, tc, true, own, own3, slot, v.cth);
    else
      compute_packed_geo2<PK, MODE, 1, OFF, true>(cp, lconst, pv, tb, tc, true, own, own3, slot, v.cth);
  }
  if (tc < te) {
    if (two)
      compute_packed_geo2<PK, MODE, 2, OFF>(cp, lconst, pv, tc, te, tc == tb, own, own3, slot);
    else
      compute_packed_geo2<PK, MODE, 1, OFF>(cp, lconst, pv, tc, te, tc == tb, own, own3, slot);
  }
  if (rw_shift < 5) {
    const int RW = 1 << rw_shift;
    __syncwarp();
    for (int r = 0; r < (two ? 2 : 1); ++r) {
      float4 a = slot[32 * r];
      if (MODE == MODE_LSE && OFF) {
        for (int o = RW; o < 32; o <<= 1) a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      } else if (MODE == MODE_LSE) {
        for (int o = RW; o < 32; o <<= 1) {
          const float m2 = __shfl_xor_sync(0xffffffffu, a.x, o), s2 = __shfl_xor_sync(0xffffffffu, a.y, o);
          const float nm = fmaxf(a.x, m2);
          a.y = a.y * ex2_approx(a.x - nm) + s2 * ex2_approx(m2 - nm);
          a.x = nm;
        }
      } else {
        for (int o = RW; o < 32; o <<= 1) {
          a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
          a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
          a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
          a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
        }
      }
      __syncwarp();
      if (lane < RW) slot[32 * r] = a;
    }
  }
}

__device__ __forceinline__ void lean_signal(int* done_b) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    atomicAdd(done_b, 1);
  }
}

// pairs of this CTA: one pair (Q dedicated CTAs each) when the batch fits the grid, otherwise whole pairs in turn
__device__ __forceinline__ void lean_pairs(const SinkParams& prm, const LeanGeom& gm, int& b0, int& bstep, int& c) {
  if (gm.whole_pairs) {
    b0 = blockIdx.x;
    bstep = gridDim.x;
    c = 0;
  } else {
    b0 = blockIdx.x / gm.Q;
    bstep = prm.B;  // exactly one pair
    c = blockIdx.x % gm.Q;
  }
}

// ================================================================================================================
// Forward
// ================================================================================================================
template <int FAST>
__global__ void __launch_bounds__(SK_THREADS, 1) sinkhorn_fwd_lean_kernel(const SinkParams prm, const LeanGeom gm) {
  extern __shared__ float4 smem4[];
  __shared__ int s_ls;
  __shared__ SweepIO s_ios[2];
  constexpr int PK = is_packed_cost(FAST) ? FAST : FAST_GEO2;
  const int HL = prm.hist_levels, L = prm.iters;
  const int rw_shift = gm.rw_shift, RW = 1 << rw_shift, nsub = 32 >> rw_shift;
  float4* part = lean_part(smem4, gm, false);

  int b0, bstep, c;
  lean_pairs(prm, gm, b0, bstep, c);
  PROF_INIT();
  for (int b = b0; b < prm.B; b += bstep) {
    const LeanView V0 = lean_view<false>(smem4, gm, 0, c, prm.N, prm.M), V1 = lean_view<false>(smem4, gm, 1, c, prm.M, prm.N);
    __syncthreads();  // the previous pair's readers are done
    lean_stage_resident(V0, prm.Y + (size_t)b * prm.M, prm.X + (size_t)b * prm.N, rw_shift);
    lean_stage_resident(V1, prm.X + (size_t)b * prm.N, prm.Y + (size_t)b * prm.M, rw_shift);
    {  // padding potentials never change: -inf
      float* P0 = reinterpret_cast<float*>(V0.P);
      for (int q = V0.n_str + threadIdx.x; q < 2 * V0.T; q += SK_THREADS) P0[packed_pos(q, V0.T)] = -INFINITY;
      float* P1 = reinterpret_cast<float*>(V1.P);
      for (int q = V1.n_str + threadIdx.x; q < 2 * V1.T; q += SK_THREADS) P1[packed_pos(q, V1.T)] = -INFINITY;
    }
    // beta^0 = 0 (kept in the history so the backward can stream it)
    for (int j = V1.r0 + threadIdx.x; j < V1.r1; j += SK_THREADS) prm.beta[((size_t)b * HL) * prm.M + j] = 0.f;

    const LeanThread LT0 = lean_thread(V0, rw_shift), LT1 = lean_thread(V1, rw_shift);
    __syncthreads();  // the resident coordinates and owner records are in place
    lean_fill_cache<PK>(prm.cp, V0, LT0);
    lean_fill_cache<PK>(prm.cp, V1, LT1);
    for (int h = 0; h < 2 * L; ++h) {
      const int type = h & 1;
      const int l = (h >> 1) + 1;
      const LeanView v = type ? V1 : V0;
      const LeanThread lt = type ? LT1 : LT0;
      if (v.r1 <= v.r0) continue;  // (CTA-uniform) no owners of this kind in the band
      const float lconst = type ? prm.lb2 : prm.la2;
      float* P = reinterpret_cast<float*>(v.P);
      PROF_MARK(5);
      if (type == 0 && l == 1) {
        for (int q = threadIdx.x; q < v.n_str; q += SK_THREADS) P[packed_pos(q, v.T)] = 0.f;
      } else {
        const float* src = type ? prm.alpha + ((size_t)b * HL + l) * prm.N : prm.beta + ((size_t)b * HL + (l - 1)) * prm.M;
        lean_poll<false>(src, v.n_str, v.T, P, prm.status);
      }
      PROF_MARK(0);
      __syncthreads();
      PROF_MARK(1);
      float* out_pot = type ? prm.beta + ((size_t)b * HL + l) * prm.M : prm.alpha + ((size_t)b * HL + l) * prm.N;
      float* out_lo = type ? prm.beta_lo + ((size_t)b * HL + l) * prm.M : prm.alpha_lo + ((size_t)b * HL + l) * prm.N;
      bool off_try = SHWD_OFFSET_LSE && h >= 2;
      for (;;) {
        if (off_try)
          lean_compute<PK, MODE_LSE, true>(prm.cp, lconst, v, lt, part, rw_shift);
        else
          lean_compute<PK, MODE_LSE, false>(prm.cp, lconst, v, lt, part, rw_shift);
        PROF_MARK(2);
        __syncthreads();
        PROF_MARK(6);
        // merge the NS slice partials of every owner in fixed order
        const int o = v.r0 + threadIdx.x;
        const bool owner = lt.pp >= 0;
        float mx = NEG_BIG, sum = 0.f;
        bool bad = false;
        if (owner) {
          const float4* pp = part + lt.pp;
          if (off_try) {
            mx = pp[0].x;  // the common offset
            for (int s = 0; s < lt.NS; ++s) sum += pp[s * lt.pp_stride].y;
            bad = !(sum >= 0x1p-60f && sum <= 0x1p60f);
          } else {
            for (int s = 0; s < lt.NS; ++s) mx = fmaxf(mx, pp[s * lt.pp_stride].x);
            for (int s = 0; s < lt.NS; ++s) {
              const float4 st = pp[s * lt.pp_stride];
              sum += st.y * ex2_approx(st.x - mx);
            }
          }
        }
        if (off_try && __syncthreads_or(bad)) {  // (CTA-uniform) redo the half-step with the running maximum
          off_try = false;
          continue;
        }
        PROF_MARK(3);
        if (owner) {
          // new potential in double, float32 rounding residual kept for the backward (see finalize_visit)
          const double npd = (double)lconst - ((double)mx + fast_log2d(sum));
          const float np = (float)npd;
          out_lo[o] = (float)(npd - (double)np);
          out_pot[o] = np;
          float4* ow = v.own + lt.pp;
          for (int s = 0; s < nsub; ++s) ow[s << rw_shift].w = np;
        }
        PROF_MARK(4);
        break;
      }
    }
    lean_signal(prm.done + b);
  }
  __syncthreads();
  // the two FINAL sweeps and the cost, on the flattened deal with the general kernel's shared-memory carve-up
  float4* sS = smem4;
  float4* part_l = smem4 + 2 * CHUNK_PAD;
  float2* sAdj = reinterpret_cast<float2*>(part_l + SK_WARPS * GMAX * 32);
  float4* sOwn = reinterpret_cast<float4*>(sAdj + 2 * CHUNK_PAD);
  fwd_tail<FAST>(prm, gm.whole_pairs ? 1 : gm.Q, sS, sAdj, part_l, sOwn, s_ios, s_ls);
}

// ================================================================================================================
// Backward
// ================================================================================================================
template <int FAST>
__global__ void __launch_bounds__(SK_THREADS, 1) sinkhorn_bwd_lean_kernel(const SinkParams prm, const LeanGeom gm) {
  extern __shared__ float4 smem4[];
  __shared__ SweepIO s_ios[2];
  __shared__ ResidentType s_RT[2];
  constexpr int PK = is_packed_cost(FAST) ? FAST : FAST_GEO2;
  if (threadIdx.x == 0) s_RT[0].ok = s_RT[1].ok = 0;
  PROF_INIT();
  const int HL = prm.hist_levels;
  const int Ls = *prm.iters_run;
  const int rw_shift = gm.rw_shift, RW = 1 << rw_shift, nsub = 32 >> rw_shift;
  const int gr = (prm.N + 31) / 32, gc = (prm.M + 31) / 32;
  const size_t BN = (size_t)prm.B * prm.N, BM = (size_t)prm.B * prm.M;
  {
    // phases 0 and 1: the FINAL sweeps (l = L*), on the flattened deal; they leave abar^L*, bbar^(L*-1) and the first
    // gradient terms in global memory and count gr + gc groups per pair
    float4* sS = smem4;
    float4* part_l = smem4 + 2 * CHUNK_PAD;
    float2* sAdj = reinterpret_cast<float2*>(part_l + SK_WARPS * GMAX * 32);
    float4* sOwn = reinterpret_cast<float4*>(sAdj + 2 * CHUNK_PAD);
    __syncthreads();
    bwd_phase<FAST>(prm, 0, Ls, false, sS, sAdj, part_l, sOwn, s_ios, s_RT);
    bwd_phase<FAST>(prm, 1, Ls, false, sS, sAdj, part_l, sOwn, s_ios, s_RT);
  }
  float4* part = lean_part(smem4, gm, true);

  int b0, bstep, c;
  lean_pairs(prm, gm, b0, bstep, c);
  PROF_INIT();
  for (int b = b0; b < prm.B; b += bstep) {
    const LeanView V0 = lean_view<true>(smem4, gm, 0, c, prm.N, prm.M), V1 = lean_view<true>(smem4, gm, 1, c, prm.M, prm.N);
    wait_done(prm.done + b, gr + gc, prm.status);  // (ends with a barrier: the previous users of the shared arrays are done)
    lean_stage_resident(V0, prm.Y + (size_t)b * prm.M, prm.X + (size_t)b * prm.N, rw_shift);
    lean_stage_resident(V1, prm.X + (size_t)b * prm.N, prm.Y + (size_t)b * prm.M, rw_shift);
    for (int type = 0; type < 2; ++type) {  // padding records never change: potential -inf, adjoint 0, addend -inf
      const LeanView v = type ? V1 : V0;
      float* P = reinterpret_cast<float*>(v.P);
      float* A = reinterpret_cast<float*>(v.A);
      float* S = reinterpret_cast<float*>(v.S);
      for (int q = v.n_str + threadIdx.x; q < 2 * v.T; q += SK_THREADS) {
        const int o = packed_pos(q, v.T);
        P[o] = -INFINITY;
        A[o] = 0.f;
        S[o] = -INFINITY;
      }
    }
    // the owners' running state: latest adjoint (abar^L* / bbar^(L*-1) from the FINAL sweeps) and the gradient so far
    const float* al = prm.alpha + (size_t)b * HL * prm.N;
    const float* be = prm.beta + (size_t)b * HL * prm.M;
    const float* al_lo = prm.alpha_lo + (size_t)b * HL * prm.N;
    const float* be_lo = prm.beta_lo + (size_t)b * HL * prm.M;
    float* abar_b = prm.abar + (size_t)b * prm.N;  // + plane * BN
    float* bbar_b = prm.bbar + (size_t)b * prm.M;  // + plane * BM
    float4 Gx = make_float4(0.f, 0.f, 0.f, 0.f), Gy = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ox = V0.r0 + threadIdx.x, oy = V1.r0 + threadIdx.x;
    const bool own_x = threadIdx.x < (V0.ng << rw_shift) && ox < V0.r1;
    const bool own_y = threadIdx.x < (V1.ng << rw_shift) && oy < V1.r1;
    if (own_x) {
      Gx = __ldcg(prm.g4x + (size_t)b * prm.N + ox);
      V0.ownadj[threadIdx.x] = __ldcg(abar_b + (size_t)Ls * BN + ox);
    }
    if (own_y) {
      Gy = __ldcg(prm.g4y + (size_t)b * prm.M + oy);
      V1.ownadj[threadIdx.x] = __ldcg(bbar_b + (size_t)(Ls - 1) * BM + oy);
    }

    const LeanThread LT0 = lean_thread(V0, rw_shift), LT1 = lean_thread(V1, rw_shift);
    __syncthreads();  // the resident coordinates and owner records are in place
    if (PK == FAST_GEO2 || PK == FAST_GEO1) {
      lean_fill_cache<PK>(prm.cp, V0, LT0);
      lean_fill_cache<PK>(prm.cp, V1, LT1);
    }
    for (int ph = 2; ph <= 2 * Ls; ++ph) {
      const bool last = (ph == 2 * Ls);
      const int type = last ? 0 : (ph & 1);
      const int l = last ? 0 : Ls - (ph >> 1);
      const LeanView v = type ? V1 : V0;
      const LeanThread lt = type ? LT1 : LT0;
      if (v.r1 <= v.r0) continue;  // (CTA-uniform)
      float* P = reinterpret_cast<float*>(v.P);
      float* A = reinterpret_cast<float*>(v.A);
      float* S = reinterpret_cast<float*>(v.S);
      PROF_MARK(5);
      // ---- PRE (independent of the previous phase): streamed potential of the forward history, its float32 addend and
      // the 2^res correction (parked in A until the adjoint arrives); the owners' exponents
      {
        const float* spot = type ? al + (size_t)l * prm.N : be + (size_t)l * prm.M;
        const float* slo = type ? al_lo + (size_t)l * prm.N : be_lo + (size_t)l * prm.M;
        const double c1 = type ? (double)prm.la2 : (double)prm.lb2;
        for (int q = threadIdx.x; q < v.n_str; q += SK_THREADS) {
          const float pot = __ldcg(spot + q);
          const double full = (double)pot + (last ? 0.0 : (double)__ldcg(slo + q)) - c1;
          const float sadd = (float)full;
          const int o = packed_pos(q, v.T);
          P[o] = pot;
          S[o] = sadd;
          A[o] = last ? 0.f : exp2f((float)(full - (double)sadd));
        }
        const int g = threadIdx.x >> rw_shift, oin = threadIdx.x & (RW - 1);
        if (threadIdx.x < (v.ng << rw_shift)) {
          const int o = v.r0 + threadIdx.x;
          float4 e = make_float4(-INFINITY, -INFINITY, 0.f, 0.f);
          if (o < v.r1) {
            const float* p1 = type ? be + (size_t)(l - 1) * prm.M : al + (size_t)l * prm.N;
            const float* p2 = type ? be + (size_t)l * prm.M : al + (size_t)(l + 1) * prm.N;
            const float* lo2p = type ? be_lo + (size_t)l * prm.M : al_lo + (size_t)(l + 1) * prm.N;
            const double c2 = type ? (double)prm.lb2 : (double)prm.la2;
            if (!last) e.x = __ldcg(p1 + o);
            const double full = (double)__ldcg(p2 + o) + (double)__ldcg(lo2p + o) - c2;
            e.y = (float)full;
            e.z = v.ownadj[threadIdx.x] * exp2f((float)(full - (double)e.y));
          }
          float4* ow = v.own3 + (g << 5) + oin;
          for (int s = 0; s < nsub; ++s) ow[s << rw_shift] = e;
        }
      }
      PROF_MARK(1);
      // ---- POST: the streamed adjoint of the previous phase (same thread -> same records as PRE: no barrier needed)
      if (!last) {
        const float* sadj = type ? abar_b + (size_t)l * BN : bbar_b + (size_t)l * BM;
        lean_poll<true>(sadj, v.n_str, v.T, A, prm.status);
      }
      PROF_MARK(0);
      __syncthreads();
      PROF_MARK(6);
      lean_compute<PK, MODE_BWD, false>(prm.cp, 0.f, v, lt, part, rw_shift);
      PROF_MARK(2);
      __syncthreads();
      PROF_MARK(3);
      // ---- merge + publish
      if (lt.pp >= 0) {
        const int o = v.r0 + threadIdx.x;
        const float4* pp = part + lt.pp;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < lt.NS; ++s) {
          const float4 st = pp[s * lt.pp_stride];
          sum.x += st.x;
          sum.y += st.y;
          sum.z += st.z;
          sum.w += st.w;
        }
        if (!last) {
          const float adj = -sum.w;
          float* out = type ? bbar_b + (size_t)(l - 1) * BM : abar_b + (size_t)l * BN;
          out[o] = adj;
          v.ownadj[threadIdx.x] = adj;
        }
        const float gxs = sum.x * prm.cp.gscale, gys = sum.y * prm.cp.gscale, gzs = sum.z * prm.cp.gscale;
        if (type) {
          Gy.x += gxs;
          Gy.y += gys;
          Gy.z += gzs;
        } else {
          Gx.x += gxs;
          Gx.y += gys;
          Gx.z += gzs;
        }
      }
      PROF_MARK(4);
    }
    if (own_x) __stcg(prm.g4x + (size_t)b * prm.N + ox, make_float4(Gx.x, Gx.y, Gx.z, 0.f));
    if (own_y) __stcg(prm.g4y + (size_t)b * prm.M + oy, make_float4(Gy.x, Gy.y, Gy.z, 0.f));
  }
}

// ================================================================================================================
// Host side
// ================================================================================================================
static int g_path_mode = 0;  // 0: automatic, 1: always the flattened-deal kernels, 2: lean whenever eligible

bool lean_plan(int B, int N, int M, LeanGeom* out) {
  const int G = sm_count();
  const int big = N > M ? N : M;
  if (B <= 0 || N <= 0 || M <= 0 || big > CHUNK || G < 1) return false;
  LeanGeom gm = {};
  gm.whole_pairs = B > G;
  int Q = gm.whole_pairs ? 1 : G / B;
  const int need = (big + Q - 1) / Q;
  gm.rw_shift = need > 16 ? 5 : (need > 8 ? 4 : 3);
  const int RW = 1 << gm.rw_shift, nsub = 32 >> gm.rw_shift;
  int q_eff = 1;
  for (int type = 0; type < 2; ++type) {
    const int n_own = type ? M : N;
    int R = ((n_own + Q - 1) / Q + RW - 1) / RW * RW;
    if (R < RW) R = RW;
    if ((R >> gm.rw_shift) > GMAX) return false;
    gm.R[type] = R;
    const int q = (n_own + R - 1) / R;
    if (q > q_eff) q_eff = q;
  }
  gm.Q = q_eff;
  for (int type = 0; type < 2; ++type) {
    const int n_str = type ? N : M;
    const int ng = gm.R[type] >> gm.rw_shift;
    const int ngp = (ng + 1) >> 1;
    const int NGP = ngp <= 1 ? 1 : (ngp <= 2 ? 2 : 4);
    const int unit = (SK_WARPS / NGP) * nsub * 4;
    gm.T[type] = ((n_str + 1) / 2 + unit - 1) / unit * unit;
  }
  gm.grid = gm.whole_pairs ? G : B * gm.Q;
  // shared memory: the lean carve-up, then as much of the cached-intermediate table (lean_fill_cache) as fits in the SM's
  // 227 KB; the FINAL sweeps use the general kernel's carve-up over the same allocation
  const size_t lean_bytes = sizeof(float2) * 6 * (size_t)(gm.T[0] + gm.T[1]) + sizeof(float4) * (LEAN_PART_ROWS * 32 + 4 * GMAX * 32) +
                            sizeof(float) * 2 * GMAX * 32;
  if (lean_bytes > sinkhorn_smem_bytes()) return false;
  static const bool cache_on = !(getenv("SHWD_LEAN_CACHE") && getenv("SHWD_LEAN_CACHE")[0] == '0');
  long long slots = cache_on ? (long long)(226 * 1024 - lean_bytes) / (long long)(sizeof(float4) * SK_THREADS) : 0;  // float4 per thread
  for (int type = 0; type < 2; ++type) {
    const int ng = gm.R[type] >> gm.rw_shift;
    const int ngp = (ng + 1) >> 1;
    const int NGP = ngp <= 1 ? 1 : (ngp <= 2 ? 2 : 4);
    const int sls = (gm.T[type] / (SK_WARPS / NGP)) >> (5 - gm.rw_shift);  // records per lane sub-slice
    gm.cr[type] = ng >= 2 ? 2 : 1;
    long long nc = slots / gm.cr[type];
    if (nc > sls / 2) nc = sls / 2;
    nc &= ~1LL;  // the running-maximum loop advances two record pairs at a time
    gm.nc[type] = (int)nc;
    slots -= nc * gm.cr[type];
  }
  const size_t total = lean_bytes + sizeof(float4) * SK_THREADS * ((size_t)gm.nc[0] * gm.cr[0] + (size_t)gm.nc[1] * gm.cr[1]);
  gm.smem = (int)(total > sinkhorn_smem_bytes() ? total : sinkhorn_smem_bytes());
  if (out) *out = gm;
  return true;
}

// Automatic choice (measured on B200, profiles/r02a_ot_kernels.md).  The lean skeleton saves ~3 us of bookkeeping per
// half-step; what it can lose is SMs: Q is an integer, so B = 32, N = 1024 runs on 4 x 32 = 128 of 148 SMs.  It wins
// wherever a CTA's share of a half-step is small next to that saving (B = 16, N = 1024: 7.2 vs 7.8 ms) or no SM is left
// idle anyway; at the benchmark shape (8 owner groups x 1024 streamed points per CTA, ~24 us per half-step) the
// flattened deal's 148 balanced CTAs win (12.96 vs 13.46 ms).
static bool lean_auto(int B, int N, int M, const LeanGeom& gm) {
  const int G = sm_count();
  if (gm.whole_pairs) {  // whole pairs in rounds: worth it unless the last round is mostly empty
    const int rounds = (B + G - 1) / G;
    return 10LL * B >= 8LL * rounds * G;
  }
  if (100LL * gm.grid >= 98LL * G) return true;
  const long long per_cta = (long long)(gm.R[0] > gm.R[1] ? gm.R[0] : gm.R[1]) * (N > M ? N : M);  // element-evals per half-step
  return per_cta <= 160000;
}

bool lean_selected(int B, int N, int M, int fast, int hist_levels, float thresh) {
  if (g_path_mode == 1) return false;
  if (!is_packed_cost(fast) || hist_levels <= 1 || thresh > 0.f) return false;
  LeanGeom gm;
  if (!lean_plan(B, N, M, &gm)) return false;
  return g_path_mode == 2 || lean_auto(B, N, M, gm);
}

template <typename K>
static int launch_lean(K kernel, const SinkParams& prm, const LeanGeom& gm, cudaStream_t s) {
  const size_t smem = (size_t)gm.smem;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return SHWD_ERR_CUDA;
  }
  void* args[] = {const_cast<SinkParams*>(&prm), const_cast<LeanGeom*>(&gm)};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(gm.grid), dim3(SK_THREADS), args, smem, s);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return SHWD_ERR_CUDA;
  }
  return SHWD_OK;
}

int launch_lean_fwd(int fast, const SinkParams& prm, const LeanGeom& gm, cudaStream_t s) {
  switch (fast) {
    case FAST_GEO2: return launch_lean(sinkhorn_fwd_lean_kernel<FAST_GEO2>, prm, gm, s);
    case FAST_SQE2: return launch_lean(sinkhorn_fwd_lean_kernel<FAST_SQE2>, prm, gm, s);
    case FAST_GEO1: return launch_lean(sinkhorn_fwd_lean_kernel<FAST_GEO1>, prm, gm, s);
    case FAST_SQE1: return launch_lean(sinkhorn_fwd_lean_kernel<FAST_SQE1>, prm, gm, s);
    case FAST_EUC2: return launch_lean(sinkhorn_fwd_lean_kernel<FAST_EUC2>, prm, gm, s);
    case FAST_OMC2: return launch_lean(sinkhorn_fwd_lean_kernel<FAST_OMC2>, prm, gm, s);
    default: return SHWD_ERR_UNSUPPORTED;
  }
}

int launch_lean_bwd(int fast, const SinkParams& prm, const LeanGeom& gm, cudaStream_t s) {
  switch (fast) {
    case FAST_GEO2: return launch_lean(sinkhorn_bwd_lean_kernel<FAST_GEO2>, prm, gm, s);
    case FAST_SQE2: return launch_lean(sinkhorn_bwd_lean_kernel<FAST_SQE2>, prm, gm, s);
    case FAST_GEO1: return launch_lean(sinkhorn_bwd_lean_kernel<FAST_GEO1>, prm, gm, s);
    case FAST_SQE1: return launch_lean(sinkhorn_bwd_lean_kernel<FAST_SQE1>, prm, gm, s);
    case FAST_EUC2: return launch_lean(sinkhorn_bwd_lean_kernel<FAST_EUC2>, prm, gm, s);
    case FAST_OMC2: return launch_lean(sinkhorn_bwd_lean_kernel<FAST_OMC2>, prm, gm, s);
    default: return SHWD_ERR_UNSUPPORTED;
  }
}

}  // namespace shwd

extern "C" int shwd_sinkhorn_set_path(int mode) {
  if (mode < 0 || mode > 2) return SHWD_ERR_INVALID_ARGUMENT;
  shwd::g_path_mode = mode;
  return SHWD_OK;
}

extern "C" int shwd_sinkhorn_lean_regime(int B, int N, int M) {
  // geometry only (cost kind, history and early stop are checked per launch): 1 when the lean kernels would take this shape
  if (shwd::g_path_mode == 1) return 0;
  shwd::LeanGeom gm;
  if (!shwd::lean_plan(B, N, M, &gm)) return 0;
  return (shwd::g_path_mode == 2 || shwd::lean_auto(B, N, M, gm)) ? 1 : 0;
}

#ifdef SHWD_PROFILE
extern "C" int shwd_prof_read_lean(unsigned long long* out8, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out8, shwd::g_prof_lean, sizeof(unsigned long long) * 8);
  if (e != cudaSuccess) return SHWD_ERR_CUDA;
  if (reset) {
    unsigned long long z[8] = {0};
    e = cudaMemcpyToSymbol(shwd::g_prof_lean, z, sizeof(z));
    if (e != cudaSuccess) return SHWD_ERR_CUDA;
  }
  return SHWD_OK;
}
#endif

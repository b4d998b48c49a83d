// Shared machinery of the entropic-OT sweeps (see sinkhorn.cu for the formulation): parameter blocks, the per-visit
// sweep with its packed / scalar inner loops, staging, merge, inter-CTA protocols and the host-side workspace carve-up.
// Included by sinkhorn.cu (flattened-deal persistent kernels) and sinkhorn_lean.cu (dedicated-CTA kernels for small
// problems); everything here is a template, inline or static, so both translation units stay self-contained.
#pragma once
#include "common.cuh"
#include "cost.cuh"
#include "f32x2.cuh"
#include <math.h>

namespace shwd {

#ifndef SHWD_THREADS
#define SHWD_THREADS 512
#endif
#ifndef SHWD_UNROLL
#define SHWD_UNROLL 1
#endif
#ifndef SHWD_OFF_UNROLL
#define SHWD_OFF_UNROLL 2
#endif
#ifndef SHWD_SPIN
#define SHWD_SPIN 1
#endif
// Adjoint planes of the backward.  The flattened-deal kernel keeps the counter protocol with two planes reused by level
// parity (hot in L2): making them write-once costs more in fresh-line write traffic than the protocol saves at the
// benchmark shape (A/B on B200 at B=32, N=1024: 8.62 ms vs 8.38 ms).  The lean kernels (sinkhorn_lean.cu) are
// latency-bound instead and poll level-indexed write-once planes; SinkParams::adj_planes says which layout a launch uses.
#ifndef SHWD_RESIDENT
#define SHWD_RESIDENT 1
#endif
#ifndef SHWD_OFFSET_LSE
#define SHWD_OFFSET_LSE 1
#endif
constexpr int SK_UNROLL = SHWD_UNROLL;
constexpr int OFF_UNROLL = SHWD_OFF_UNROLL;
constexpr int SK_THREADS = SHWD_THREADS;  // 256: two CTAs per SM (one CTA's inter-CTA wait / staging overlaps the other's compute)
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_CTAS_PER_SM = 512 / SK_THREADS;
constexpr int GMAX = 8;          // owner groups (of 32) per visit
#ifndef SHWD_CHUNK
#define SHWD_CHUNK 2048
#endif
constexpr int CHUNK = SHWD_CHUNK;  // streamed points staged per pass
constexpr int CHUNK_PAD = CHUNK + 4 * SK_WARPS;
constexpr long long WAIT_TIMEOUT_CYCLES = 6000000000LL;  // ~3 s: a lost signal ends the launch instead of hanging the GPU
constexpr float LN2F = 0.6931471805599453f;
constexpr float NEG_BIG = -3.0e38f;

enum { MODE_LSE = 0, MODE_FINAL = 1, MODE_BWD = 2 };

// Optional in-kernel phase timing (build with -DSHWD_PROFILE): thread 0 of every CTA accumulates clock64 deltas per
// phase into the workspace's err area tail; read back by tools/phase_profile.py.  Off in the product build.
#ifdef SHWD_PROFILE
#ifndef SHWD_PROF_SYM
#define SHWD_PROF_SYM g_prof
#endif
__device__ unsigned long long SHWD_PROF_SYM[8];  // one array per translation unit (no relocatable device code)
__shared__ long long s_prof_t;
#define PROF_INIT()                          \
  do {                                       \
    if (threadIdx.x == 0) s_prof_t = clock64(); \
  } while (0)
#define PROF_MARK(slot)                                                     \
  do {                                                                      \
    if (threadIdx.x == 0) {                                                 \
      long long _n = clock64();                                             \
      atomicAdd(&SHWD_PROF_SYM[slot], (unsigned long long)(_n - s_prof_t));        \
      s_prof_t = _n;                                                        \
    }                                                                       \
  } while (0)
#else
#define PROF_INIT()
#define PROF_MARK(slot)
#endif

struct SinkParams {
  const float4* X;
  const float4* Y;
  int B, N, M;
  CostParams cp;
  int iters;
  int hist_levels;
  float* alpha;  // (B, hist_levels, N)   stored iterate (what the next half-step consumes)
  float* beta;   // (B, hist_levels, M)
  float* alpha_lo;  // same shape: residual (la2 - lse2) - alpha of the float32 rounding, kept so the backward's
  float* beta_lo;   // softmax factors are normalised by the UNROUNDED log-sum-exp (see sweep())
  float* row_pc;
  float* col_pc;
  float* cost;
  int* iters_run;
  float la2, lb2, inv_k, bval;
  float thresh;
  // backward only
  const float* grad_cost;
  float4* g4x;
  float4* g4y;
  float* abar;  // (adj_planes, B, N): 2 planes by level parity, or iters+1 write-once level-indexed planes
  float* bbar;  // (adj_planes, B, M)
  int adj_planes;
  // workspace
  int* done;    // (B)
  int* status;  // (1)
  unsigned long long* err;   // (iters, B) early-stop statistic in 2^-40 fixed point (integer sums: order-independent)
  int spin_ready;  // the host pre-filled the write-once planes with SPIN_SENTINEL
};

// Per-(pair, half-step) description of one sweep.
struct SweepIO {
  const float4* own;
  int n_own;
  const float4* str;
  int n_str;
  const float* str_pot;  // nullptr -> 0
  // MODE_LSE
  float lconst;
  float* out_pot;
  float* out_pot_lo;     // nullptr -> residual not kept
  const float* old_pot;  // for the early-stop statistic (nullptr -> 0)
  unsigned long long* err_out;  // nullptr -> not recorded
  // MODE_FINAL: the plan is evaluated in its column-normalised form P_ij = b * S^v,L_ij,
  //   S^v,L_ij = 2^(fl(M(alpha^L_i) + fl(beta^L_j + lo_j - lb2))) * 2^res_j,
  // i.e. with exactly the roundings of the last beta half-step, so P and the softmax factor it cancels against in the
  // backward are the same float32 numbers.
  const float* own_pot;
  const float* own_lo;   // residual plane of own_pot (column sweep)
  int own_is_beta;       // 0: owners are x (row sums r_i); 1: owners are y (column sums c_j)
  float* out_pc;
  float pc_scale;        // bval / k: the sweep accumulates S * (k C)
  // MODE_BWD.  primary   S1 = 2^(fl(M(own_pot1) + sadd_j)) * 2^res_j,  sadd_j = fl(str_pot_j - c1), res folded into adj
  //            secondary S2 = 2^(fl(M(str_pot_j) + o2))    * 2^res2,   o2 = fl(own_pot2 - c2),     res2 folded into oadj
  const float* str_adj;  // nullptr -> 0
  float str_adj_scale;
  const float* str_lo;   // residual plane of the streamed potential (nullptr -> 0)
  float c1;
  const float* own_pot1;  // nullptr -> primary term disabled
  const float* own_pot2;  // nullptr -> secondary term disabled
  float c2;
  const float* own_lo2;
  const float* own_adj2;
  float own_adj2_scale;
  // FINAL sweeps (l = L*): the direct term g P (1 - C/eps) and the first adjoint term share S^v,L and are combined
  // analytically: weight = g (b/k) S' [1 + ln2 (k lambda_j - kC_ij)], lambda_j = c_j / b the column-mean cost; likewise
  // abar^L_i = g ln2 (b/k) sum_j S'_ij (kC_ij - k lambda_j).  No difference of separately rounded large terms is formed.
  const float* fin_cpc;   // col_pc of the pair (c_j)
  float fin_A;            // g_b * bval / k
  float fin_klam_scale;   // k / bval
  float* adj_out;  // nullptr -> not written
  float4* G;
  int G_accumulate;
};

__device__ __forceinline__ void wait_done(const int* done_b, int target, int* status) {
  if (threadIdx.x == 0 && target > 0) {
    if (ld_acquire_gpu(done_b) < target) {
      long long t0 = clock64();
      while (ld_acquire_gpu(done_b) < target) {
        if (*reinterpret_cast<volatile int*>(status) != 0) break;
        if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
          atomicExch(status, 1);
          break;
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void signal_done(int* done_b, int n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(done_b, n);
  }
}

// log2 of a positive, normal float32 as a double, accurate to ~1e-12 -- far inside what the float32 sum it is applied to
// carries.  The library's double log2 is ~150 dependent instructions on the one warp every other CTA of the pair is
// waiting for; this is ~20: exponent split, s = (f-1)/(f+1) with a float reciprocal seed and one Newton step,
// 2 atanh(s) by its odd series through s^13 (|s| <= 0.172).
__device__ __forceinline__ double fast_log2d(float x) {
  const int bits = __float_as_int(x);
  int e = ((bits >> 23) & 0xff) - 127;
  float f = __int_as_float((bits & 0x007fffff) | 0x3f800000);  // [1, 2)
  if (f > 1.41421356f) {
    f *= 0.5f;
    e += 1;
  }
  const double fd = (double)f;
  const double den = fd + 1.0;
  double r = (double)__frcp_rn((float)den);
  r = r * (2.0 - den * r);
  const double t = (fd - 1.0) * r;
  const double t2 = t * t;
  double p = 1.0 / 13.0;
  p = fma(p, t2, 1.0 / 11.0);
  p = fma(p, t2, 1.0 / 9.0);
  p = fma(p, t2, 1.0 / 7.0);
  p = fma(p, t2, 1.0 / 5.0);
  p = fma(p, t2, 1.0 / 3.0);
  p = fma(p, t2, 1.0);
  return fma(t * p, 2.8853900817779268, (double)e);  // 2 / ln 2
}

// ---- finish a visit: merge the SK_WARPS partials of every owner in fixed order, write the half-step's outputs.
// Returns true (CTA-uniform, nothing written) when a fixed-offset LSE visit has to be redone with the running maximum.
template <int MODE, bool FINAL_TERM>
__device__ __forceinline__ bool finalize_visit(const CostParams& cp, const SweepIO (&ios)[2], const int (&glo)[2], int n0, int c0v, int ng,
                                               const float4* part, const float4* sOwn, const float4* sOwn2, float* oldp_out,
                                               bool off_try) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // this thread's group (visit-local index = its warp index) and the segment it belongs to
  const int cidx = c0v + (threadIdx.x >> 5);
  const int seg = cidx >= n0;
  const SweepIO& io = ios[seg];
  const int gown = seg ? glo[1] + cidx - n0 : glo[0] + cidx;
    // ---- merge the SK_WARPS partials of every owner in fixed order and finish the half-step for these owners
    float errv = 0.f;
    float mx = NEG_BIG, sum = 0.f;
    if (MODE == MODE_LSE) {
      bool bad = false;
      if (threadIdx.x < ng * 32 && gown * 32 + lane < io.n_own) {
        const int g = threadIdx.x >> 5;
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) mx = fmaxf(mx, part[(w * GMAX + g) * 32 + lane].x);
#pragma unroll
        for (int w = 0; w < SK_WARPS; ++w) {
          float4 st = part[(w * GMAX + g) * 32 + lane];
          sum += st.y * ex2_approx(st.x - mx);
        }
        bad = !(sum >= 0x1p-60f && sum <= 0x1p60f);
      }
      if (off_try && __syncthreads_or(bad)) return true;
    }
    if (threadIdx.x < ng * 32) {
      const int g = threadIdx.x >> 5;
      const int o = gown * 32 + lane;
      if (o < io.n_own) {
        if (MODE == MODE_LSE) {
          // new potential = lconst - lse2 in double (a correctly rounded, monotone map lets the iteration settle on a
          // bitwise fixed point like the reference does -- the early-stop rule of sinkhorn.py:42-44 needs that);
          // keep the float32 rounding residual for the backward
          // (A/B at B=32, N=1024: fast_log2d here changes nothing measurable -- 4.70 vs 4.68 ms -- the merge of one visit
          //  overlaps the other warps' waits; the library log2 stays, as the early-stop rule needs its correct rounding)
          const double npd = (double)io.lconst - ((double)mx + log2((double)sum));
          const float np = (float)npd;
          if (io.out_pot_lo) io.out_pot_lo[o] = (float)(npd - (double)np);
          if (io.err_out)
            errv = fabsf(np - (sOwn ? sOwn[threadIdx.x].w : (io.old_pot ? __ldcg(io.old_pot + o) : 0.f)));
          io.out_pot[o] = np;
          if (oldp_out) oldp_out[threadIdx.x] = np;
        } else if (MODE == MODE_FINAL) {
          float sum = 0.f;
#pragma unroll
          for (int w = 0; w < SK_WARPS; ++w) sum += part[(w * GMAX + g) * 32 + lane].x;
          if (io.own_is_beta) {
            const double full = (double)__ldcg(io.own_pot + o) + (double)__ldcg(io.own_lo + o) - (double)io.c2;
            sum *= exp2f((float)(full - (double)(float)full));
          }
          io.out_pc[o] = sum * io.pc_scale;
        } else {
          float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w = 0; w < SK_WARPS; ++w) {
            float4 st = part[(w * GMAX + g) * 32 + lane];
            sum.x += st.x;
            sum.y += st.y;
            sum.z += st.z;
            sum.w += st.w;
          }
          if (io.adj_out) {
            // regular sweeps: -sum_j adj_j S_ij ; row sweep(L*): abar^L_i = -ln2 * sum_j u_ij (k lambda_j - kC_ij)
            io.adj_out[o] = (FINAL_TERM && !io.own_is_beta) ? -LN2F * sum.w : -sum.w;
          }
          float4 gv = make_float4(sum.x * cp.gscale, sum.y * cp.gscale, sum.z * cp.gscale, 0.f);
          if (io.G_accumulate) {
            const float4 old = sOwn2 ? sOwn2[threadIdx.x] : __ldcg(io.G + o);
            gv.x += old.x;
            gv.y += old.y;
            gv.z += old.z;
          }
          __stcg(io.G + o, gv);
        }
      }
    }
    if (MODE == MODE_LSE && ios[0].err_out) {
      // early-stop statistic sum_i |u_new - u_old| (sinkhorn.py:42): one atomic per warp (= per owner group), in 2^-40
      // fixed point so that the sum does not depend on the order the CTAs arrive in (an iterate sitting exactly at the
      // threshold is selected the same way in every run); a warp's share is clamped at 1024 -- astronomically above any
      // threshold -- which keeps 2048 groups inside 64 bits
      errv = warp_sum(errv);
      if (lane == 0 && threadIdx.x < ng * 32)
        atomicAdd(io.err_out, (unsigned long long)((double)fminf(errv, 1024.f) * 1099511627776.0));
    }
  __syncthreads();
  return false;
}

// ---- streamed chunk staging, scalar layout: sS[j] = (x, y, z, potential), sAdj[j] = (adjoint', addend)
template <int MODE, bool FINAL_TERM>
__device__ __forceinline__ void stage_scalar(const SweepIO& io, int c0, int cnt, int total, float4* sS, float2* sAdj) {
  for (int j = threadIdx.x; j < total; j += SK_THREADS) {
    float4 r = make_float4(0.f, 0.f, 0.f, -INFINITY);
    float2 a = make_float2(0.f, -INFINITY);
    if (MODE == MODE_BWD && FINAL_TERM && !io.own_is_beta) r.w = 0.f;  // here .w carries k*lambda_j (0 * inf = NaN otherwise)
    if (j < cnt) {
      r = __ldg(io.str + c0 + j);
      r.w = io.str_pot ? __ldcg(io.str_pot + c0 + j) : 0.f;
      if (MODE == MODE_BWD || (MODE == MODE_FINAL && !io.own_is_beta)) {
        // the streamed potential normalises the primary softmax: apply (pot + lo - c1) as a float32 addend plus a
        // multiplicative correction 2^res folded into the adjoint, so the normalisation is exact to ~1e-7
        const double full = (double)r.w + (io.str_lo ? (double)__ldcg(io.str_lo + c0 + j) : 0.0) - (double)io.c1;
        a.y = (float)full;
        const float corr = exp2f((float)(full - (double)a.y));
        if (MODE == MODE_FINAL) {
          a.x = corr;
        } else if (FINAL_TERM && !io.own_is_beta) {
          a.x = io.fin_A * corr;
          r.w = __ldcg(io.fin_cpc + c0 + j) * io.fin_klam_scale;  // k * lambda_j (the secondary term is off at l = L*)
        } else {
          a.x = io.str_adj ? __ldcg(io.str_adj + c0 + j) * io.str_adj_scale * corr : 0.f;
        }
      }
    }
    sS[j] = r;
    if (MODE != MODE_LSE) sAdj[j] = a;
  }
}

// ---- packed layout (geodesic p=2 fast path, regular sweeps): record t holds streamed points j = t ("lo" half) and
// j = t + T ("hi" half) as float2 pairs in six SoA arrays of T entries: X, Y, Z, POT (+ ADJ, ADD for the backward).
struct PackedSmem {
  float2 *X, *Y, *Z, *P, *A, *S;
};
__device__ __forceinline__ PackedSmem packed_view(float4* sS, float2* sAdj, int T) {
  PackedSmem v;
  float2* b = reinterpret_cast<float2*>(sS);
  v.X = b;
  v.Y = b + T;
  v.Z = b + 2 * T;
  v.P = b + 3 * T;
  v.A = sAdj;
  v.S = sAdj + T;
  return v;
}
// Data-flow synchronisation ("spin" mode, geodesic-p2 kernels with a history).  Potentials (forward) and adjoints
// (backward) are written ONCE per launch into level-indexed planes that the host pre-fills with the bit pattern
// 0xFFFFFFFF (a NaN no computation produces).  A consumer simply re-loads an element until it differs from the
// sentinel: the poll IS the data load, so the per-half-step chain  barrier -> fence -> atomic -> poll -> load  of the
// counter protocol collapses to one store -> load hop, and producers need no fence at all (a 32-bit store is atomic and
// becomes visible on its own; nothing else is communicated between CTAs inside these half-steps).
constexpr unsigned SPIN_SENTINEL = 0xFFFFFFFFu;
__device__ __forceinline__ unsigned ld_relaxed_u32(const float* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Staging is split around the inter-CTA wait.  PRE (before the wait): everything that does not depend on the previous
// half-step -- the streamed coordinates, and for the backward the streamed potential (forward history), its float32 addend
// and the 2^res correction.  POST (after the wait): the forward's streamed potential / the backward's streamed adjoint.
// The same thread handles the same records in both parts, so no barrier is needed between them.
template <int MODE, bool COORDS>
__device__ __forceinline__ void stage_packed_pre(const SweepIO& io, int c0, int cnt, int T, const PackedSmem& v) {
  if (!COORDS && MODE != MODE_BWD) return;  // resident coordinates: nothing to do for the forward
  float* X = reinterpret_cast<float*>(v.X);
  float* Y = reinterpret_cast<float*>(v.Y);
  float* Z = reinterpret_cast<float*>(v.Z);
  float* P = reinterpret_cast<float*>(v.P);
  float* A = reinterpret_cast<float*>(v.A);
  float* S = reinterpret_cast<float*>(v.S);
  for (int q = threadIdx.x; q < 2 * T; q += SK_THREADS) {
    const int half = q >= T, t = half ? q - T : q;
    const int j = q;  // lo half: j = t; hi half: j = t + T
    float4 r = make_float4(0.f, 0.f, 0.f, -INFINITY);
    float2 a = make_float2(0.f, -INFINITY);
    if (j < cnt) {
      if (COORDS) {
        const float4 c = __ldg(io.str + c0 + j);
        r.x = c.x;
        r.y = c.y;
        r.z = c.z;
      }
      if (MODE == MODE_BWD) {
        r.w = io.str_pot ? __ldcg(io.str_pot + c0 + j) : 0.f;
        const double full = (double)r.w + (io.str_lo ? (double)__ldcg(io.str_lo + c0 + j) : 0.0) - (double)io.c1;
        a.y = (float)full;
        a.x = exp2f((float)(full - (double)a.y));  // corr; multiplied by the adjoint in stage_packed_post
      }
    }
    const int o = 2 * t + half;
    if (COORDS) {
      X[o] = r.x;
      Y[o] = r.y;
      Z[o] = r.z;
    }
    if (MODE == MODE_BWD) {
      P[o] = r.w;
      A[o] = a.x;
      S[o] = a.y;
    }
  }
}
template <int MODE>
__device__ __forceinline__ void stage_packed_post(const SweepIO& io, int c0, int cnt, int T, const PackedSmem& v, bool spin, int* status) {
  float* dst = reinterpret_cast<float*>(MODE == MODE_LSE ? v.P : v.A);
  const float* src = (MODE == MODE_LSE) ? io.str_pot : io.str_adj;
  const float scale = (MODE == MODE_LSE) ? 1.f : io.str_adj_scale;
  // batches of 4 records per thread: the loads of a batch are all in flight before its first store
  for (int q0 = threadIdx.x; q0 < 2 * T; q0 += 4 * SK_THREADS) {
    float val[4];
    if (spin && src) {
      unsigned raw[4];
      long long t0 = 0;
      for (unsigned tries = 0;; ++tries) {
        bool all = true;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q0 + u * SK_THREADS;
          if (q < cnt && (tries == 0 || raw[u] == SPIN_SENTINEL)) raw[u] = ld_relaxed_u32(src + c0 + q);
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
      for (int u = 0; u < 4; ++u) val[u] = __uint_as_float(raw[u]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = q0 + u * SK_THREADS;
        val[u] = (src && q < cnt) ? __ldcg(src + c0 + q) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * SK_THREADS;
      if (q < 2 * T) {
        const int half = q >= T, t = half ? q - T : q;
        const int o = 2 * t + half;
        if (MODE == MODE_LSE)
          dst[o] = (q < cnt) ? val[u] : -INFINITY;
        else
          dst[o] = (src && q < cnt) ? val[u] * scale * dst[o] : 0.f;
      }
    }
  }
}

// Resident mode.  A CTA's share of a half-step type (row / column owners) is the same in every half-step, so when that
// share is one item (<= GMAX owner groups of <= 2 pairs) and the clouds fit, the packed coordinates of the streamed clouds
// and the owner records are staged ONCE per launch and stay in shared memory; a half-step then only moves potentials
// (and, in the backward, adjoints): the coordinate staging and its L2 round trip leave the critical path
// wait -> stage -> sweep -> merge -> signal that every half-step of every pair is serialised on.
struct ResidentType {
  int ok;
  int T;           // packed records per array
  // offsets, not pointers: the struct lives in local memory, and a pointer loaded from there would be generic (LD
  // instead of LDS in the sweeps); an offset added to the shared-memory base keeps the address space known
  int xyz[2];      // per segment, float2 units from the sS base: X at xyz, Y at xyz + T, Z at xyz + 2T
  int P[2];        // per segment: streamed potential
  int own;         // GMAX*32-entry slot index (0 / 1) of the owner coordinate records and owner potentials
};

// Owner records of a visit, staged once per CTA (PRE: none of it depends on the previous half-step): sOwn[g*32+lane] =
// (x, y, z, old potential of the same kind [LSE]); the backward also stages sOwn2 = the owner's accumulated gradient
// (read-modify-written by this very thread two half-steps ago) and sOwn3 = (own_pot1, o2, oadj).
template <int MODE>
__device__ __forceinline__ void stage_owners(const SweepIO (&ios)[2], const int (&glo)[2], int n0, int c0v, int ng, float4* sOwn,
                                             float4* sOwn2, float4* sOwn3, const float4* rt_ownc, const float* rt_oldp) {
  const bool rt = rt_ownc != nullptr;
  if (threadIdx.x < ng * 32) {
    const int lane = threadIdx.x & 31;
    const int cidx = c0v + (threadIdx.x >> 5);
    const int seg = cidx >= n0;
    const SweepIO& io = ios[seg];
    const int o = (seg ? glo[1] + cidx - n0 : glo[0] + cidx) * 32 + lane;
    const bool live = o < io.n_own;
    float4 rec = rt ? rt_ownc[threadIdx.x] : (live ? __ldg(io.own + o) : make_float4(0.f, 0.f, 0.f, 0.f));
    rec.w = 0.f;
    if (MODE == MODE_LSE) {
      if (rt)
        rec.w = rt_oldp[threadIdx.x];
      else if (live && io.old_pot)
        rec.w = __ldcg(io.old_pot + o);
    } else if (MODE == MODE_BWD) {
      float4 e = make_float4(-INFINITY, -INFINITY, 0.f, 0.f);
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        if (io.own_pot1) e.x = __ldcg(io.own_pot1 + o);
        if (io.own_pot2) {
          const double full = (double)__ldcg(io.own_pot2 + o) + (io.own_lo2 ? (double)__ldcg(io.own_lo2 + o) : 0.0) - (double)io.c2;
          e.y = (float)full;
          e.z = __ldcg(io.own_adj2 + o) * io.own_adj2_scale * exp2f((float)(full - (double)e.y));
        }
        if (io.G_accumulate) g = __ldcg(io.G + o);
      }
      sOwn2[threadIdx.x] = g;
      sOwn3[threadIdx.x] = e;
    }
    sOwn[threadIdx.x] = rec;
  }
}

// th = sqrt(k) * acos(c) on a packed pair -- operation for operation the scalar scaled_acos() (bit-identical results)
__device__ __forceinline__ f2 scaled_acos2(const float (&q)[7], float hpi, f2 c) {
  const f2 a = abs2(c);
  const f2 w = sub2(bc2(1.f), a);
  f2 r = bc2(q[6]);
  r = fma2(r, a, bc2(q[5]));
  r = fma2(r, a, bc2(q[4]));
  r = fma2(r, a, bc2(q[3]));
  r = fma2(r, a, bc2(q[2]));
  r = fma2(r, a, bc2(q[1]));
  r = fma2(r, a, bc2(q[0]));
  const f2 sq = mk2(sqrt_approx(fabsf(lo2(w))), sqrt_approx(fabsf(hi2(w))));
  const f2 h = fma2(neg2(sq), r, bc2(hpi));
  const f2 hs = mk2(copysignf(lo2(h), lo2(c)), copysignf(hi2(h), hi2(c)));
  return sub2(bc2(hpi), hs);
}
__device__ __forceinline__ f2 dot3_2(float ox, float oy, float oz, f2 X, f2 Y, f2 Z) {
  return fma2(bc2(oz), Z, fma2(bc2(oy), Y, mul2(bc2(ox), X)));
}

// The two fast costs in packed form.  pk_e: the per-element intermediate of Cost<FAST>::E on a packed pair (GEO2:
// th = sqrt(k) theta; SQE2: |o - s|^2); pk_m: the canonical exponent Cost<FAST>::m(e, pot), bit-identical to the scalar one.
template <int FAST>
__device__ __forceinline__ f2 pk_e(const CostParams& cp, const float4& o, f2 X, f2 Y, f2 Z) {
  if (FAST == FAST_OMC2) return fma2(bc2(-cp.sk), dot3_2(o.x, o.y, o.z, X, Y, Z), bc2(cp.sk));
  if (fast_is_geo(FAST)) return scaled_acos2(cp.q, cp.hpi, dot3_2(o.x, o.y, o.z, X, Y, Z));
  const f2 dx = sub2(bc2(o.x), X), dy = sub2(bc2(o.y), Y), dz = sub2(bc2(o.z), Z);
  if (FAST == FAST_SQE1) return add2(add2(abs2(dx), abs2(dy)), abs2(dz));
  const f2 sq = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
  if (FAST == FAST_EUC2) return mk2(sqrt_approx(lo2(sq)), sqrt_approx(hi2(sq)));
  return sq;
}
template <int FAST>
__device__ __forceinline__ f2 pk_m(const CostParams& cp, f2 e, f2 pot) {
  if (FAST == FAST_GEO2 || FAST == FAST_OMC2) return fma2(neg2(e), e, pot);
  if (FAST == FAST_GEO1) return fma2(bc2(-cp.sk), e, pot);
  return fma2(bc2(-cp.k), e, pot);
}
__host__ __device__ constexpr bool is_packed_cost(int fast) { return fast_is_geo(fast) || fast_is_sqe(fast); }
// sign(d) per half, 0 at 0 (torch.abs backward)
__device__ __forceinline__ f2 sign2(f2 d) {
  const float a = lo2(d), b = hi2(d);  // FSET (1.0 / 0.0) + LOP3 (sign copy) per half
  return mk2(copysignf(a != 0.f ? 1.f : 0.f, a), copysignf(b != 0.f ? 1.f : 0.f, b));
}

// One (R owner groups, warp-slice) pass of a regular geodesic-p2 sweep in packed arithmetic: every lane owns R points
// (one per group) and shares each streamed record between them, so one LDS.128 per array feeds 4R elements.
// tb..te (multiple of 4) is the warp's range of packed records.  Results go to the warp's partial slots (slot[r*32]),
// same format as the scalar path.
//
// LSE sweeps come in two flavours.  The safe one keeps a running maximum (online log-sum-exp).  OFF = true replaces it by
// a fixed per-owner offset: the owner's log-sum-exp of the PREVIOUS iterate, lse_old_i = lconst - pot_old_i.  The map
// streamed potential -> log-sum-exp is 1-Lipschitz in the sup norm, so once the iteration has settled a little the sum of
// 2^(m - lse_old) is close to 1.  The merge checks it: a total inside [2^-60, 2^60] is exact to float32 (terms flushed
// below 2^-126 cannot matter, an overflow makes it inf); anything else -- only the first few, wildly moving iterations --
// makes the CTA redo the visit with the running maximum.  The fixed offset removes the max / rescale work (0.25 MUFU
// and ~2 FP32/ALU slots per element) and the dependency of every ex2 on the max of its batch.
// CACHED (sinkhorn_lean.cu): the per-element intermediate pk_e -- the expensive, potential-independent part of an element
// (dot product, acos) -- of records t, t+1 against owner r comes from cth[(((t - tb) / 2) * R + r) * SK_THREADS] (the
// thread's own column of a shared-memory table filled once per launch by lean_fill_cache with the same pk_e calls, so
// the values are the ones the uncached loop would compute, bit for bit) instead of being recomputed.
template <int FAST, int MODE, int R, bool OFF = false, bool CACHED = false>
__device__ __forceinline__ void compute_packed_geo2(const CostParams& cp, float lconst, const PackedSmem& v, int tb, int te,
                                                    bool first_chunk, const float4* own, const float4* own3, float4* slot,
                                                    const float4* cth = nullptr) {
  float4 op[R];  // staged owner records of this lane (stage_owners); entries of dead owners are zero / -inf
#pragma unroll
  for (int r = 0; r < R; ++r) op[r] = own[32 * r];
  if (MODE == MODE_LSE && OFF) {
    float off[R];
    f2 rs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      off[r] = __fsub_rn(lconst, op[r].w);  // log-sum-exp of the previous iterate
      rs[r] = first_chunk ? bc2(0.f) : mk2(slot[32 * r].y, 0.f);
    }
    // software-pipelined by hand: the ex2 of batch i are issued among the dot / acos FMAs of batch i+1, so every
    // stretch of the instruction stream feeds the XU and the FMA pipe at their steady ratio
    f2 mp[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) mp[r][0] = mp[r][1] = bc2(-INFINITY);
#pragma unroll OFF_UNROLL
    for (int t = tb; t < te; t += 2) {
      const float4 X = *reinterpret_cast<const float4*>(v.X + t), Y = *reinterpret_cast<const float4*>(v.Y + t);
      const float4 Z = *reinterpret_cast<const float4*>(v.Z + t), P = *reinterpret_cast<const float4*>(v.P + t);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const f2 e0 = ex2_2(sub2(mp[r][0], bc2(off[r]))), e1 = ex2_2(sub2(mp[r][1], bc2(off[r])));
        f2 th0, th1;
        if (CACHED) {
          const float4 c = cth[(((t - tb) >> 1) * R + r) * SK_THREADS];
          th0 = mk2(c.x, c.y);
          th1 = mk2(c.z, c.w);
        } else {
          th0 = pk_e<FAST>(cp, op[r], mk2(X.x, X.y), mk2(Y.x, Y.y), mk2(Z.x, Z.y));
          th1 = pk_e<FAST>(cp, op[r], mk2(X.z, X.w), mk2(Y.z, Y.w), mk2(Z.z, Z.w));
        }
        mp[r][0] = pk_m<FAST>(cp, th0, mk2(P.x, P.y));  // the canonical exponent Cost::m, as in the safe path
        mp[r][1] = pk_m<FAST>(cp, th1, mk2(P.z, P.w));
        rs[r] = add2(rs[r], add2(e0, e1));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      rs[r] = add2(rs[r], add2(ex2_2(sub2(mp[r][0], bc2(off[r]))), ex2_2(sub2(mp[r][1], bc2(off[r])))));
#pragma unroll
    for (int r = 0; r < R; ++r) slot[32 * r] = make_float4(off[r], lo2(rs[r]) + hi2(rs[r]), 0.f, 0.f);
  } else if (MODE == MODE_LSE) {
    f2 rm[R], rs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rm[r] = bc2(NEG_BIG);
      rs[r] = bc2(0.f);
      if (!first_chunk) {
        float4 st = slot[32 * r];  // (max, sum) merged so far: restart both halves from it, the sum in the lo half only
        rm[r] = bc2(st.x);
        rs[r] = mk2(st.y, 0.f);
      }
    }
#pragma unroll SK_UNROLL
    for (int t = tb; t < te; t += 4) {
      f2 m[R][4];
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const float4 X = *reinterpret_cast<const float4*>(v.X + t + e), Y = *reinterpret_cast<const float4*>(v.Y + t + e);
        const float4 Z = *reinterpret_cast<const float4*>(v.Z + t + e), P = *reinterpret_cast<const float4*>(v.P + t + e);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          f2 th0, th1;
          if (CACHED) {
            const float4 c = cth[(((t + e - tb) >> 1) * R + r) * SK_THREADS];
            th0 = mk2(c.x, c.y);
            th1 = mk2(c.z, c.w);
          } else {
            th0 = pk_e<FAST>(cp, op[r], mk2(X.x, X.y), mk2(Y.x, Y.y), mk2(Z.x, Z.y));
            th1 = pk_e<FAST>(cp, op[r], mk2(X.z, X.w), mk2(Y.z, Y.w), mk2(Z.z, Z.w));
          }
          m[r][e] = pk_m<FAST>(cp, th0, mk2(P.x, P.y));
          m[r][e + 1] = pk_m<FAST>(cp, th1, mk2(P.z, P.w));
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const f2 nm = max2(max2(max2(m[r][0], m[r][1]), max2(m[r][2], m[r][3])), rm[r]);
        rs[r] = mul2(rs[r], ex2_2(sub2(rm[r], nm)));
        rs[r] = add2(rs[r], add2(add2(ex2_2(sub2(m[r][0], nm)), ex2_2(sub2(m[r][1], nm))),
                                 add2(ex2_2(sub2(m[r][2], nm)), ex2_2(sub2(m[r][3], nm)))));
        rm[r] = nm;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float M = fmaxf(lo2(rm[r]), hi2(rm[r]));
      const float S = lo2(rs[r]) * ex2_approx(lo2(rm[r]) - M) + hi2(rs[r]) * ex2_approx(hi2(rm[r]) - M);
      slot[32 * r] = make_float4(M, S, 0.f, 0.f);
    }
  } else {
    float opot1[R], o2[R], oadj[R];
    f2 ax[R], ay[R], az[R], aw[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 e = own3[32 * r];
      opot1[r] = e.x;
      o2[r] = e.y;
      oadj[r] = e.z;
      ax[r] = ay[r] = az[r] = aw[r] = bc2(0.f);
    }
#pragma unroll SK_UNROLL
    for (int t = tb; t < te; t += 2) {
      const float4 X = *reinterpret_cast<const float4*>(v.X + t), Y = *reinterpret_cast<const float4*>(v.Y + t);
      const float4 Z = *reinterpret_cast<const float4*>(v.Z + t), P = *reinterpret_cast<const float4*>(v.P + t);
      const float4 A = *reinterpret_cast<const float4*>(v.A + t), S = *reinterpret_cast<const float4*>(v.S + t);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const f2 x2 = h ? mk2(X.z, X.w) : mk2(X.x, X.y), y2 = h ? mk2(Y.z, Y.w) : mk2(Y.x, Y.y);
        const f2 z2 = h ? mk2(Z.z, Z.w) : mk2(Z.x, Z.y), p2 = h ? mk2(P.z, P.w) : mk2(P.x, P.y);
        const f2 a2 = h ? mk2(A.z, A.w) : mk2(A.x, A.y), s2 = h ? mk2(S.z, S.w) : mk2(S.x, S.y);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (fast_is_geo(FAST)) {
            const f2 c = dot3_2(op[r].x, op[r].y, op[r].z, x2, y2, z2);
            f2 th, gs;
            if (FAST == FAST_OMC2) {
              th = fma2(bc2(-cp.sk), c, bc2(cp.sk));  // sqrt(k) (1 - c); d(kC)/dc = -2 sqrt(k) th
              gs = th;
            } else {
              if (CACHED) {
                const float4 cc = cth[(((t - tb) >> 1) * R + r) * SK_THREADS];
                th = h ? mk2(cc.z, cc.w) : mk2(cc.x, cc.y);
              } else {
                th = scaled_acos2(cp.q, cp.hpi, c);
              }
              const f2 om = fma2(neg2(c), c, bc2(1.f));
              const f2 rs = mk2(rsqrt_approx(fmaxf(lo2(om), 1e-12f)), rsqrt_approx(fmaxf(hi2(om), 1e-12f)));
              gs = (FAST == FAST_GEO2) ? mul2(th, rs) : rs;                   // p = 1: d(k theta)/dc = -k rs (constant in gscale)
            }
            const f2 nth = (FAST == FAST_GEO1) ? bc2(-cp.sk) : neg2(th);      // -kC = nth * th in every case
            const f2 S1 = ex2_2(add2(fma2(nth, th, bc2(opot1[r])), s2));
            const f2 S2 = ex2_2(add2(fma2(nth, th, p2), bc2(o2[r])));
            const f2 w1 = mul2(a2, S1);
            aw[r] = add2(aw[r], w1);
            const f2 wg = mul2(fma2(bc2(oadj[r]), S2, w1), gs);
            ax[r] = fma2(wg, x2, ax[r]);
            ay[r] = fma2(wg, y2, ay[r]);
            az[r] = fma2(wg, z2, az[r]);
          } else {  // squared Euclidean: kC = k |o - s|^2, d(kC)/d(owner) = 2k (o - s) (2k = cp.gscale, applied per owner);
                    // L1: kC = k |o - s|_1, d(kC)/d(owner) = k sign(o - s)
            const f2 dx = sub2(bc2(op[r].x), x2), dy = sub2(bc2(op[r].y), y2), dz = sub2(bc2(op[r].z), z2);
            f2 sq = (FAST == FAST_SQE1) ? add2(add2(abs2(dx), abs2(dy)), abs2(dz)) : fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
            f2 ri = bc2(1.f);
            if (FAST == FAST_EUC2) {  // kC = k |d|, d(kC)/d(owner) = k d / |d| (0 at d = 0); |d| = sqrt.approx as in the forward
              ri = mk2(lo2(sq) > 0.f ? rsqrt_approx(lo2(sq)) : 0.f, hi2(sq) > 0.f ? rsqrt_approx(hi2(sq)) : 0.f);
              sq = mk2(sqrt_approx(lo2(sq)), sqrt_approx(hi2(sq)));
            }
            const f2 nk = bc2(-cp.k);
            const f2 S1 = ex2_2(add2(fma2(nk, sq, bc2(opot1[r])), s2));
            const f2 S2 = ex2_2(add2(fma2(nk, sq, p2), bc2(o2[r])));
            const f2 w1 = mul2(a2, S1);
            aw[r] = add2(aw[r], w1);
            f2 wg = fma2(bc2(oadj[r]), S2, w1);
            if (FAST == FAST_EUC2) wg = mul2(wg, ri);
            ax[r] = fma2(wg, (FAST == FAST_SQE1) ? sign2(dx) : dx, ax[r]);
            ay[r] = fma2(wg, (FAST == FAST_SQE1) ? sign2(dy) : dy, ay[r]);
            az[r] = fma2(wg, (FAST == FAST_SQE1) ? sign2(dz) : dz, az[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float4 acc = first_chunk ? make_float4(0.f, 0.f, 0.f, 0.f) : slot[32 * r];
      acc.x += lo2(ax[r]) + hi2(ax[r]);
      acc.y += lo2(ay[r]) + hi2(ay[r]);
      acc.z += lo2(az[r]) + hi2(az[r]);
      acc.w += lo2(aw[r]) + hi2(aw[r]);
      slot[32 * r] = acc;
    }
  }
}

// One item of a half-step: up to two SEGMENTS -- owner-group ranges [glo[s], ghi[s]) of two different pairs (a CTA whose
// share straddles a pair boundary) -- swept together: both pairs' streamed data are staged side by side, all groups are
// computed in one pass and finalised together, so a straddling CTA pays the fixed per-item cost (staging, barriers,
// finalise, signal) once instead of twice (every other CTA of both pairs waits for it each half-step).

// acquire: every segment's pair has finished the previous half-step
__device__ __forceinline__ void wait_done2(const int* done, int b0, int b1, int nseg, int target, int* status) {
  if (threadIdx.x == 0 && target > 0) {
    for (int s = 0; s < nseg; ++s) {
      const int* d = done + (s ? b1 : b0);
      if (ld_acquire_gpu(d) < target) {
        long long t0 = clock64();
        while (ld_acquire_gpu(d) < target) {
          if (*reinterpret_cast<volatile int*>(status) != 0) break;
          if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
            atomicExch(status, 1);
            break;
          }
        }
      }
    }
  }
  __syncthreads();
}

struct WaitSpec {
  const int* done;
  int b0, b1;
  int target;
  int* status;
  int try_off;  // forward LSE: attempt the fixed-offset sum (see compute_packed_geo2)
  int spin;     // consumers poll the data itself (see above); done/target are not used by this item
};

template <int FAST, int MODE, bool FINAL_TERM>
__device__ void sweep(const CostParams& cp, SweepIO (&ios)[2], const int (&glo)[2], const int (&ghi)[2], int nseg, float4* sS0,
                      float2* sAdj0, float4* part, float4* sOwn, const WaitSpec& ws, const ResidentType* rt) {
  typedef Cost<FAST> CF;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr bool PACKED = is_packed_cost(FAST) && (MODE == MODE_LSE || (MODE == MODE_BWD && !FINAL_TERM));
  constexpr int PK = is_packed_cost(FAST) ? FAST : FAST_GEO2;  // (the packed branch is dead code for the generic kernel)
  const int n0 = ghi[0] - glo[0];
  const int ntot = n0 + (nseg > 1 ? ghi[1] - glo[1] : 0);
  float4* sSb[2] = {sS0, sS0 + CHUNK_PAD};
  float2* sAdjb[2] = {sAdj0, sAdj0 + CHUNK_PAD};
  const int n_str = ios[0].n_str;  // every pair of a launch has the same cloud sizes
  float4* sOwn2 = sOwn + GMAX * 32;
  float4* sOwn3 = sOwn2 + GMAX * 32;
  // resident arrays (see resident_setup for the carve-up)
  float2* sS2 = reinterpret_cast<float2*>(sS0);
  const float4* rt_ownc = rt ? sOwn + (3 + rt->own) * GMAX * 32 : nullptr;
  float* rt_oldp = rt ? reinterpret_cast<float*>(sOwn + 5 * GMAX * 32) + rt->own * GMAX * 32 : nullptr;
  bool waited = false;
  if (!PACKED) {
    wait_done2(ws.done, ws.b0, ws.b1, nseg, ws.target, ws.status);
    PROF_MARK(0);
    waited = true;
  }
  for (int c0v = 0; c0v < ntot; c0v += GMAX) {
    const int ng = min(GMAX, ntot - c0v);
    const bool use0 = c0v < n0, use1 = (c0v + ng) > n0;  // which segments this visit touches
    bool off_try = PACKED && MODE == MODE_LSE && ws.try_off;
    for (;;) {  // (a fixed-offset LSE visit whose sums left the safe range is redone once with the running maximum)
    for (int c0 = 0; c0 < n_str; c0 += CHUNK) {
      const int cnt = min(CHUNK, n_str - c0);
      if (PACKED) {
        // per-warp slice of packed records (two streamed points each), multiple of 4
        const int SLt = ((((cnt + 1) / 2 + SK_WARPS - 1) / SK_WARPS) + 3) & ~3;
        const int T = SLt * SK_WARPS;
        PackedSmem pv0 = packed_view(sSb[0], sAdjb[0], T), pv1 = packed_view(sSb[1], sAdjb[1], T);
        if (rt) {
          pv0.X = sS2 + rt->xyz[0];
          pv0.Y = pv0.X + T;
          pv0.Z = pv0.Y + T;
          pv0.P = sS2 + rt->P[0];
          pv1.X = sS2 + rt->xyz[1];
          pv1.Y = pv1.X + T;
          pv1.Z = pv1.Y + T;
          pv1.P = sS2 + rt->P[1];
          if (use0) stage_packed_pre<MODE, false>(ios[0], c0, cnt, T, pv0);
          if (use1) stage_packed_pre<MODE, false>(ios[1], c0, cnt, T, pv1);
        } else {
          if (use0) stage_packed_pre<MODE, true>(ios[0], c0, cnt, T, pv0);
          if (use1) stage_packed_pre<MODE, true>(ios[1], c0, cnt, T, pv1);
        }
        if (c0 == 0) stage_owners<MODE>(ios, glo, n0, c0v, ng, sOwn, sOwn2, sOwn3, rt_ownc, rt_oldp);
        if (!waited) {
          PROF_MARK(1);
          if (!ws.spin) wait_done2(ws.done, ws.b0, ws.b1, nseg, ws.target, ws.status);
          PROF_MARK(0);
          waited = true;
        }
        if (use0) stage_packed_post<MODE>(ios[0], c0, cnt, T, pv0, ws.spin != 0, ws.status);
        if (use1) stage_packed_post<MODE>(ios[1], c0, cnt, T, pv1, ws.spin != 0, ws.status);
        __syncthreads();
        PROF_MARK(6);
        for (int g = 0; g < ng;) {  // two owner groups of the same pair per pass share every streamed record
          const int seg = (c0v + g) >= n0;
          const int gown = seg ? glo[1] + c0v + g - n0 : glo[0] + c0v + g;
          const bool two = (g + 1 < ng) && (((c0v + g + 1) >= n0) == (seg != 0));
          float4* slot = part + (warp * GMAX + g) * 32 + lane;
          if (MODE == MODE_LSE && off_try) {
            if (two)
              compute_packed_geo2<PK, MODE, 2, true>(cp, ios[seg].lconst, seg ? pv1 : pv0, warp * SLt, warp * SLt + SLt, c0 == 0, sOwn + g * 32 + lane, sOwn3 + g * 32 + lane, slot);
            else
              compute_packed_geo2<PK, MODE, 1, true>(cp, ios[seg].lconst, seg ? pv1 : pv0, warp * SLt, warp * SLt + SLt, c0 == 0, sOwn + g * 32 + lane, sOwn3 + g * 32 + lane, slot);
          } else if (two)
            compute_packed_geo2<PK, MODE, 2>(cp, ios[seg].lconst, seg ? pv1 : pv0, warp * SLt, warp * SLt + SLt, c0 == 0, sOwn + g * 32 + lane, sOwn3 + g * 32 + lane, slot);
          else
            compute_packed_geo2<PK, MODE, 1>(cp, ios[seg].lconst, seg ? pv1 : pv0, warp * SLt, warp * SLt + SLt, c0 == 0, sOwn + g * 32 + lane, sOwn3 + g * 32 + lane, slot);
          g += two ? 2 : 1;
        }
      } else {
      const int SL = (((cnt + SK_WARPS - 1) / SK_WARPS) + 3) & ~3;  // per-warp slice, multiple of 4
      if (use0) stage_scalar<MODE, FINAL_TERM>(ios[0], c0, cnt, SL * SK_WARPS, sSb[0], sAdjb[0]);
      if (use1) stage_scalar<MODE, FINAL_TERM>(ios[1], c0, cnt, SL * SK_WARPS, sSb[1], sAdjb[1]);
      __syncthreads();
      PROF_MARK(1);
      const int j0 = warp * SL, j1 = j0 + SL;
      for (int g = 0; g < ng; ++g) {
        const int seg = (c0v + g) >= n0;
        const SweepIO& io = ios[seg];
        const float4* sS = sS0 + seg * CHUNK_PAD;
        const float2* sAdj = sAdj0 + seg * CHUNK_PAD;
        const int o = (seg ? glo[1] + c0v + g - n0 : glo[0] + c0v + g) * 32 + lane;
        const bool live = o < io.n_own;
        float4 op = live ? __ldg(io.own + o) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4* slot = part + (warp * GMAX + g) * 32 + lane;
        if (MODE == MODE_LSE) {
          float rm = NEG_BIG, rs = 0.f;
          if (c0 > 0) {
            float4 st = *slot;
            rm = st.x;
            rs = st.y;
          }
#pragma unroll 2
          for (int j = j0; j < j1; j += 4) {
            float4 s0 = sS[j], s1 = sS[j + 1], s2 = sS[j + 2], s3 = sS[j + 3];
            float m0 = CF::m(cp, CF::eval(cp, op.x, op.y, op.z, s0.x, s0.y, s0.z), s0.w);
            float m1 = CF::m(cp, CF::eval(cp, op.x, op.y, op.z, s1.x, s1.y, s1.z), s1.w);
            float m2 = CF::m(cp, CF::eval(cp, op.x, op.y, op.z, s2.x, s2.y, s2.z), s2.w);
            float m3 = CF::m(cp, CF::eval(cp, op.x, op.y, op.z, s3.x, s3.y, s3.z), s3.w);
            float nm = fmaxf(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), rm);
            rs *= ex2_approx(rm - nm);
            rs += (ex2_approx(m0 - nm) + ex2_approx(m1 - nm)) + (ex2_approx(m2 - nm) + ex2_approx(m3 - nm));
            rm = nm;
          }
          *slot = make_float4(rm, rs, 0.f, 0.f);
        } else if (MODE == MODE_FINAL) {
          float acc = (c0 > 0) ? slot->x : 0.f;
          float opot = live ? __ldcg(io.own_pot + o) : -INFINITY;
          const bool col = io.own_is_beta != 0;
          if (col && live) opot = (float)((double)opot + (double)__ldcg(io.own_lo + o) - (double)io.c2);  // o2 of S^v,L
#pragma unroll 2
          for (int j = j0; j < j1; j += 4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float4 s = sS[j + e];
              typename CF::E ce = CF::eval(cp, op.x, op.y, op.z, s.x, s.y, s.z);
              float S;
              if (col) {
                S = ex2_approx(__fadd_rn(CF::m(cp, ce, s.w), opot));
              } else {
                float2 ad = sAdj[j + e];
                S = ex2_approx(__fadd_rn(CF::m(cp, ce, opot), ad.y)) * ad.x;
              }
              acc = fmaf(S, CF::kc(cp, ce), acc);
            }
          }
          *slot = make_float4(acc, 0.f, 0.f, 0.f);
        } else {
          float4 acc = (c0 > 0) ? *slot : make_float4(0.f, 0.f, 0.f, 0.f);
          float opot1 = -INFINITY, o2 = -INFINITY, oadj = 0.f, oklam = 0.f;
          const bool col = io.own_is_beta != 0;
          if (live) {
            if (io.own_pot1) opot1 = __ldcg(io.own_pot1 + o);
            if (io.own_pot2) {
              const double full = (double)__ldcg(io.own_pot2 + o) + (io.own_lo2 ? (double)__ldcg(io.own_lo2 + o) : 0.0) - (double)io.c2;
              o2 = (float)full;
              const float corr = exp2f((float)(full - (double)o2));
              if (FINAL_TERM && col) {
                oadj = io.fin_A * corr;
                oklam = __ldcg(io.fin_cpc + o) * io.fin_klam_scale;
              } else {
                oadj = __ldcg(io.own_adj2 + o) * io.own_adj2_scale * corr;
              }
            }
          }
#pragma unroll 2
          for (int j = j0; j < j1; j += 2) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float4 s = sS[j + e];
              float2 ad = sAdj[j + e];
              float gs, vx, vy, vz;
              typename CF::E ce = CF::eval_grad(cp, op.x, op.y, op.z, s.x, s.y, s.z, gs, vx, vy, vz);
              float S1 = ex2_approx(__fadd_rn(CF::m(cp, ce, opot1), ad.y));
              float wt;
              if (FINAL_TERM && !col) {
                // row sweep(L*): u = g (b/k) S'_ij ; weight u [1 + ln2 (k lambda_j - kC)] ; abar accumulates u (kC - k lambda_j)
                float u = ad.x * S1;
                float d = __fsub_rn(s.w, CF::kc(cp, ce));
                acc.w = fmaf(u, d, acc.w);
                wt = fmaf(u * LN2F, d, u);
              } else {
                float S2 = ex2_approx(__fadd_rn(CF::m(cp, ce, s.w), o2));
                float w1 = ad.x * S1;
                acc.w += w1;
                if (FINAL_TERM) {  // col sweep(L*): secondary + direct term combined the same way
                  float u = oadj * S2;
                  float d = __fsub_rn(oklam, CF::kc(cp, ce));
                  wt = fmaf(u * LN2F, d, u) + w1;
                } else {
                  wt = fmaf(oadj, S2, w1);
                }
              }
              float wg = wt * gs;
              acc.x = fmaf(wg, vx, acc.x);
              acc.y = fmaf(wg, vy, acc.y);
              acc.z = fmaf(wg, vz, acc.z);
            }
          }
          *slot = acc;
        }
      }
      }
      __syncthreads();
      PROF_MARK(2);
    }
    const bool redo = finalize_visit<MODE, FINAL_TERM>(cp, ios, glo, n0, c0v, ng, part, PACKED ? sOwn : nullptr,
                                                       PACKED ? sOwn2 : nullptr, (PACKED && MODE == MODE_LSE) ? rt_oldp : nullptr, off_try);
    PROF_MARK(3);
    if (!redo) break;
    off_try = false;
    }
  }
}

// release: this CTA's groups of every segment are done
__device__ __forceinline__ void signal_done2(int* done, const int (&segb)[2], const int (&glo)[2], const int (&ghi)[2], int nseg) {
  __syncthreads();
  if (threadIdx.x == 0) {
    // release at gpu scope (cumulative over the bar.sync above); __threadfence() would be the heavier fence.sc.gpu
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    atomicAdd(done + segb[0], ghi[0] - glo[0]);
    if (nseg > 1) atomicAdd(done + segb[1], ghi[1] - glo[1]);
  }
}

__device__ __forceinline__ void cta_range(long long total, int& g0, int& g1) {
  g0 = (int)((total * blockIdx.x) / gridDim.x);
  g1 = (int)((total * (blockIdx.x + 1)) / gridDim.x);
}

// Work split of one half-step: the flattened (pair, owner group) list is dealt to the CTAs in contiguous, balanced runs
// (cta_range).  (A two-set interleave that lets a CTA alternate between two halves of the batch to hide the inter-CTA
// wait was tried and lost: it doubles the per-item fixed cost -- staging, merge, signal -- which outweighs the wait.)

// One-time staging for the resident mode (see ResidentType).  sS (4*CHUNK_PAD float2 slots) is carved into the packed
// coordinate arrays of both types' streamed clouds plus one potential array per segment.
template <int FAST>
__device__ void resident_setup(const SinkParams& prm, int gr, int gc, float4* sS, float4* sOwnC, float* sOldP, ResidentType (&R)[2]) {
  int g0[2], g1[2], T[2], ok[2], xyz[2][2], Pof[2];
  long long need = 0;
  int Tmax = 0;
#pragma unroll
  for (int type = 0; type < 2; ++type) {
    const int gpp = type ? gc : gr, n_str = type ? prm.N : prm.M;
    cta_range((long long)prm.B * gpp, g0[type], g1[type]);
    const int ntot = g1[type] - g0[type];
    const int npairs = ntot > 0 ? (g1[type] - 1) / gpp - g0[type] / gpp + 1 : 0;
    const int SLt = ((((n_str + 1) / 2 + SK_WARPS - 1) / SK_WARPS) + 3) & ~3;
    T[type] = SLt * SK_WARPS;
    ok[type] = is_packed_cost(FAST) && SHWD_RESIDENT && ntot >= 1 && ntot <= GMAX && npairs <= 2 && n_str <= CHUNK;
    need += 2LL * 3 * T[type];
    Tmax = max(Tmax, T[type]);
  }
  need += 2LL * Tmax;
  if (need > 4LL * CHUNK_PAD) ok[0] = ok[1] = 0;
  float2* sS2 = reinterpret_cast<float2*>(sS);
  int base = 0;
#pragma unroll
  for (int type = 0; type < 2; ++type) {
#pragma unroll
    for (int sg = 0; sg < 2; ++sg) {
      xyz[type][sg] = base;
      base += 3 * T[type];
    }
  }
  Pof[0] = base;
  Pof[1] = base + Tmax;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int type = 0; type < 2; ++type) {
      R[type].ok = ok[type];
      R[type].T = T[type];
      R[type].own = type;
      R[type].xyz[0] = xyz[type][0];
      R[type].xyz[1] = xyz[type][1];
      R[type].P[0] = Pof[0];
      R[type].P[1] = Pof[1];
    }
  }
  __syncthreads();  // the previous users of sS are done
#pragma unroll
  for (int type = 0; type < 2; ++type) {
    if (!ok[type]) continue;
    const int gpp = type ? gc : gr, n_str = type ? prm.N : prm.M, n_own = type ? prm.M : prm.N;
    const float4* str = type ? prm.X : prm.Y;
    const float4* own = type ? prm.Y : prm.X;
    const int b0 = g0[type] / gpp;
    const int TT = T[type];
#pragma unroll
    for (int sg = 0; sg < 2; ++sg) {
      const int b = b0 + sg;
      if ((long long)b * gpp >= g1[type]) break;
      float* X = reinterpret_cast<float*>(sS2 + xyz[type][sg]);
      float* Y = X + 2 * TT;
      float* Z = Y + 2 * TT;
      for (int q = threadIdx.x; q < 2 * TT; q += SK_THREADS) {
        const int half = q >= TT, t = half ? q - TT : q;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < n_str) c = __ldg(str + (size_t)b * n_str + q);
        const int o = 2 * t + half;
        X[o] = c.x;
        Y[o] = c.y;
        Z[o] = c.z;
      }
    }
    if (threadIdx.x < GMAX * 32) {
      const int g = g0[type] + (threadIdx.x >> 5);
      float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < g1[type]) {
        const int b = g / gpp, o = (g % gpp) * 32 + (threadIdx.x & 31);
        if (o < n_own) rec = __ldg(own + (size_t)b * n_own + o);
      }
      rec.w = 0.f;
      sOwnC[type * GMAX * 32 + threadIdx.x] = rec;
      sOldP[type * GMAX * 32 + threadIdx.x] = 0.f;
    }
  }
  __syncthreads();
}

// ---- forward tail (shared by both forward kernels): pick the iterate, the two FINAL sweeps, the cost.
// lse_done = per-pair counter value once every LSE half-step of the pair is published.
template <int FAST>
__device__ void fwd_tail(const SinkParams& prm, int lse_done, float4* sS, float2* sAdj, float4* part, float4* sOwn,
                         SweepIO (&s_ios)[2], int& s_ls) {
  const int gr = (prm.N + 31) / 32, gc = (prm.M + 31) / 32;
  const int HL = prm.hist_levels;
  const int L = prm.iters;
  auto slot = [&](int l) { return HL > 1 ? l : 0; };
  // ---- which iterate is the result?  (sinkhorn.py:42-44: first l with mean_b sum_i |u^l - u^{l-1}| < thresh)
  int Ls = L;
  if (prm.thresh > 0.f) {
    for (int b = 0; b < prm.B; ++b) wait_done(prm.done + b, lse_done, prm.status);
    if (threadIdx.x == 0) {
      int found = L;
      for (int l = 0; l < L; ++l) {
        unsigned long long si = 0ull;
        for (int b = 0; b < prm.B; ++b) si += __ldcg(prm.err + (size_t)l * prm.B + b);
        const float s = (float)((double)si * 9.094947017729282e-13);
        // err is in alpha units (k u); the reference tests u
        if (s * prm.inv_k / prm.B < prm.thresh) {
          found = l + 1;
          break;
        }
      }
      s_ls = found;
    }
    __syncthreads();
    Ls = s_ls;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *prm.iters_run = Ls;

  // ---- final sweeps: r_i = sum_j P_ij C_ij (row owners), c_j = sum_i P_ij C_ij (col owners)
  for (int type = 0; type < 2; ++type) {
    const int gpp = type ? gc : gr;
    int gbeg, gend;
    cta_range((long long)prm.B * gpp, gbeg, gend);
    for (int g = gbeg; g < gend;) {
      // gather up to two segments (a CTA whose share straddles a pair boundary) into one item
      SweepIO (&ios)[2] = s_ios;
      int segb[2] = {0, 0}, seg0[2] = {0, 0}, seg1[2] = {0, 0}, nseg = 0;
      while (nseg < 2 && g < gend) {
      const int b = g / gpp, lg0 = g % gpp;
      const int lg1 = min(gpp, lg0 + (gend - g));
      segb[nseg] = b;
      seg0[nseg] = lg0;
      seg1[nseg] = lg1;
      g += lg1 - lg0;
      if (threadIdx.x == 32 * nseg) {
      SweepIO& io = s_ios[nseg];
      io.pc_scale = prm.bval * prm.inv_k;
      const float* al = prm.alpha + ((size_t)b * HL + slot(Ls)) * prm.N;
      const float* be = prm.beta + ((size_t)b * HL + slot(Ls)) * prm.M;
      const float* be_lo = prm.beta_lo + ((size_t)b * HL + slot(Ls)) * prm.M;
      io.c1 = io.c2 = prm.lb2;
      io.str_lo = nullptr;
      io.own_lo = nullptr;
      if (type == 0) {
        io.own = prm.X + (size_t)b * prm.N;
        io.n_own = prm.N;
        io.str = prm.Y + (size_t)b * prm.M;
        io.n_str = prm.M;
        io.str_pot = be;
        io.str_lo = be_lo;
        io.own_pot = al;
        io.own_is_beta = 0;
        io.out_pc = prm.row_pc + (size_t)b * prm.N;
      } else {
        io.own = prm.Y + (size_t)b * prm.M;
        io.n_own = prm.M;
        io.str = prm.X + (size_t)b * prm.N;
        io.n_str = prm.N;
        io.str_pot = al;
        io.own_pot = be;
        io.own_lo = be_lo;
        io.own_is_beta = 1;
        io.out_pc = prm.col_pc + (size_t)b * prm.M;
      }
      }
      ++nseg;
      }
      __syncthreads();  // the item descriptors (written by thread 0) are visible to the CTA
      PROF_MARK(5);
      WaitSpec ws = {prm.done, segb[0], segb[1], lse_done, prm.status, 0, 0};
      sweep<FAST, MODE_FINAL, false>(prm.cp, ios, seg0, seg1, nseg, sS, sAdj, part, sOwn, ws, nullptr);
      signal_done2(prm.done, segb, seg0, seg1, nseg);
      PROF_MARK(4);
    }
  }

  // ---- cost_b = sum_i r_i, summed in a fixed order by the CTA that owns the pair's first row group
  {
    int g0, g1;
    cta_range((long long)prm.B * gr, g0, g1);
    for (int b = (g0 + gr - 1) / gr; b * gr < g1 && b < prm.B; ++b) {
      if (b * gr < g0) continue;
      wait_done(prm.done + b, lse_done + gr + gc, prm.status);
      float s = 0.f;
      for (int i = threadIdx.x; i < prm.N; i += SK_THREADS) s += __ldcg(prm.row_pc + (size_t)b * prm.N + i);
      s = warp_sum(s);
      __shared__ float sc[SK_WARPS];
      if ((threadIdx.x & 31) == 0) sc[threadIdx.x >> 5] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < SK_WARPS; ++w) t += sc[w];
        prm.cost[b] = t;
      }
      __syncthreads();
    }
  }
}

// ---- one phase of the backward in the flattened deal (shared: the lean backward runs phases 0 and 1 -- the FINAL
// sweeps -- through this code).  spin: consumers poll write-once adjoint planes (never with two parity planes).
template <int FAST>
__device__ void bwd_phase(const SinkParams& prm, int ph, int Ls, bool spin, float4* sS, float2* sAdj, float4* part, float4* sOwn,
                          SweepIO (&s_ios)[2], ResidentType (&RT)[2]) {
  const int gr = (prm.N + 31) / 32, gc = (prm.M + 31) / 32;
  const int HL = prm.hist_levels;
  const size_t BN = (size_t)prm.B * prm.N, BM = (size_t)prm.B * prm.M;
  auto ADJ_PLANE = [&](int l) { return prm.adj_planes == 2 ? (l & 1) : l; };
  const bool last = (ph == 2 * Ls);
  const int type = last ? 0 : (ph & 1);
  const int l = last ? 0 : Ls - (ph >> 1);
  const int gpp = type ? gc : gr;
  const int nrow_before = last ? Ls : ((ph + 1) >> 1), ncol_before = last ? Ls : (ph >> 1);
  int gbeg, gend;
  cta_range((long long)prm.B * gpp, gbeg, gend);
  for (int g = gbeg; g < gend;) {
    // gather up to two segments (a CTA whose share straddles a pair boundary) into one item
    SweepIO (&ios)[2] = s_ios;
    int segb[2] = {0, 0}, seg0[2] = {0, 0}, seg1[2] = {0, 0}, nseg = 0;
    while (nseg < 2 && g < gend) {
    const int b = g / gpp, lg0 = g % gpp;
    const int lg1 = min(gpp, lg0 + (gend - g));
    segb[nseg] = b;
    seg0[nseg] = lg0;
    seg1[nseg] = lg1;
    g += lg1 - lg0;
    if (threadIdx.x == 32 * nseg) {
    SweepIO& io = s_ios[nseg];
    const float gb = __ldg(prm.grad_cost + b);
    const float* al = prm.alpha + (size_t)b * HL * prm.N;  // level 0
    const float* be = prm.beta + (size_t)b * HL * prm.M;
    const float* al_lo = prm.alpha_lo + (size_t)b * HL * prm.N;
    const float* be_lo = prm.beta_lo + (size_t)b * HL * prm.M;
    io.fin_A = gb * prm.bval * prm.inv_k;
    io.fin_cpc = prm.col_pc + (size_t)b * prm.M;
    io.fin_klam_scale = 1.f / (prm.bval * prm.inv_k);
    io.own_is_beta = type;
    io.str_lo = nullptr;
    io.own_lo2 = nullptr;
    io.G_accumulate = (l != Ls);
    if (type == 0) {
      io.own = prm.X + (size_t)b * prm.N;
      io.n_own = prm.N;
      io.str = prm.Y + (size_t)b * prm.M;
      io.n_str = prm.M;
      io.str_pot = be + (size_t)l * prm.M;
      io.G = prm.g4x + (size_t)b * prm.N;
      io.c1 = prm.lb2;
      io.c2 = prm.la2;
      if (last) {
        io.str_adj = nullptr;
        io.str_adj_scale = 0.f;
        io.own_pot1 = nullptr;
        io.adj_out = nullptr;
      } else {
        if (l == Ls) {
          io.str_adj = nullptr;  // FINAL row sweep: the streamed scalars are g (b/k) 2^res_j and k lambda_j
          io.str_adj_scale = 0.f;
        } else {
          io.str_adj = prm.bbar + (size_t)ADJ_PLANE(l) * BM + (size_t)b * prm.M;
          io.str_adj_scale = 1.f;
        }
        io.own_pot1 = al + (size_t)l * prm.N;
        io.str_lo = be_lo + (size_t)l * prm.M;  // S^v,l is normalised by beta^l (streamed)
        io.adj_out = prm.abar + (size_t)ADJ_PLANE(l) * BN + (size_t)b * prm.N;
      }
      if (l < Ls) {
        io.own_pot2 = al + (size_t)(l + 1) * prm.N;
        io.own_lo2 = al_lo + (size_t)(l + 1) * prm.N;  // S^u,l+1 is normalised by alpha^{l+1} (owner)
        io.own_adj2 = prm.abar + (size_t)ADJ_PLANE(l + 1) * BN + (size_t)b * prm.N;
        io.own_adj2_scale = 1.f;
      } else {
        io.own_pot2 = nullptr;
        io.own_adj2 = nullptr;
        io.own_adj2_scale = 0.f;
      }
    } else {
      io.own = prm.Y + (size_t)b * prm.M;
      io.n_own = prm.M;
      io.str = prm.X + (size_t)b * prm.N;
      io.n_str = prm.N;
      io.str_pot = al + (size_t)l * prm.N;
      io.str_adj = prm.abar + (size_t)ADJ_PLANE(l) * BN + (size_t)b * prm.N;
      io.str_adj_scale = 1.f;
      io.str_lo = al_lo + (size_t)l * prm.N;  // S^u,l is normalised by alpha^l (streamed)
      io.G = prm.g4y + (size_t)b * prm.M;
      io.own_pot1 = be + (size_t)(l - 1) * prm.M;
      io.c1 = prm.la2;
      io.own_pot2 = be + (size_t)l * prm.M;
      io.own_lo2 = be_lo + (size_t)l * prm.M;  // S^v,l is normalised by beta^l (owner)
      io.c2 = prm.lb2;
      if (l == Ls) {
        io.own_adj2 = nullptr;  // FINAL col sweep: the owner scalars are g (b/k) 2^res_j and k lambda_j
        io.own_adj2_scale = 0.f;
      } else {
        io.own_adj2 = prm.bbar + (size_t)ADJ_PLANE(l) * BM + (size_t)b * prm.M;
        io.own_adj2_scale = 1.f;
      }
      io.adj_out = prm.bbar + (size_t)ADJ_PLANE(l - 1) * BM + (size_t)b * prm.M;
    }
    }
    ++nseg;
    }
    __syncthreads();  // the item descriptors (written by thread 0) are visible to the CTA
    PROF_MARK(5);
    // spin mode: phase 0 publishes through the counters (phase 1 stages through the scalar path); from phase 2 on the
    // consumers poll the adjoints themselves
    WaitSpec ws = {prm.done, segb[0], segb[1], nrow_before * gr + ncol_before * gc, prm.status, 0, (spin && ph >= 2) ? 1 : 0};
    if (l == Ls)
      sweep<FAST, MODE_BWD, true>(prm.cp, ios, seg0, seg1, nseg, sS, sAdj, part, sOwn, ws, nullptr);
    else
      sweep<FAST, MODE_BWD, false>(prm.cp, ios, seg0, seg1, nseg, sS, sAdj, part, sOwn, ws, RT[type].ok ? &RT[type] : nullptr);
    if (!spin || ph == 0) signal_done2(prm.done, segb, seg0, seg1, nseg);
    PROF_MARK(4);
  }
}

// dynamic shared memory of every OT kernel (the general carve-up; the lean kernels alias a smaller one inside it)
static inline size_t sinkhorn_smem_bytes() {
  return sizeof(float4) * (2 * CHUNK_PAD + (size_t)SK_WARPS * GMAX * 32 + 5 * GMAX * 32) + sizeof(float2) * 2 * CHUNK_PAD +
         sizeof(float) * 2 * GMAX * 32;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  int* done;
  int* status;
  unsigned long long* err;
  float* abar;
  float* bbar;
  size_t head_bytes;  // counters + status + err (memset on every launch)
  size_t total;
};

static Workspace carve(void* base, int B, int N, int M, int iters, int adj_planes) {
  Workspace w;
  size_t off = 0;
  char* p = static_cast<char*>(base);
  w.status = reinterpret_cast<int*>(p + off);
  off += 256;
  w.done = reinterpret_cast<int*>(p + off);
  off = align_up(off + sizeof(int) * (size_t)B, 256);
  w.err = reinterpret_cast<unsigned long long*>(p + off);
  off = align_up(off + sizeof(unsigned long long) * (size_t)B * (size_t)iters, 256);
  w.head_bytes = off;
  w.abar = reinterpret_cast<float*>(p + off);
  off = align_up(off + sizeof(float) * (size_t)adj_planes * B * N, 256);
  w.bbar = reinterpret_cast<float*>(p + off);
  off = align_up(off + sizeof(float) * (size_t)adj_planes * B * M, 256);
  w.total = off;
  return w;
}

static inline void fill_marginals(SinkParams& prm, int N, int M, float eps) {
  // log(fill_(1.0/n) + 1e-8) in float32, as the reference builds it (sinkhorn.py:25-26,39-40)
  const float a = (float)(1.0 / (double)N) + 1e-8f, b = (float)(1.0 / (double)M) + 1e-8f;
  prm.la2 = (float)log2((double)a);
  prm.lb2 = (float)log2((double)b);
  prm.bval = b;
  prm.inv_k = (float)((double)eps / 1.4426950408889634);
}

template <typename K>
static int launch_persistent(K kernel, const SinkParams& prm, size_t smem, int max_groups, cudaStream_t s) {
  static thread_local int configured_dev = -1;
  (void)configured_dev;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return SHWD_ERR_CUDA;
  }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SK_THREADS, smem);
  if (e != cudaSuccess || per_sm < 1) {
    set_last_cuda_error(e == cudaSuccess ? cudaErrorLaunchOutOfResources : e);
    return SHWD_ERR_CUDA;
  }
  int grid = sm_count() * (per_sm < SK_CTAS_PER_SM ? per_sm : SK_CTAS_PER_SM);  // persistent CTAs, all co-resident
  if (grid > max_groups) grid = max_groups;
  if (grid < 1) grid = 1;
  void* args[] = {const_cast<SinkParams*>(&prm)};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(grid), dim3(SK_THREADS), args, smem, s);
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return SHWD_ERR_CUDA;
  }
  return SHWD_OK;
}

}  // namespace shwd

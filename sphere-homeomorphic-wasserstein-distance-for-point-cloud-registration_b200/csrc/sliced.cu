// Sliced paths: projection -> per-slice stable radix sort in shared memory -> fused 1-D (circular) Wasserstein
// reductions, with the gradients w.r.t. the sorted values and the scatter back through the sort permutation.
//
// Replaces   sliced_cost projection + normalise + atan2     Point_Cloud_Resistration/losses/max_spherical_sliced_w.py:270-279
//            torch.sort calls                               max_spherical_sliced_w.py:163-164, 224-225, 232, 235
//            emd1D_circle (p = 1, level median)             max_spherical_sliced_w.py:230-247
//            Euclidean sliced W                             Wasserstein_flow_problem/Flow_ellipsoid.ipynb:208-220 (cell 5)
//
// Sort contract: the permutation equals torch.sort(keys, dim=-1, stable=True) bit for bit -- ascending, ties in input
// order, -0.0 == +0.0, NaN last.  Keys go through the usual order-preserving float->uint32 map
// (with -0 canonicalised and every NaN mapped to 0xFFFFFFFF) and a 4-pass LSD radix sort (8-bit digits) whose
// ranking is warp-synchronous: each warp owns a contiguous item range, counts its digits with shared-memory atomics on a
// warp-private histogram, then ranks 32 consecutive items per step by grouping equal digits with eight ballots, so the
// scatter needs no atomics and a pass has only three block barriers.  Records ping-pong between two shared-memory buffers:
// (key, index) uint2 pairs for rows up to 4224 keys, separate 4-byte key / 2-byte index arrays up to 16384 keys
// (segmented_sort_compact_kernel), and between two global scratch buffers beyond that.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

namespace shwd {

constexpr int PJ_THREADS = 256;
constexpr float TWO_PI_F = 6.283185307179586f;
constexpr float PI_F = 3.141592653589793f;

// ------------------------------------------------------------------------------------------------ projections ----
// One key: the stand-alone projection kernels and the sort kernels that compute their own keys (segmented_sort_project_*)
// call the same functions, so both produce the same bits.
//
// circle coordinate t = (atan2(-c, -a) + pi) / (2 pi) of the projected point (a, c).  The reference normalises (a, c) first
// (F.normalize, eps 1e-12) -- atan2 does not depend on the scale, so the square root and the two divisions are skipped -- and
// calls atan2f (libdevice: ~45 instructions with an IEEE division inside).  Here: q = min / max of the magnitudes,
// atan q = q P8(q^2) (degree-8 interpolant at Chebyshev nodes of [0, 1], 1.2e-8 rad), the octant / sign fix-ups of
// IEEE atan2 on the sign BITS (so +-0 behave as in atan2f: (0, 0) -> t = 0, 0.5 or 1 like the reference), NaN / inf
// coordinates -> NaN like normalize gives.  Against the exact value on 2e6 random points: max 7.9e-8, mean 1.4e-8; the
// reference's own float32 chain (torch CPU): max 1.0e-7, mean 1.8e-8 (tools/fit_circle_key.py).  The key was 40 % of the
// fused projection + sort kernel's instructions (ncu source page, r02h) before this.
__device__ __forceinline__ float circle_key(const float* u /* U[p][d][k] at d*2+k */, float x0, float x1, float x2) {
  const float a = fmaf(u[4], x2, fmaf(u[2], x1, u[0] * x0));
  const float c = fmaf(u[5], x2, fmaf(u[3], x1, u[1] * x0));
  const float x = -a, y = -c;
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float q = mx > 0.f ? __fdividef(mn, mx) : 0.f;
  const float s2 = q * q;
  float p = 0.0028340641874819994f;
  p = fmaf(p, s2, -0.016005029901862144f);
  p = fmaf(p, s2, 0.042587608098983765f);
  p = fmaf(p, s2, -0.07495445758104324f);
  p = fmaf(p, s2, 0.10636754333972931f);
  p = fmaf(p, s2, -0.14202570915222168f);
  p = fmaf(p, s2, 0.19992484152317047f);
  p = fmaf(p, s2, -0.3333306610584259f);
  p = fmaf(p, s2, 1.0f);
  float r = p * q;
  if (ay > ax) r = 1.5707963267948966f - r;
  if (__float_as_uint(x) >> 31) r = PI_F - r;
  r = copysignf(r, y);
  return fmaf(r, 0.15915494309189535f, 0.5f) + 0.f * (a + c);  // (the last term: NaN for NaN / inf coordinates)
}
__device__ __forceinline__ float line_key(const float* t /* theta[p][d] */, float x0, float x1, float x2) {
  return fmaf(t[2], x2, fmaf(t[1], x1, t[0] * x0));
}

// keys[b,p,n] = (atan2(-q1, -q0) + pi) / (2 pi), q = normalize(U_p^T x_n)            (sliced_cost :270-279)
__global__ void __launch_bounds__(PJ_THREADS) project_circle_kernel(const float* __restrict__ x, const float* __restrict__ U,
                                                                    int N, int P, float* __restrict__ keys, int pp) {
  // pp: U holds one frame set PER PAIR, (B,P,3,2) (max_spherical_sliced_w_fast.py:298-319), instead of one (P,3,2) for all
  extern __shared__ float sU[];  // P_TILE * 6
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * 32;
  const int pc = min(32, P - p0);
  if (pp) U += (size_t)b * P * 6;
  for (int i = threadIdx.x; i < pc * 6; i += PJ_THREADS) sU[i] = __ldg(U + (size_t)p0 * 6 + i);
  __syncthreads();
  const int n = blockIdx.x * PJ_THREADS + threadIdx.x;
  if (n >= N) return;
  const float* xp = x + ((size_t)b * N + n) * 3;
  const float x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2);
#pragma unroll 4
  for (int p = 0; p < pc; ++p) keys[((size_t)b * P + p0 + p) * N + n] = circle_key(sU + p * 6, x0, x1, x2);
}

// gx[b,n,:] = sum_p gk[b,p,n] * ( dt/da * U[p,:,0] + dt/dc * U[p,:,1] ),  dt/da = -c / (2 pi r^2), dt/dc = a / (2 pi r^2)
// A CTA owns 32 consecutive points (lane = point); its 8 warps split the slices (warp w takes p = w, w + 8, ...), so a row
// read is one 128-B line per warp and B * N / 32 CTAs x 8 independent load streams are in flight (the first version walked
// all P slices sequentially in one thread per point: 128 CTAs, latency-bound, 250 us at cfg3).  The 8 partial sums are
// added in warp order through shared memory (deterministic).
constexpr int PB_WARPS = PJ_THREADS / 32;
constexpr int PB_TILE = 256;  // slices staged per tile

__global__ void __launch_bounds__(PJ_THREADS) project_circle_bwd_kernel(const float* __restrict__ x, const float* __restrict__ U,
                                                                        int N, int P, const float* __restrict__ gk,
                                                                        const float* __restrict__ gw, float* __restrict__ gx, int pp) {
  // gw (nullable): per-pair upstream gradient of the slice MEAN -- the result is scaled by gw[b] / P here instead of by
  // an elementwise pass afterwards (same two roundings: the division, then the product)
  __shared__ float sU[PB_TILE * 6];
  __shared__ float red[PB_WARPS][3][32];
  const int b = blockIdx.y;
  if (pp) U += (size_t)b * P * 6;  // one frame set per pair
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool ok = n < N;
  float x0 = 0.f, x1 = 0.f, x2 = 0.f;
  if (ok) {
    const float* xp = x + ((size_t)b * N + n) * 3;
    x0 = __ldg(xp);
    x1 = __ldg(xp + 1);
    x2 = __ldg(xp + 2);
  }
  const float* gkb = gk + (size_t)b * P * N + (ok ? n : 0);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  for (int p0 = 0; p0 < P; p0 += PB_TILE) {
    const int pc = min(PB_TILE, P - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < pc * 6; i += PJ_THREADS) sU[i] = __ldg(U + (size_t)p0 * 6 + i);
    __syncthreads();
    // sixteen rows in flight per lane, the loads before the arithmetic (same accumulation order)
    for (int pb = warp; pb < pc; pb += 16 * PB_WARPS) {
      float gq[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int p = pb + q * PB_WARPS;
        gq[q] = (ok && p < pc) ? __ldg(gkb + (size_t)(p0 + p) * N) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int p = pb + q * PB_WARPS;
        if (p < pc) {
          const float* u = sU + p * 6;
          float a = fmaf(u[4], x2, fmaf(u[2], x1, u[0] * x0));
          float c = fmaf(u[5], x2, fmaf(u[3], x1, u[1] * x0));
          float r2 = fmaxf(fmaf(c, c, a * a), 1e-24f);
          float g = gq[q] / (TWO_PI_F * r2);
          float ta = -c * g, tc = a * g;
          g0 = fmaf(ta, u[0], fmaf(tc, u[1], g0));
          g1 = fmaf(ta, u[2], fmaf(tc, u[3], g1));
          g2 = fmaf(ta, u[4], fmaf(tc, u[5], g2));
        }
      }
    }
  }
  red[warp][0][lane] = g0;
  red[warp][1][lane] = g1;
  red[warp][2][lane] = g2;
  __syncthreads();
  if (threadIdx.x < 96) {  // output element (point pt, component k) of this CTA's contiguous 96-float span
    const int pt = threadIdx.x / 3, k = threadIdx.x - pt * 3;
    if (blockIdx.x * 32 + pt < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < PB_WARPS; ++w) t += red[w][k][pt];
      if (gw) t = t * (__ldg(gw + b) / (float)P);
      gx[((size_t)b * N + blockIdx.x * 32) * 3 + threadIdx.x] = t;
    }
  }
}

// The same with FOUR consecutive points per lane (N % 4 == 0): a warp reads 512 contiguous bytes of a gradient row instead of
// 128 (one LDG.128 per lane), a CTA owns 128 points.  The rows are N * 4 bytes apart, so the narrow version's requests were
// isolated 128-byte lines -- DRAM pages opened for one line each: 0.96 TB/s at cfg3 (67 MB per cloud in 70 us).  Per point the
// slices are visited by the same warp in the same order and the eight partial sums are added in the same warp order, so the
// result is the narrow kernel's, bit for bit.
__global__ void __launch_bounds__(PJ_THREADS) project_circle_bwd4_kernel(const float* __restrict__ x, const float* __restrict__ U,
                                                                         int N, int P, const float* __restrict__ gk,
                                                                         const float* __restrict__ gw, float* __restrict__ gx, int pp) {
  __shared__ float sU[PB_TILE * 6];
  __shared__ float red[PB_WARPS][3][128];
  const int b = blockIdx.y;
  if (pp) U += (size_t)b * P * 6;  // one frame set per pair
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4 = (blockIdx.x * 32 + lane) * 4;
  const bool ok = n4 < N;  // N % 4 == 0: a lane's four points are all inside or all outside
  float X[4][3];
#pragma unroll
  for (int e = 0; e < 4; ++e) X[e][0] = X[e][1] = X[e][2] = 0.f;
  if (ok) {
    const float4* xp = reinterpret_cast<const float4*>(x + ((size_t)b * N + n4) * 3);
    const float4 A = __ldg(xp), Bv = __ldg(xp + 1), C = __ldg(xp + 2);
    X[0][0] = A.x; X[0][1] = A.y; X[0][2] = A.z;
    X[1][0] = A.w; X[1][1] = Bv.x; X[1][2] = Bv.y;
    X[2][0] = Bv.z; X[2][1] = Bv.w; X[2][2] = C.x;
    X[3][0] = C.y; X[3][1] = C.z; X[3][2] = C.w;
  }
  const float* gkb = gk + (size_t)b * P * N + (ok ? n4 : 0);
  float G[4][3];
#pragma unroll
  for (int e = 0; e < 4; ++e) G[e][0] = G[e][1] = G[e][2] = 0.f;
  for (int p0 = 0; p0 < P; p0 += PB_TILE) {
    const int pc = min(PB_TILE, P - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < pc * 6; i += PJ_THREADS) sU[i] = __ldg(U + (size_t)p0 * 6 + i);
    __syncthreads();
    // eight rows in flight per lane (the loads first, then the arithmetic): the kernel is bound by load latency at 1.7 CTAs per SM
    for (int pb = warp; pb < pc; pb += 8 * PB_WARPS) {
      float4 gq[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int p = pb + q * PB_WARPS;
        gq[q] = (ok && p < pc) ? __ldg(reinterpret_cast<const float4*>(gkb + (size_t)(p0 + p) * N)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int p = pb + q * PB_WARPS;
        if (p < pc) {
          const float* u = sU + p * 6;
          const float gkv[4] = {gq[q].x, gq[q].y, gq[q].z, gq[q].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a = fmaf(u[4], X[e][2], fmaf(u[2], X[e][1], u[0] * X[e][0]));
            float c = fmaf(u[5], X[e][2], fmaf(u[3], X[e][1], u[1] * X[e][0]));
            float r2 = fmaxf(fmaf(c, c, a * a), 1e-24f);
            float g = gkv[e] / (TWO_PI_F * r2);
            float ta = -c * g, tc = a * g;
            G[e][0] = fmaf(ta, u[0], fmaf(tc, u[1], G[e][0]));
            G[e][1] = fmaf(ta, u[2], fmaf(tc, u[3], G[e][1]));
            G[e][2] = fmaf(ta, u[4], fmaf(tc, u[5], G[e][2]));
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    red[warp][0][lane * 4 + e] = G[e][0];
    red[warp][1][lane * 4 + e] = G[e][1];
    red[warp][2][lane * 4 + e] = G[e][2];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 384; o += PJ_THREADS) {  // output element (point pt, component k) of this CTA's contiguous span
    const int pt = o / 3, k = o - pt * 3;
    if (blockIdx.x * 128 + pt < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < PB_WARPS; ++w) t += red[w][k][pt];
      if (gw) t = t * (__ldg(gw + b) / (float)P);
      gx[((size_t)b * N + blockIdx.x * 128) * 3 + o] = t;
    }
  }
}

// keys[b,p,n] = <x_n, theta_p>                                              (Flow_ellipsoid.ipynb:214-216)
__global__ void __launch_bounds__(PJ_THREADS) project_line_kernel(const float* __restrict__ x, const float* __restrict__ th, int N,
                                                                  int P, float* __restrict__ keys) {
  extern __shared__ float sT[];
  const int b = blockIdx.z;
  const int p0 = blockIdx.y * 32;
  const int pc = min(32, P - p0);
  for (int i = threadIdx.x; i < pc * 3; i += PJ_THREADS) sT[i] = __ldg(th + (size_t)p0 * 3 + i);
  __syncthreads();
  const int n = blockIdx.x * PJ_THREADS + threadIdx.x;
  if (n >= N) return;
  const float* xp = x + ((size_t)b * N + n) * 3;
  const float x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2);
  for (int p = 0; p < pc; ++p) keys[((size_t)b * P + p0 + p) * N + n] = line_key(sT + p * 3, x0, x1, x2);
}

__global__ void __launch_bounds__(PJ_THREADS) project_line_bwd_kernel(const float* __restrict__ th, int N, int P,
                                                                      const float* __restrict__ gk, const float* __restrict__ gw,
                                                                      float* __restrict__ gx) {
  __shared__ float sT[PB_TILE * 3];
  __shared__ float red[PB_WARPS][3][32];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool ok = n < N;
  const float* gkb = gk + (size_t)b * P * N + (ok ? n : 0);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  for (int p0 = 0; p0 < P; p0 += PB_TILE) {
    const int pc = min(PB_TILE, P - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < pc * 3; i += PJ_THREADS) sT[i] = __ldg(th + (size_t)p0 * 3 + i);
    __syncthreads();
#pragma unroll 8
    for (int p = warp; p < pc; p += PB_WARPS) {
      const float g = ok ? __ldg(gkb + (size_t)(p0 + p) * N) : 0.f;
      g0 = fmaf(g, sT[p * 3], g0);
      g1 = fmaf(g, sT[p * 3 + 1], g1);
      g2 = fmaf(g, sT[p * 3 + 2], g2);
    }
  }
  red[warp][0][lane] = g0;
  red[warp][1][lane] = g1;
  red[warp][2][lane] = g2;
  __syncthreads();
  if (threadIdx.x < 96) {
    const int pt = threadIdx.x / 3, k = threadIdx.x - pt * 3;
    if (blockIdx.x * 32 + pt < N) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < PB_WARPS; ++w) t += red[w][k][pt];
      if (gw) t = t * (__ldg(gw + b) / (float)P);
      gx[((size_t)b * N + blockIdx.x * 32) * 3 + threadIdx.x] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------ radix sort -----
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_SMEM_MAX = 8192;  // items per segment whose ping-pong buffers fit in shared memory

// Integer-only (immune to -ftz): every NaN -> 0xFFFFFFFF (last), -0.0 -> +0.0 (the two zeros tie), then the usual
// order-preserving map.
__device__ __forceinline__ uint32_t float_sort_key(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;
  if ((u << 1) == 0u) u = 0u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_sort_key(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// Block-wide exclusive scan of one unsigned value per thread (SORT_THREADS threads); returns the exclusive prefix.
__device__ __forceinline__ uint32_t block_exscan_u32(uint32_t v, uint32_t* warp_tot /* SORT_WARPS */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  uint32_t base = 0;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  __syncthreads();
  return base + inc - v;
}

// Lanes holding the same 9-bit value (8-bit digit + "invalid" flag), from nine ballots.  __match_any_sync gives the same mask
// but is microcoded on sm_100 (hundreds of cycles on the ADU pipe: it made this kernel ADU-bound, ncu 84 % busy).
template <int BITS>
__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
  uint32_t peers = 0xffffffffu;
  // peers &= (bit set) ? ballot : ~ballot -- spelled in PTX so that it stays ~3 instructions per bit (R2P sets the
  // predicates of four bits at once, then vote, predicated complement, and); the compiler's own lowering of the C++
  // form takes six.  The bits are fed four at a time from separately shifted copies so that ptxas does not try to keep
  // all eight predicates alive at once (it has seven).
#pragma unroll
  for (int base = 0; base < BITS; base += 4) {
    uint32_t dd;
    asm volatile("shr.u32 %0, %1, %2;" : "=r"(dd) : "r"(d), "r"(base));
#pragma unroll
    for (int bit = 0; bit < 4 && base + bit < BITS; ++bit) {
      asm volatile(
          "{\n .reg .pred p;\n .reg .b32 t, m;\n and.b32 t, %1, %2;\n setp.ne.u32 p, t, 0;\n"
          " vote.sync.ballot.b32 m, p, 0xffffffff;\n @!p not.b32 m, m;\n and.b32 %0, %0, m;\n}"
          : "+r"(peers)
          : "r"(dd), "r"(1u << bit));
    }
  }
  return peers;
}

// Record buffers of the sort live in shared memory (segments up to 8192 keys) or in global scratch.  In the shared case
// they are addressed through 32-bit shared-window addresses with explicit ld.shared / st.shared: through generic pointers
// the compiler emitted LD.E.64 / ST.E.64 plus, for every predicated counter store, an S2R SR_SWINHI / SR_CgaCtaId pair to
// rebuild the window base inside the batch loop (the kernel is ALU / issue bound, so those instructions were not free).
template <bool SMEM>
struct RecBuf {
  uint2* g;     // global (or generic) pointer
  uint32_t sa;  // shared-window byte address (SMEM only)
};
template <bool SMEM>
__device__ __forceinline__ RecBuf<SMEM> make_recbuf(uint2* p) {
  RecBuf<SMEM> r;
  r.g = p;
  r.sa = SMEM ? (uint32_t)__cvta_generic_to_shared(p) : 0u;
  return r;
}
template <bool SMEM>
__device__ __forceinline__ uint2 rec_ld(const RecBuf<SMEM>& b, int i) {
  if (SMEM) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(b.sa + 8u * (uint32_t)i));
    return v;
  }
  return b.g[i];
}
template <bool SMEM>
__device__ __forceinline__ uint32_t rec_ld_key(const RecBuf<SMEM>& b, int i) {
  if (SMEM) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(b.sa + 8u * (uint32_t)i));
    return v;
  }
  return b.g[i].x;
}
template <bool SMEM>
__device__ __forceinline__ void rec_st(const RecBuf<SMEM>& b, int i, uint2 v) {
  if (SMEM)
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(b.sa + 8u * (uint32_t)i), "r"(v.x), "r"(v.y) : "memory");
  else
    b.g[i] = v;
}
// shared-memory store under a predicate, without a branch (32-bit shared-window address)
__device__ __forceinline__ void st_shared_if(uint32_t saddr, uint32_t v, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.u32 [%0], %1; }" ::"r"(saddr), "r"(v), "r"((int)pred) : "memory");
}

// Stable LSD radix sort of n (key, index) records: a -> ... -> result buffer returned (either a or b).
// hist: SORT_WARPS*256 words of shared memory; wt: SORT_WARPS words.
template <bool SMEM>
__device__ RecBuf<SMEM> block_radix_sort(RecBuf<SMEM> a, RecBuf<SMEM> b, int n, uint32_t* hist, uint32_t* wt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = (((n + SORT_WARPS - 1) / SORT_WARPS) + 31) & ~31;
  const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t* wh = hist + warp * 256;
  const uint32_t wh_s = (uint32_t)__cvta_generic_to_shared(wh);
  for (int pass = 0; pass < 4; ++pass) {
    const int sh = pass * 8;
    for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) hist[i] = 0;
    __syncthreads();
    // The per-warp digit counters are read by every lane and advanced by the leader of each match group with a PREDICATED
    // store: no divergent branch inside the batch loops (branch / reconvergence bookkeeping runs on the ADU pipe, which
    // was the busiest unit of this kernel), one __syncwarp per batch orders the store against the next batch's reads.
    // counting needs no order: shared-memory atomics on the warp's private histogram
    for (int i = beg + lane; i < end; i += 32) atomicAdd(wh + ((rec_ld_key<SMEM>(a, i) >> sh) & 255u), 1u);
    __syncthreads();
    {
      // digit-major exclusive offsets: thread d owns digit d
      const int d = threadIdx.x;
      uint32_t tot = 0;
#pragma unroll
      for (int w = 0; w < SORT_WARPS; ++w) tot += hist[w * 256 + d];
      uint32_t run = block_exscan_u32(tot, wt);
#pragma unroll
      for (int w = 0; w < SORT_WARPS; ++w) {
        uint32_t t = hist[w * 256 + d];
        hist[w * 256 + d] = run;
        run += t;
      }
    }
    __syncthreads();
    int i0 = beg;
#pragma unroll 1
    for (; i0 + 32 <= end; i0 += 32) {  // full batches: every lane holds a record, eight ballots
      const uint2 rec = rec_ld<SMEM>(a, i0 + lane);
      const uint32_t d = (rec.x >> sh) & 255u;
      const uint32_t peers = match_digit<8>(d);
      const uint32_t cur = wh[d];
      rec_st<SMEM>(b, cur + __popc(peers & lt), rec);
      st_shared_if(wh_s + 4u * d, cur + __popc(peers), (peers & lt) == 0);
      __syncwarp();
    }
    if (i0 < end) {  // the ragged tail of the warp's range: a ninth "invalid" bit keeps the empty lanes apart
      const int i = i0 + lane;
      const bool valid = i < end;
      const uint2 rec = valid ? rec_ld<SMEM>(a, i) : make_uint2(0u, 0u);
      const uint32_t d = valid ? ((rec.x >> sh) & 255u) : 256u;
      const uint32_t peers = match_digit<9>(d);
      const uint32_t cur = wh[d & 255u];
      if (valid) rec_st<SMEM>(b, cur + __popc(peers & lt), rec);
      st_shared_if(wh_s + 4u * (d & 255u), cur + __popc(peers), valid && (peers & lt) == 0);
      __syncwarp();
    }
    __syncthreads();
    RecBuf<SMEM> t = a;
    a = b;
    b = t;
  }
  return a;
}

template <bool SMEM>
__global__ void __launch_bounds__(SORT_THREADS) segmented_sort_kernel(const float* __restrict__ keys, int len,
                                                                      float* __restrict__ sorted, int64_t* __restrict__ perm,
                                                                      int32_t* __restrict__ perm32, uint2* __restrict__ gscratch) {
  extern __shared__ uint2 sbuf[];
  __shared__ uint32_t hist[SORT_WARPS * 256];
  __shared__ uint32_t wt[SORT_WARPS];
  const size_t seg = blockIdx.x;
  const float* k = keys + seg * len;
  RecBuf<SMEM> a, b;
  if (SMEM) {
    a = make_recbuf<SMEM>(sbuf);
    b = make_recbuf<SMEM>(sbuf + len);
  } else {
    a = make_recbuf<SMEM>(gscratch + seg * 2 * (size_t)len);
    b = make_recbuf<SMEM>(gscratch + seg * 2 * (size_t)len + len);
  }
  for (int i = threadIdx.x; i < len; i += SORT_THREADS) rec_st<SMEM>(a, i, make_uint2(float_sort_key(__ldg(k + i)), (uint32_t)i));
  __syncthreads();
  const RecBuf<SMEM> r = block_radix_sort<SMEM>(a, b, len, hist, wt);
  for (int i = threadIdx.x; i < len; i += SORT_THREADS) {
    const uint32_t j = rec_ld<SMEM>(r, i).y;
    if (sorted) sorted[seg * len + i] = __ldg(k + j);
    if (perm) perm[seg * len + i] = (int64_t)j;
    if (perm32) perm32[seg * len + i] = (int32_t)j;
  }
}

// ---- digit trimming + fused projection (the sliced losses' own sort) ---------------------------------------------------
// An LSD radix sort only has to look at the bits in which the keys of a row differ.  Circle coordinates lie in [0, 1): sign
// and the top exponent bits are the same for every key of a row, and a row of 4096 coordinates almost always differs in 27
// bits -- three passes of 9-bit digits instead of four of 8 (the bits are found per row from the OR / AND of its keys, so
// any input is still sorted exactly; signed Euclidean projections keep four passes).  A pass is the same warp-synchronous
// ranking as block_radix_sort, with per-warp digit counters as 16-bit halves (512 bins x 8 warps = the same 8 KB).
// The kernel can also COMPUTE its keys: a slice's keys are a function of the (N,3) cloud and one frame, so the sort CTA of
// slice (b, p) projects the cloud itself instead of reading a (B,P,N) key array that a projection kernel wrote -- two
// launches and 8 B per key of HBM traffic less per cloud -- and rebuilds the sorted values from the sorted keys (the key map
// is invertible; -0.0 comes back as +0.0, which no consumer of the sorted values can tell apart) instead of gathering them.
constexpr int SORT_TRIM_MAXW = 9;

template <int W>
__device__ __forceinline__ void radix_pass_trim(const RecBuf<true>& a, const RecBuf<true>& b, int beg, int end, int sh, uint32_t dmask,
                                                uint32_t* hist32, uint32_t* wt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int BINS = 1 << SORT_TRIM_MAXW;          // counters are laid out for the widest digit
  constexpr int WORDS = BINS / 2;                    // 16-bit counters, two per word
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t* wh = hist32 + warp * WORDS;
  const uint32_t wh_s = (uint32_t)__cvta_generic_to_shared(wh);
  for (int i = threadIdx.x; i < SORT_WARPS * WORDS; i += SORT_THREADS) hist32[i] = 0;
  __syncthreads();
  for (int i = beg + lane; i < end; i += 32) {
    const uint32_t d = (rec_ld_key<true>(a, i) >> sh) & dmask;
    atomicAdd(wh + (d >> 1), 1u << ((d & 1u) * 16u));
  }
  __syncthreads();
  {
    // digit-major exclusive offsets: thread t owns digits 2t and 2t+1 (one counter word per warp)
    const int t = threadIdx.x;
    uint32_t c[SORT_WARPS];
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      c[w] = hist32[w * WORDS + t];
      tot += (c[w] & 0xffffu) + (c[w] >> 16);
    }
    uint32_t run = block_exscan_u32(tot, wt);
    uint32_t lo[SORT_WARPS];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      lo[w] = run;
      run += c[w] & 0xffffu;
    }
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      hist32[w * WORDS + t] = lo[w] | (run << 16);
      run += c[w] >> 16;
    }
  }
  __syncthreads();
  int i0 = beg;
#pragma unroll 1
  for (; i0 + 32 <= end; i0 += 32) {
    const uint2 rec = rec_ld<true>(a, i0 + lane);
    const uint32_t d = (rec.x >> sh) & dmask;
    const uint32_t peers = match_digit<W>(d);
    uint32_t cur;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(cur) : "r"(wh_s + 2u * d));
    rec_st<true>(b, cur + __popc(peers & lt), rec);
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.u16 [%0], %1; }" ::"r"(wh_s + 2u * d), "r"(cur + __popc(peers)),
                 "r"((int)((peers & lt) == 0))
                 : "memory");
    __syncwarp();
  }
  if (i0 < end) {  // ragged tail: an extra "invalid" bit keeps the empty lanes apart
    const int i = i0 + lane;
    const bool valid = i < end;
    const uint2 rec = valid ? rec_ld<true>(a, i) : make_uint2(0u, 0u);
    const uint32_t d = valid ? ((rec.x >> sh) & dmask) : (1u << W);
    const uint32_t peers = match_digit<W + 1>(d);
    const uint32_t dd = d & dmask;
    uint32_t cur;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(cur) : "r"(wh_s + 2u * dd));
    if (valid) rec_st<true>(b, cur + __popc(peers & lt), rec);
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.u16 [%0], %1; }" ::"r"(wh_s + 2u * dd), "r"(cur + __popc(peers)),
                 "r"((int)(valid && (peers & lt) == 0))
                 : "memory");
    __syncwarp();
  }
  __syncthreads();
}

// ---- one-pass bucket sort for rows whose keys are spread out (what projections of a cloud are) -------------------------
// torch.sort(stable=True) orders a row by the pair (key, input index), a total order: any method that realises it is "the
// stable sort".  The keys of a slice are ~evenly spread over their range, so a MONOTONE map of the value onto SORT_NB buckets,
//     bucket(t) = min(floor((t - tmin) * SORT_NB / (tmax - tmin)), SORT_NB - 1)        (float32 ops are monotone),
// leaves about one record per bucket: count (shared-memory atomics on 16-bit halves), scan, scatter into the bucket's range
// in ARRIVAL order (atomicAdd returning the slot), then every record counts the records of its own bucket that precede it
// in (key, index) order -- a loop of the bucket's length, ~2 on average -- and moves to its final place.  Four light sweeps
// over the row instead of three ranking passes of nine ballots each (cfg3: 280 -> see DESIGN.md 4.4).  Rows with a non-finite
// key or a bucket above SORT_BUCKET_CAP records (ties, clustered keys) take the radix passes below; the decision is made
// before anything is moved.
constexpr int SORT_NB = 2 * SORT_WARPS * (1 << SORT_TRIM_MAXW) / 2;  // 4096 buckets: the radix counters' 8 KB as 16-bit halves
constexpr int SORT_BUCKET_CAP = 48;
constexpr int SORT_BUCKET_MIN_LEN = 128;

__device__ __forceinline__ uint32_t sort_bucket(uint32_t key, float tmin, float scale) {
  const float t = float_from_sort_key(key);
  return min(__float2uint_rz(__fmul_rn(__fsub_rn(t, tmin), scale)), (uint32_t)(SORT_NB - 1));
}

// Returns false (CTA-uniform, nothing moved) if the row must take the radix passes; true: the row's outputs are written
// (straight from the rank step: a record's final slot lies within a bucket's length of the slot it is read from, so a
// warp's stores still fall into the same few 128-byte lines).  COUNTED: the caller filled the histogram while it produced
// the keys (circle coordinates: the map of [0, 1) is known beforehand -- tmin = 0, scale = SORT_NB).
template <bool COUNTED>
__device__ __forceinline__ bool bucket_sort_row(const RecBuf<true>& a, const RecBuf<true>& b, int len, uint32_t kmin, uint32_t kmax,
                                                uint32_t* hist32, uint32_t* wt, float* __restrict__ sorted_row,
                                                int32_t* __restrict__ perm_row) {
  constexpr int WORDS = SORT_NB / 2;
  constexpr int WPT = WORDS / SORT_THREADS;  // counter words per thread in the scan
  if (len < SORT_BUCKET_MIN_LEN || kmax >= 0xFF800000u || kmin <= 0x007FFFFFu) return false;  // short row / +-inf / NaN
  float tmin = 0.f, scale = (float)SORT_NB;
  if (!COUNTED) {
    tmin = float_from_sort_key(kmin);
    scale = __fdiv_rn((float)SORT_NB, __fsub_rn(float_from_sort_key(kmax), tmin));
    if (!(scale < 3.0e38f)) return false;  // range below ~1e-35: every key would land in the two end buckets anyway
    for (int i = threadIdx.x; i < WORDS; i += SORT_THREADS) hist32[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += SORT_THREADS) {
      const uint32_t bk = sort_bucket(rec_ld_key<true>(a, i), tmin, scale);
      atomicAdd(hist32 + (bk >> 1), 1u << ((bk & 1u) * 16u));
    }
    __syncthreads();
  } else if (kmin < 0x80000000u) {
    return false;  // a negative key: not circle coordinates after all
  }
  const uint32_t hs = (uint32_t)__cvta_generic_to_shared(hist32);
  uint32_t c[WPT];
  uint32_t tot = 0, mx = 0;
#pragma unroll
  for (int w = 0; w < WPT; ++w) {
    c[w] = hist32[threadIdx.x * WPT + w];
    tot += (c[w] & 0xffffu) + (c[w] >> 16);
    mx = max(mx, max(c[w] & 0xffffu, c[w] >> 16));
  }
  if (__syncthreads_or(mx > (uint32_t)SORT_BUCKET_CAP)) return false;
  uint32_t run = block_exscan_u32(tot, wt);
#pragma unroll
  for (int w = 0; w < WPT; ++w) {  // counters -> first slot of each bucket (the scatter advances them to the bucket's end)
    const uint32_t lo = run, hi = run + (c[w] & 0xffffu);
    hist32[threadIdx.x * WPT + w] = lo | (hi << 16);
    run = hi + (c[w] >> 16);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += SORT_THREADS) {
    const uint2 rec = rec_ld<true>(a, i);
    const uint32_t bk = sort_bucket(rec.x, tmin, scale);
    const uint32_t old = atomicAdd(hist32 + (bk >> 1), 1u << ((bk & 1u) * 16u));
    rec_st<true>(b, (int)((old >> ((bk & 1u) * 16u)) & 0xffffu), rec);
  }
  __syncthreads();
  for (int sl = threadIdx.x; sl < len; sl += SORT_THREADS) {
    const uint2 rec = rec_ld<true>(b, sl);
    const uint32_t bk = sort_bucket(rec.x, tmin, scale);
    uint32_t end, beg = 0u;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(end) : "r"(hs + 2u * bk));
    if (bk > 0u) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(beg) : "r"(hs + 2u * (bk - 1u)));
    uint32_t r = beg;
    for (uint32_t j = beg; j < end; ++j) {
      const uint2 o = rec_ld<true>(b, (int)j);
      r += (o.x < rec.x || (o.x == rec.x && o.y < rec.y)) ? 1u : 0u;
    }
    if (sorted_row) sorted_row[r] = float_from_sort_key(rec.x);
    if (perm_row) perm_row[r] = (int32_t)rec.y;
  }
  return true;
}

// KEYS: 0 = keys read from `keys` (B*P rows of len), 1 = circle keys of cloud x through frames fr (P,3,2), 2 = line keys
// through directions fr (P,3).  seg = b * P + p.
template <int KEYS, bool BUCKETS>
__global__ void __launch_bounds__(SORT_THREADS) segmented_sort_trim_kernel(const float* __restrict__ keys, const float* __restrict__ x,
                                                                           const float* __restrict__ fr, int P, int len,
                                                                           float* __restrict__ sorted, int32_t* __restrict__ perm32,
                                                                           int pp) {
  extern __shared__ uint2 sbuf[];
  __shared__ uint32_t hist32[SORT_WARPS * (1 << SORT_TRIM_MAXW) / 2];
  __shared__ uint32_t wt[SORT_WARPS];
  __shared__ uint32_t s_or[SORT_WARPS], s_and[SORT_WARPS], s_min[SORT_WARPS], s_max[SORT_WARPS];
  __shared__ float s_fr[6];
  const size_t seg = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RecBuf<true> a = make_recbuf<true>(sbuf), b = make_recbuf<true>(sbuf + len);
  constexpr bool COUNTED = BUCKETS && KEYS == 1;  // circle coordinates: the bucket histogram is filled as the keys are made
  if (KEYS != 0) {
    const size_t p = pp ? seg : seg % P;  // pp: one frame set per pair, fr is (B,P,...)
    const int nf = KEYS == 1 ? 6 : 3;
    if (threadIdx.x < nf) s_fr[threadIdx.x] = __ldg(fr + p * nf + threadIdx.x);
    if (COUNTED)
      for (int i = threadIdx.x; i < SORT_NB / 2; i += SORT_THREADS) hist32[i] = 0;
    __syncthreads();
  }
  uint32_t vor = 0u, vand = 0xffffffffu, kmin = 0xffffffffu, kmax = 0u;
  for (int i = threadIdx.x; i < len; i += SORT_THREADS) {
    float kf;
    if (KEYS == 0) {
      kf = __ldg(keys + seg * len + i);
    } else {
      const float* xp = x + ((seg / P) * (size_t)len + i) * 3;
      const float x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2);
      kf = KEYS == 1 ? circle_key(s_fr, x0, x1, x2) : line_key(s_fr, x0, x1, x2);
    }
    const uint32_t k = float_sort_key(kf);
    vor |= k;
    vand &= k;
    kmin = min(kmin, k);
    kmax = max(kmax, k);
    rec_st<true>(a, i, make_uint2(k, (uint32_t)i));
    if (COUNTED && len >= SORT_BUCKET_MIN_LEN) {
      const uint32_t bk = sort_bucket(k, 0.f, (float)SORT_NB);
      atomicAdd(hist32 + (bk >> 1), 1u << ((bk & 1u) * 16u));
    }
  }
  vor = __reduce_or_sync(0xffffffffu, vor);
  vand = __reduce_and_sync(0xffffffffu, vand);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) {
    s_or[warp] = vor;
    s_and[warp] = vand;
    s_min[warp] = kmin;
    s_max[warp] = kmax;
  }
  __syncthreads();
  vor = 0u;
  vand = 0xffffffffu;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; ++w) {
    vor |= s_or[w];
    vand &= s_and[w];
    kmin = min(kmin, s_min[w]);
    kmax = max(kmax, s_max[w]);
  }
  const uint32_t varying = vor ^ vand;  // (CTA-uniform) bits in which the row's keys differ
  if (BUCKETS && varying != 0u &&
      bucket_sort_row<COUNTED>(a, b, len, kmin, kmax, hist32, wt, sorted ? sorted + seg * len : nullptr,
                               perm32 ? perm32 + seg * len : nullptr))
    return;
  if (varying != 0u) {
    const int lo = __ffs(varying) - 1, nbits = 32 - __clz(varying) - lo;
    const int passes = (nbits + SORT_TRIM_MAXW - 1) / SORT_TRIM_MAXW;
    const int w = (nbits + passes - 1) / passes;  // 1 .. 9 bits per pass
    const uint32_t dmask = (1u << w) - 1u;
    const int chunk = (((len + SORT_WARPS - 1) / SORT_WARPS) + 31) & ~31;
    const int beg = min(len, warp * chunk), end = min(len, beg + chunk);
    for (int pass = 0; pass < passes; ++pass) {
      const int sh = lo + pass * w;
      if (w <= 7)
        radix_pass_trim<7>(a, b, beg, end, sh, dmask, hist32, wt);
      else if (w == 8)
        radix_pass_trim<8>(a, b, beg, end, sh, dmask, hist32, wt);
      else
        radix_pass_trim<9>(a, b, beg, end, sh, dmask, hist32, wt);
      const RecBuf<true> t = a;
      a = b;
      b = t;
    }
  }
  for (int i = threadIdx.x; i < len; i += SORT_THREADS) {
    const uint2 r = rec_ld<true>(a, i);
    if (sorted) sorted[seg * len + i] = float_from_sort_key(r.x);
    if (perm32) perm32[seg * len + i] = (int32_t)r.y;
  }
}

// ---- compact shared-memory layout ------------------------------------------------------------------------------------
// The same stable LSD radix sort with 4-byte keys and 2-byte indices in SEPARATE ping-pong arrays (12 B per item instead of
// the 16 B of two uint2 buffers) and the per-warp digit counters as 16-bit halves (offsets < 65536): rows of up to
// SORT_COMPACT_MAX = 16384 keys stay in shared memory (the uint2 layout ends at 8192 and falls back to global scratch), a
// row of 8192 keys leaves room for two CTAs per SM instead of one, a row of 4096 keys for four instead of three.
constexpr int SORT_COMPACT_MAX = 16384;

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts_u16_if(uint32_t a, uint32_t v, bool pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.u16 [%0], %1; }" ::"r"(a), "h"((uint16_t)v), "r"((int)pred) : "memory");
}

// nb > 0: try the one-pass bucket sort first (bucket_sort_row's method on this layout: nb buckets of 16-bit counters behind the
// arrays -- the launcher sizes the counter area for max(radix counters, nb) -- keys scattered from keyA into (keyB, idxB) in
// arrival order, then ranked inside their bucket by (key, index)); rows it does not take (a non-finite key, a bucket above
// SORT_BUCKET_CAP, fewer than SORT_BUCKET_MIN_LEN keys) run the four radix passes as before, nothing having been moved.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) segmented_sort_compact_kernel(const float* __restrict__ keys, int len,
                                                                            float* __restrict__ sorted, int64_t* __restrict__ perm,
                                                                            int32_t* __restrict__ perm32, int nb) {
  constexpr int THREADS = WARPS * 32;
  static_assert(THREADS >= 256, "one thread per digit in the offset scan");
  extern __shared__ __align__(16) unsigned char csm[];
  __shared__ uint32_t wt[WARPS];
  __shared__ uint32_t s_kmin[WARPS], s_kmax[WARPS];
  // keyA[len] keyB[len] (u32) | idxA[len] idxB[len] (u16) | hist[WARPS][256] (u16)
  uint32_t ka = (uint32_t)__cvta_generic_to_shared(csm), kb = ka + 4u * len;
  uint32_t ia = kb + 4u * len, ib = ia + 2u * len;
  const uint32_t hs = ib + 2u * len;  // 4-byte aligned: 12 * len
  uint32_t* hist32 = reinterpret_cast<uint32_t*>(csm + 12 * (size_t)len);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t seg = blockIdx.x;
  const float* k = keys + seg * len;
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  for (int i = tid; i < len; i += THREADS) {
    const uint32_t key = float_sort_key(__ldg(k + i));
    kmin = min(kmin, key);
    kmax = max(kmax, key);
    sts_u32(ka + 4u * i, key);
    sts_u16(ia + 2u * i, (uint32_t)i);
  }
  if (nb > 0 && len >= SORT_BUCKET_MIN_LEN) {
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) {
      s_kmin[warp] = kmin;
      s_kmax[warp] = kmax;
    }
    __syncthreads();  // (also the key stores above)
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      kmin = min(kmin, s_kmin[w]);
      kmax = max(kmax, s_kmax[w]);
    }
    // finite keys only (+-inf / NaN rows take the radix passes), and a range the map can resolve
    bool ok = kmax < 0xFF800000u && kmin > 0x007FFFFFu && kmax > kmin;
    const float tmin = float_from_sort_key(kmin);
    const float scale = ok ? __fdiv_rn((float)nb, __fsub_rn(float_from_sort_key(kmax), tmin)) : 0.f;
    ok = ok && scale < 3.0e38f;
    if (ok) {  // CTA-uniform
      const int words = nb >> 1, wpt = words / THREADS;  // nb is a multiple of 2 * THREADS (launcher)
      const uint32_t nbm1 = (uint32_t)nb - 1u;
      auto bucket = [&](uint32_t key) { return min(__float2uint_rz(__fmul_rn(__fsub_rn(float_from_sort_key(key), tmin), scale)), nbm1); };
      for (int i = tid; i < words; i += THREADS) hist32[i] = 0u;
      __syncthreads();
      for (int i = tid; i < len; i += THREADS) {
        const uint32_t bk = bucket(lds_u32(ka + 4u * i));
        atomicAdd(hist32 + (bk >> 1), 1u << ((bk & 1u) * 16u));
      }
      __syncthreads();
      uint32_t tot = 0, mx = 0;
      for (int w = 0; w < wpt; ++w) {
        const uint32_t c = hist32[tid * wpt + w];
        tot += (c & 0xffffu) + (c >> 16);
        mx = max(mx, max(c & 0xffffu, c >> 16));
      }
      if (!__syncthreads_or(mx > (uint32_t)SORT_BUCKET_CAP)) {
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) wt[warp] = inc;
        __syncthreads();
        uint32_t run = inc - tot;
        for (int w = 0; w < warp; ++w) run += wt[w];
        for (int w = 0; w < wpt; ++w) {  // counters -> first slot of each bucket (the scatter advances them to the bucket's end)
          const uint32_t c = hist32[tid * wpt + w];
          const uint32_t lo = run, hi = run + (c & 0xffffu);
          hist32[tid * wpt + w] = lo | (hi << 16);
          run = hi + (c >> 16);
        }
        __syncthreads();
        for (int i = tid; i < len; i += THREADS) {
          const uint32_t key = lds_u32(ka + 4u * i);
          const uint32_t bk = bucket(key);
          const uint32_t old = atomicAdd(hist32 + (bk >> 1), 1u << ((bk & 1u) * 16u));
          const uint32_t slot = (old >> ((bk & 1u) * 16u)) & 0xffffu;
          sts_u32(kb + 4u * slot, key);
          sts_u16(ib + 2u * slot, (uint32_t)i);
        }
        __syncthreads();
        for (int sl = tid; sl < len; sl += THREADS) {
          const uint32_t key = lds_u32(kb + 4u * sl), idx = lds_u16(ib + 2u * sl);
          const uint32_t bk = bucket(key);
          const uint32_t end = lds_u16(hs + 2u * bk), beg = bk ? lds_u16(hs + 2u * (bk - 1u)) : 0u;
          uint32_t r = beg;
          for (uint32_t j = beg; j < end; ++j) {
            const uint32_t ok2 = lds_u32(kb + 4u * j), oi = lds_u16(ib + 2u * j);
            r += (ok2 < key || (ok2 == key && oi < idx)) ? 1u : 0u;
          }
          if (sorted) sorted[seg * len + r] = __ldg(k + idx);  // the original value (keeps -0.0)
          if (perm) perm[seg * len + r] = (int64_t)idx;
          if (perm32) perm32[seg * len + r] = (int32_t)idx;
        }
        return;
      }
    }
  }
  const int chunk = (((len + WARPS - 1) / WARPS) + 31) & ~31;
  const int beg = min(len, warp * chunk), end = min(len, beg + chunk);
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t wh = hs + 512u * warp;  // this warp's 256 16-bit counters
  for (int pass = 0; pass < 4; ++pass) {
    const int sh = pass * 8;
    for (int i = tid; i < WARPS * 128; i += THREADS) hist32[i] = 0u;
    __syncthreads();  // (first pass: also the record stores above)
    // counting needs no order: one shared-memory atomic per key on the 32-bit word that holds the digit's 16-bit half
    for (int i = beg + lane; i < end; i += 32) {
      const uint32_t d = (lds_u32(ka + 4u * i) >> sh) & 255u;
      atomicAdd(hist32 + warp * 128 + (d >> 1), 1u << ((d & 1u) << 4));
    }
    __syncthreads();
    {
      // digit-major exclusive offsets: thread d < 256 owns digit d
      const bool own = tid < 256;
      uint32_t tot = 0;
      if (own) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) tot += lds_u16(hs + 512u * w + 2u * tid);
      }
      uint32_t inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) wt[warp] = inc;
      __syncthreads();
      uint32_t run = inc - tot;
      for (int w = 0; w < warp; ++w) run += wt[w];
      if (own) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
          const uint32_t a = hs + 512u * w + 2u * tid;
          const uint32_t t = lds_u16(a);
          sts_u16(a, run);
          run += t;
        }
      }
    }
    __syncthreads();
    int i0 = beg;
#pragma unroll 1
    for (; i0 + 32 <= end; i0 += 32) {  // full batches
      const uint32_t key = lds_u32(ka + 4u * (i0 + lane));
      const uint32_t idx = lds_u16(ia + 2u * (i0 + lane));
      const uint32_t d = (key >> sh) & 255u;
      const uint32_t peers = match_digit<8>(d);
      const uint32_t cur = lds_u16(wh + 2u * d);
      const uint32_t dst = cur + __popc(peers & lt);
      sts_u32(kb + 4u * dst, key);
      sts_u16(ib + 2u * dst, idx);
      sts_u16_if(wh + 2u * d, cur + __popc(peers), (peers & lt) == 0);
      __syncwarp();
    }
    if (i0 < end) {  // ragged tail of the warp's range
      const int i = i0 + lane;
      const bool valid = i < end;
      const uint32_t key = valid ? lds_u32(ka + 4u * i) : 0u;
      const uint32_t idx = valid ? lds_u16(ia + 2u * i) : 0u;
      const uint32_t d = valid ? ((key >> sh) & 255u) : 256u;
      const uint32_t peers = match_digit<9>(d);
      const uint32_t cur = lds_u16(wh + 2u * (d & 255u));
      const uint32_t dst = cur + __popc(peers & lt);
      if (valid) {
        sts_u32(kb + 4u * dst, key);
        sts_u16(ib + 2u * dst, idx);
      }
      sts_u16_if(wh + 2u * (d & 255u), cur + __popc(peers), valid && (peers & lt) == 0);
      __syncwarp();
    }
    __syncthreads();
    uint32_t t = ka;
    ka = kb;
    kb = t;
    t = ia;
    ia = ib;
    ib = t;
  }
  for (int i = tid; i < len; i += THREADS) {
    const uint32_t j = lds_u16(ia + 2u * i);
    if (sorted) sorted[seg * len + i] = __ldg(k + j);  // the original value (keeps -0.0 and NaN payloads)
    if (perm) perm[seg * len + i] = (int64_t)j;
    if (perm32) perm32[seg * len + i] = (int32_t)j;
  }
}

// gkeys[seg][perm[seg][k]] = gsorted[seg][k]
__global__ void unsort_kernel(const float* __restrict__ gs, const int64_t* __restrict__ perm, size_t total, int len,
                              float* __restrict__ gk) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  size_t seg = i / len;
  gk[seg * len + (size_t)perm[i]] = gs[i];
}

// ------------------------------------------------------------------------------------- circular W1 (level median) --
__device__ __forceinline__ float block_sum_f32(float v, float* wtot) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) wtot[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < SORT_WARPS; ++w) t += wtot[w];
  return t;
}

constexpr int CW1_THREADS = 512;     // 16 warps: with C <= 16 entries per thread (n + m <= 8192) 64 registers -> two CTAs, 32 warps per SM
constexpr int CW1_WARPS = CW1_THREADS / 32;
constexpr int CW1_PER_THREAD = 20;   // merged entries per thread with register-resident F keys (n + m <= 512 * 20 = 10240)
constexpr int CW1_PER_THREAD_MAX = 64;  // ... with recomputed F keys (n + m <= 32768: cfg4's two clouds of 16384 points)

// Fixed-order block sum with ONE barrier: the per-warp partials ping-pong between two shared-memory rows.
__device__ __forceinline__ float block_sum_pp(float v, float (*wf)[CW1_WARPS], int& phase) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) wf[phase][threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < CW1_WARPS; ++w) t += wf[phase][w];
  phase ^= 1;
  return t;
}

// shared-memory index with one pad word per 32: the merge walks su / sv at a lane stride of ~c/2 words, which without the
// pad lands every lane of a warp in one or two banks
__host__ __device__ __forceinline__ int cw1_pad(int i) { return i + (i >> 5); }

// One CTA per slice.  us (n), vs (m) sorted ascending.  Follows emd1D_circle :230-247:
//   merged = stable merge of (us, vs)  [== sort(cat(us, vs))],  w = +1/n for u entries, -1/m for v entries
//   F = cumsum(w);  delta_k = merged_{k+1} - merged_k, last = 1 - merged_last   (the arc [0, first) is omitted)
//   level median: sort F, cw = cumsum(delta[perm]) - 0.5, first k with cw >= 0 -> med = F_sorted[k]
//   W = sum_k delta_k |F_k - med|;   dW/dmerged_k = |F_{k-1} - med| - |F_k - med|  (F_{-1} term = 0)
// Nothing is sorted here.  The merge is a merge-path: thread t owns the merged positions [t*c, (t+1)*c), finds its split
// of (us, vs) with one binary search along the diagonal and merges sequentially -- INTO REGISTERS: a thread keeps
// (key(F_k), delta_k) of its own c entries plus a bit mask of which came from u, so the merged values, F and the origin
// array never exist in shared memory (the first version wrote them at a lane stride of c words, a 32-way bank conflict,
// and its 16 (n+m) bytes of shared memory allowed one CTA per SM; now 4 (n+m) bytes and two CTAs per SM).  F's scan uses
// the same chunking (thread-sequential, block scan of the chunk totals).  The level median is a SELECTION: med is the
// smallest F value f with  sum_{k: F_k <= f} delta_k >= 0.5  (the cumulative sum over the sorted order first reaches 0.5
// inside the run of entries equal to f), found by bisection on the 32-bit order-preserving key of F -- 32 rounds of a
// register-resident partial sum + a fixed-order block reduction -- instead of a radix sort of n + m records.
// C > CW1_PER_THREAD (slices of up to 32768 merged entries): 2 C words per thread no longer fit in the register file next
// to the 4 (n+m) bytes of rows in shared memory, so only delta_k stays in registers and key(F_k) is RECOMPUTED wherever it is
// needed -- F is a running sum of +-w from the chunk's base in merged order, i.e. one select + add + key map per entry,
// bit-identical every time; this replaces the reference's four-sort composition on global scratch that such slices used to take.
template <int C>
__global__ void __launch_bounds__(CW1_THREADS, (C <= 16 ? 2 : 1))
    circular_w1_kernel(const float* __restrict__ us, const float* __restrict__ vs, const int32_t* __restrict__ pu,
                       const int32_t* __restrict__ pv, int n, int m, float* __restrict__ w_out, float* __restrict__ gus,
                       float* __restrict__ gvs) {
  extern __shared__ float cw1_smem[];
  __shared__ float wf[2][CW1_WARPS];
  __shared__ float s_first[CW1_THREADS];  // first merged value of each thread's chunk
  __shared__ float s_lastF[CW1_THREADS];  // F at the last entry of each thread's chunk
  __shared__ uint32_t s_sel[3][4];  // (sum, largest key <= mid, smallest key > mid) of a selection round, triple-buffered
  const int nm = n + m;
  float* su = cw1_smem;                 // n (padded)
  float* sv = su + cw1_pad(n - 1) + 1;  // m (padded)
  const size_t s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float wu = 1.f / n, wv = 1.f / m;
  for (int i = tid; i < n; i += CW1_THREADS) su[cw1_pad(i)] = __ldg(us + s * n + i);
  for (int j = tid; j < m; j += CW1_THREADS) sv[cw1_pad(j)] = __ldg(vs + s * m + j);
  if (tid < 12) s_sel[tid >> 2][tid & 3] = ((tid & 3) == 2) ? 0xFFFFFFFFu : 0u;
  __syncthreads();
  // ---- merge path: u goes first on ties (u_i lands at i + #{v < u_i}; v_j at j + #{u <= v_j})
  const int c = (nm + CW1_THREADS - 1) / CW1_THREADS;  // <= C
  const int d0 = min(nm, tid * c), d1 = min(nm, d0 + c);
  const int cnt = d1 - d0;
  int i0, j0;
  {
    int lo = max(0, d0 - m), hi = min(d0, n);  // i = #u entries among the first d0 merged ones
    while (lo < hi) {
      const int i = (lo + hi) >> 1;
      if (su[cw1_pad(i)] <= sv[cw1_pad(d0 - i - 1)]) lo = i + 1; else hi = i;
    }
    i0 = lo;
    j0 = d0 - lo;
  }
  constexpr bool RECOMPUTE = C > CW1_PER_THREAD;
  typedef typename std::conditional<(C > 32), unsigned long long, uint32_t>::type mask_t;
  uint32_t key[RECOMPUTE ? 1 : C];  // first the merged value bits, later key(F_k)
  float dl[C];
  mask_t from_u = 0;
  float wsum = 0.f;  // chunk total of +-w, summed in merged order
  {
    int i = i0, j = j0;
    float cu = (i < n) ? su[cw1_pad(i)] : 0.f, cv = (j < m) ? sv[cw1_pad(j)] : 0.f;
    float prev = 0.f;
#pragma unroll
    for (int q = 0; q < C; ++q) {
      dl[q] = 0.f;
      if (q < cnt) {
        const bool take_u = (j >= m) || (i < n && cu <= cv);
        const float v = take_u ? cu : cv;
        if (take_u) {
          from_u |= (mask_t)1 << q;
          wsum += wu;
          ++i;
          if (i < n) cu = su[cw1_pad(i)];
        } else {
          wsum -= wv;
          ++j;
          if (j < m) cv = sv[cw1_pad(j)];
        }
        if (q == 0) s_first[tid] = v;
        if (q > 0) dl[q - 1] = v - prev;
        prev = v;
      }
    }
    // the last delta of the chunk needs the next chunk's first value: parked in key[] until the exchange below
    key[0] = __float_as_uint(prev);
  }
  // ---- exclusive block scan of the chunk totals (same association as a thread-sequential scan of F)
  float base;
  {
    float inc = wsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wf[0][warp] = inc;
    __syncthreads();  // also publishes s_first
    float wb = 0.f;
    for (int w = 0; w < warp; ++w) wb += wf[0][w];
    base = wb + inc - wsum;
  }
  if (cnt > 0) {
    const float last = __uint_as_float(key[0]);
    const float nxt = (d1 < nm) ? s_first[tid + 1] : 1.f;
#pragma unroll
    for (int q = 0; q < C; ++q)
      if (q == cnt - 1) dl[q] = nxt - last;
  }
  float tot = 0.f;
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  {
    float run = base;
#pragma unroll
    for (int q = 0; q < C; ++q) {
      if (!RECOMPUTE) key[q] = 0xFFFFFFFFu;
      if (q < cnt) {
        run += ((from_u >> q) & 1u) ? wu : -wv;
        const uint32_t kq = float_sort_key(run);
        if (!RECOMPUTE) key[q] = kq;
        kmin = min(kmin, kq);
        kmax = max(kmax, kq);
      }
      tot += dl[q];
    }
    s_lastF[tid] = run;
  }
  // ---- level median = the smallest key K with  S(K) := sum_{key <= K} delta >= 1/2.  A search on the VALUES with both ends
  // snapped to keys that exist: [lo, hi] always holds K, lo and hi are keys of the row; a round evaluates S at the key of
  // the midpoint of the two VALUES and, in the same pass, the largest key <= mid (kb) and the smallest key > mid (ka);
  // S(mid) >= 1/2 -> hi = kb (S(kb) = S(mid)), else lo = ka (every key below ka sums to less than 1/2).  The interval of
  // values at least halves and loses a key per round, so the search ends after ~log2(#distinct F values) rounds -- F takes
  // few distinct values (levels 1/n apart, plus the rounding spread inside a level): cfg3 rows end in ~10 rounds where the
  // plain bisection of the 32-bit key needed 32.  After 24 rounds the midpoint is taken between the KEYS (at most 32 more).
  // Partial sums: a thread adds its own deltas in float32 (fixed order), the CTA adds the threads' partials in 2^-31 fixed
  // point -- integers, so one redux.sync per warp and one shared-memory atomic per warp replace the shuffle tree and the
  // sixteen-way read-back, and the total does not depend on the order of arrival.
  uint32_t kmed;
  {
    constexpr float FIX = 2147483648.f;  // 2^31
    constexpr uint32_t HALF = 1u << 30;
    int set = 0;
    auto exchange = [&](float part, uint32_t kb, uint32_t ka, uint32_t& S, uint32_t& KB, uint32_t& KA) {
      uint32_t ps = __reduce_add_sync(0xffffffffu, __float2uint_rn(part * FIX));
      kb = __reduce_max_sync(0xffffffffu, kb);
      ka = __reduce_min_sync(0xffffffffu, ka);
      if (lane == 0) {
        atomicAdd(&s_sel[set][0], ps);
        atomicMax(&s_sel[set][1], kb);
        atomicMin(&s_sel[set][2], ka);
      }
      __syncthreads();
      S = s_sel[set][0];
      KB = s_sel[set][1];
      KA = s_sel[set][2];
      const int clr = set == 0 ? 2 : set - 1;  // read last in the previous round, used again two rounds from now
      if (tid < 3) s_sel[clr][tid] = (tid == 2) ? 0xFFFFFFFFu : 0u;
      set = set == 2 ? 0 : set + 1;
    };
    uint32_t S, lo, hi;
    exchange(tot, kmax, kmin, S, hi, lo);  // the whole row: total, largest and smallest key
    if (S >= HALF) {
      for (int round = 0; lo < hi; ++round) {
        uint32_t mid = (round < 24) ? float_sort_key(0.5f * float_from_sort_key(lo) + 0.5f * float_from_sort_key(hi))
                                    : lo + ((hi - lo) >> 1);
        mid = min(max(mid, lo), hi - 1u);
        float part = 0.f;
        uint32_t kb = 0u, ka = 0xFFFFFFFFu;
        if (RECOMPUTE) {
          // (rows of more than 20 entries per thread recompute their keys and have no registers to spare for the two
          //  snapped ends: plain bisection of the key interval, with the integer exchange)
          mid = lo + ((hi - lo) >> 1);
          float run = base;  // dl[q] == 0 beyond cnt, so the surplus entries add nothing whatever their recomputed key
#pragma unroll
          for (int q = 0; q < C; ++q) {
            run += ((from_u >> q) & 1u) ? wu : -wv;
            part += (float_sort_key(run) <= mid) ? dl[q] : 0.f;
          }
          kb = mid;
          ka = mid + 1u;
        } else {
#pragma unroll
          for (int q = 0; q < C; ++q) {
            const bool le = key[q] <= mid;
            part += le ? dl[q] : 0.f;
            kb = max(kb, le ? key[q] : 0u);
            ka = min(ka, le ? 0xFFFFFFFFu : key[q]);
          }
        }
        uint32_t KB, KA;
        exchange(part, kb, ka, S, KB, KA);
        if (S >= HALF) hi = KB; else lo = KA;
      }
      kmed = lo;
    } else {
      kmed = lo;  // the cumulative sum never reaches 1/2 (all cw < 0 -> all inf -> argmin = 0): the smallest F
    }
  }
  const float med = float_from_sort_key(kmed);
  int phase = 1;
  // ---- W and dW/d(merged value), the latter parked in su / sv (dead since the merge) for a coalesced write-out
  const bool want_g = gus || gvs;
  float acc = 0.f;
  {
    float aprev = (tid > 0 && cnt > 0) ? fabsf(s_lastF[tid - 1] - med) : 0.f;  // |F_{k-1} - med|, 0 before the first entry
    int i = i0, j = j0;
    float run = base;
#pragma unroll
    for (int q = 0; q < C; ++q) {
      if (q < cnt) {
        run += ((from_u >> q) & 1u) ? wu : -wv;
        const float a = fabsf(float_from_sort_key(RECOMPUTE ? float_sort_key(run) : key[RECOMPUTE ? 0 : q]) - med);
        acc += dl[q] * a;
        if (want_g) {
          if ((from_u >> q) & 1u) su[cw1_pad(i++)] = aprev - a; else sv[cw1_pad(j++)] = aprev - a;
        }
        aprev = a;
      }
    }
  }
  acc = block_sum_pp(acc, wf, phase);  // its barrier also orders the gradient stores above
  if (tid == 0) w_out[s] = acc;
  // pu / pv given: the gradient goes to the UNSORTED key positions (the scatter torch.sort's backward would do).  With two
  // CTAs per SM (C <= 16) the row is first permuted in shared memory (stage: max(n, m) floats behind the rows) and written
  // out coalesced: circular_w1_kernel<16> at cfg3 490 -> 431 us.  The one-CTA-per-SM variants lose (C = 64 at cfg4: 370 ->
  // 439 us, the extra barriers are not hidden by a second CTA) and keep the direct scatter; so does euclid_sw_kernel
  // (157 -> 236 us when staged).
  constexpr bool STAGE = C <= 16;
  float* stage = sv + cw1_pad(m - 1) + 1;
  if (gus) {
    if (pu && STAGE) {
      for (int i = tid; i < n; i += CW1_THREADS) stage[__ldg(pu + s * n + i)] = su[cw1_pad(i)];
      __syncthreads();
      for (int i = tid; i < n; i += CW1_THREADS) gus[s * n + i] = stage[i];
      __syncthreads();
    } else {
      for (int i = tid; i < n; i += CW1_THREADS) gus[s * n + (pu ? __ldg(pu + s * n + i) : i)] = su[cw1_pad(i)];
    }
  }
  if (gvs) {
    if (pv && STAGE) {
      for (int j = tid; j < m; j += CW1_THREADS) stage[__ldg(pv + s * m + j)] = sv[cw1_pad(j)];
      __syncthreads();
      for (int j = tid; j < m; j += CW1_THREADS) gvs[s * m + j] = stage[j];
    } else {
      for (int j = tid; j < m; j += CW1_THREADS) gvs[s * m + (pv ? __ldg(pv + s * m + j) : j)] = sv[cw1_pad(j)];
    }
  }
}

template <int C>
static int launch_circular_w1(const float* us, const float* vs, const int32_t* pu, const int32_t* pv, int S, int n, int m, float* w,
                              float* gus, float* gvs, cudaStream_t stream) {
  const size_t smem = (size_t)(cw1_pad(n - 1) + 1 + cw1_pad(m - 1) + 1 + ((C <= 16 && (pu || pv)) ? (n > m ? n : m) : 0)) * sizeof(float);
  if (smem > 32 * 1024)  // static + dynamic beyond 48 KB needs the opt-in (static is < 16 KB here)
    SHWD_CUDA_CHECK(cudaFuncSetAttribute(circular_w1_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  circular_w1_kernel<C><<<S, CW1_THREADS, smem, stream>>>(us, vs, pu, pv, n, m, w, gus, gvs);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// acc[s] = sum_k |xs_k - ys_k|^p on sorted projections; gradients w.r.t. the sorted values.
__global__ void __launch_bounds__(SORT_THREADS) euclid_sw_kernel(const float* __restrict__ xs, const float* __restrict__ ys,
                                                                 const int32_t* __restrict__ px, const int32_t* __restrict__ py, int n,
                                                                 float p, float* __restrict__ acc_out, float* __restrict__ gxs,
                                                                 float* __restrict__ gys) {
  __shared__ float wf[SORT_WARPS];
  const size_t s = blockIdx.x;
  float acc = 0.f;
  for (int k = threadIdx.x; k < n; k += SORT_THREADS) {
    const float d = xs[s * n + k] - ys[s * n + k];
    const float ad = fabsf(d);
    float t, g;
    if (p == 2.f) {
      t = d * d;
      g = 2.f * d;
    } else if (p == 1.f) {
      t = ad;
      g = (d > 0.f) - (d < 0.f);
    } else {
      t = powf(ad, p);
      g = copysignf(p * powf(ad, p - 1.f), d);
    }
    acc += t;
    if (gxs) gxs[s * n + (px ? __ldg(px + s * n + k) : k)] = g;
    if (gys) gys[s * n + (py ? __ldg(py + s * n + k) : k)] = -g;
  }
  acc = block_sum_f32(acc, wf);
  if (threadIdx.x == 0) acc_out[s] = acc;
}

}  // namespace shwd

using namespace shwd;

static int g_project_bwd_wide = 1;  // shwd_project_bwd_set_wide (A/B, tests)
extern "C" int shwd_project_bwd_set_wide(int on) {
  g_project_bwd_wide = on ? 1 : 0;
  return SHWD_OK;
}
static void launch_project_circle_bwd(const float* x, const float* U, int B, int N, int P, const float* gkeys, const float* gw,
                                      float* gx, cudaStream_t s, int pp = 0) {
  // four points per lane when the rows allow vector loads and the grid still fills the SMs
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gkeys)) & 15) == 0;
  if (g_project_bwd_wide && (N % 4) == 0 && al && (long long)B * ((N + 127) / 128) >= 148) {
    dim3 grid((N + 127) / 128, B);
    project_circle_bwd4_kernel<<<grid, PJ_THREADS, 0, s>>>(x, U, N, P, gkeys, gw, gx, pp);
  } else {
    dim3 grid((N + 31) / 32, B);
    project_circle_bwd_kernel<<<grid, PJ_THREADS, 0, s>>>(x, U, N, P, gkeys, gw, gx, pp);
  }
}

static int project_circle_any(const float* x, const float* U, int B, int N, int P, float* keys, void* stream, int pp) {
  if (!x || !U || !keys || B <= 0 || N <= 0 || P <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B > 65535 || (P + 31) / 32 > 65535) return SHWD_ERR_UNSUPPORTED;
  dim3 grid((N + PJ_THREADS - 1) / PJ_THREADS, (P + 31) / 32, B);
  project_circle_kernel<<<grid, PJ_THREADS, 32 * 6 * sizeof(float), static_cast<cudaStream_t>(stream)>>>(x, U, N, P, keys, pp);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}
extern "C" int shwd_project_circle(const float* x, const float* U, int B, int N, int P, float* keys, void* stream) {
  return project_circle_any(x, U, B, N, P, keys, stream, 0);
}
extern "C" int shwd_project_circle_pp(const float* x, const float* U, int B, int N, int P, float* keys, void* stream) {
  return project_circle_any(x, U, B, N, P, keys, stream, 1);
}

extern "C" int shwd_project_circle_bwd(const float* x, const float* U, int B, int N, int P, const float* gkeys, float* gx,
                                       void* stream) {
  if (!x || !U || !gkeys || !gx || B <= 0 || N <= 0 || P <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  launch_project_circle_bwd(x, U, B, N, P, gkeys, nullptr, gx, static_cast<cudaStream_t>(stream));
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_project_line(const float* x, const float* theta, int B, int N, int P, float* keys, void* stream) {
  if (!x || !theta || !keys || B <= 0 || N <= 0 || P <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B > 65535 || (P + 31) / 32 > 65535) return SHWD_ERR_UNSUPPORTED;
  dim3 grid((N + PJ_THREADS - 1) / PJ_THREADS, (P + 31) / 32, B);
  project_line_kernel<<<grid, PJ_THREADS, 32 * 3 * sizeof(float), static_cast<cudaStream_t>(stream)>>>(x, theta, N, P, keys);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_project_line_bwd(const float* theta, int B, int N, int P, const float* gkeys, float* gx, void* stream) {
  if (!theta || !gkeys || !gx || B <= 0 || N <= 0 || P <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  dim3 grid((N + 31) / 32, B);
  project_line_bwd_kernel<<<grid, PJ_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(theta, N, P, gkeys, nullptr, gx);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// The same with the slice-mean's chain rule folded in: gx[b] = (gw[b] / P) * sum_p ...  (gw: (B,) upstream gradient of the
// per-pair mean over the P slices, on the device) -- what ops.SlicedLossFn.backward launches.
extern "C" int shwd_project_circle_bwd_scaled(const float* x, const float* U, int B, int N, int P, const float* gkeys,
                                              const float* gw, float* gx, void* stream) {
  if (!x || !U || !gkeys || !gw || !gx || B <= 0 || N <= 0 || P <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  launch_project_circle_bwd(x, U, B, N, P, gkeys, gw, gx, static_cast<cudaStream_t>(stream));
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}
extern "C" int shwd_project_circle_bwd_scaled_pp(const float* x, const float* U, int B, int N, int P, const float* gkeys,
                                                 const float* gw, float* gx, void* stream) {
  if (!x || !U || !gkeys || !gw || !gx || B <= 0 || N <= 0 || P <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  launch_project_circle_bwd(x, U, B, N, P, gkeys, gw, gx, static_cast<cudaStream_t>(stream), 1);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}
extern "C" int shwd_project_line_bwd_scaled(const float* theta, int B, int N, int P, const float* gkeys, const float* gw,
                                            float* gx, void* stream) {
  if (!theta || !gkeys || !gw || !gx || B <= 0 || N <= 0 || P <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  dim3 grid((N + 31) / 32, B);
  project_line_bwd_kernel<<<grid, PJ_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(theta, N, P, gkeys, gw, gx);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// Which layout sorts a row of `len` keys.  Measured on B200 (tools/time_sort.py, Gkeys/s, wide / compact-8 warps / compact-16):
//   len 1024: 52.9 / 45.5 / 34.7    2048: 66.5 / 56.1 / 47.5    4096: 62.7 / 56.7 / 53.3    8192: 32.8 / 42.7 / 47.5
//   10240: 24.6 / 29.0 / 36.8       16384: 18.0 / 27.0 / 34.0 (1024 rows: 11.2 / 30.3 / 38.7)
// -> the uint2 records win while three of their CTAs fit on an SM (one 64-bit access per record and pass instead of two
// narrower ones), the compact layout with 16 warps wins from there on.  SHWD_SORT_LAYOUT=wide|compact and
// SHWD_SORT_WARPS=8|16 override the choice for A/B timing (read once).
constexpr int SORT_WIDE_BEST_MAX = 4224;
static int sort_env(const char* name, const char* a, int va, const char* b, int vb, int dflt) {
  const char* v = getenv(name);
  if (!v) return dflt;
  if (!strcmp(v, a)) return va;
  if (!strcmp(v, b)) return vb;
  return dflt;
}
static bool sort_use_compact(int len) {
  static const int forced = sort_env("SHWD_SORT_LAYOUT", "wide", 1, "compact", 2, 0);
  if (len > SORT_COMPACT_MAX) return false;
  if (forced == 1) return false;
  if (forced == 2) return true;
  return len > SORT_WIDE_BEST_MAX;
}

extern "C" size_t shwd_segmented_sort_workspace_bytes(int segs, int len) {
  if (segs <= 0 || len <= SORT_SMEM_MAX || sort_use_compact(len)) return 0;
  return (size_t)segs * 2 * (size_t)len * sizeof(uint2);
}

static int g_sort_method = 0;  // shwd_sort_set_method
template <int WARPS>
static int launch_sort_compact(const float* keys, int segs, int len, float* sorted, int64_t* perm, int32_t* perm32, cudaStream_t s) {
  // bucket counters: ~2 keys per bucket (8192 buckets beyond 8192 keys, where one CTA per SM fits anyway; 4096 below, which
  // keeps two CTAs of 8192 keys on an SM); 0 = radix passes only (shwd_sort_set_method(1))
  const int nb = g_sort_method == 1 ? 0 : (len > 8192 ? 8192 : 4096);
  const size_t hist = (size_t)WARPS * 512 > (size_t)nb * 2 ? (size_t)WARPS * 512 : (size_t)nb * 2;
  const size_t smem = 12 * (size_t)len + hist;
  if (smem > 32 * 1024)
    SHWD_CUDA_CHECK(cudaFuncSetAttribute(segmented_sort_compact_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  segmented_sort_compact_kernel<WARPS><<<segs, WARPS * 32, smem, s>>>(keys, len, sorted, perm, perm32, nb);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_sort_set_method(int method) {
  if (method < 0 || method > 1) return SHWD_ERR_INVALID_ARGUMENT;
  g_sort_method = method;
  return SHWD_OK;
}

// the trimmed-digit kernel (int32 permutation, values rebuilt from the keys): rows the uint2 layout is best for
template <int KEYS>
static int launch_sort_trim(const float* keys, const float* x, const float* fr, int P, int segs, int len, float* sorted, int32_t* perm32,
                            cudaStream_t s, int pp = 0) {
  const size_t smem = 2 * (size_t)len * sizeof(uint2);
  if (g_sort_method == 1) {
    if (smem > 32 * 1024)
      SHWD_CUDA_CHECK(
          cudaFuncSetAttribute(segmented_sort_trim_kernel<KEYS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    segmented_sort_trim_kernel<KEYS, false><<<segs, SORT_THREADS, smem, s>>>(keys, x, fr, P, len, sorted, perm32, pp);
  } else {
    if (smem > 32 * 1024)
      SHWD_CUDA_CHECK(
          cudaFuncSetAttribute(segmented_sort_trim_kernel<KEYS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    segmented_sort_trim_kernel<KEYS, true><<<segs, SORT_THREADS, smem, s>>>(keys, x, fr, P, len, sorted, perm32, pp);
  }
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}
static bool sort_use_trim(int len) {
  static const int forced = sort_env("SHWD_SORT_TRIM", "0", 1, "1", 2, 0);
  if (forced == 1) return false;
  return len <= SORT_WIDE_BEST_MAX && !sort_use_compact(len);
}

extern "C" int shwd_sort_projected_max_points(void) {
  static const int forced = sort_env("SHWD_SORT_TRIM", "0", 1, "1", 2, 0);
  return forced == 1 ? 0 : SORT_WIDE_BEST_MAX;
}

static int sort_projected_any(const float* x, const float* frames, int B, int N, int P, int mode, float* sorted, int32_t* perm,
                              void* stream, int pp) {
  if (!x || !frames || B < 0 || N <= 0 || P <= 0 || (mode != 1 && mode != 2) || (!sorted && !perm)) return SHWD_ERR_INVALID_ARGUMENT;
  if (N > shwd_sort_projected_max_points()) return SHWD_ERR_UNSUPPORTED;
  if (B == 0) return SHWD_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return mode == 1 ? launch_sort_trim<1>(nullptr, x, frames, P, B * P, N, sorted, perm, s, pp)
                   : launch_sort_trim<2>(nullptr, x, frames, P, B * P, N, sorted, perm, s, pp);
}
extern "C" int shwd_sort_projected(const float* x, const float* frames, int B, int N, int P, int mode, float* sorted, int32_t* perm,
                                   void* stream) {
  return sort_projected_any(x, frames, B, N, P, mode, sorted, perm, stream, 0);
}
extern "C" int shwd_sort_projected_pp(const float* x, const float* frames, int B, int N, int P, int mode, float* sorted, int32_t* perm,
                                      void* stream) {
  return sort_projected_any(x, frames, B, N, P, mode, sorted, perm, stream, 1);
}

static int launch_segmented_sort(const float* keys, int segs, int len, float* sorted, int64_t* perm, int32_t* perm32,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!keys || segs < 0 || len <= 0 || (!sorted && !perm && !perm32)) return SHWD_ERR_INVALID_ARGUMENT;
  if (segs == 0) return SHWD_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!perm && sort_use_trim(len)) return launch_sort_trim<0>(keys, nullptr, nullptr, 1, segs, len, sorted, perm32, s);
  if (sort_use_compact(len)) {
    static const int warps_env = sort_env("SHWD_SORT_WARPS", "8", 8, "16", 16, 0);
    const int warps = warps_env ? warps_env : 16;  // (32 warps for the one-CTA-per-SM rows: 34.1 vs 33.7 Gkeys/s at 16384, slower below)
    // rows beyond 8192 keys are alone on their SM: 32 warps (bucket pass 0.158 -> 0.141 ms at 512 x 16384, 0.106 -> 0.083 at
    // 512 x 10000; the radix passes do not care: 0.250 -> 0.242).  SHWD_SORT_BIG32=0 keeps 16 (A/B).
    static const int big32 = sort_env("SHWD_SORT_BIG32", "1", 1, "0", 2, 1);
    if (big32 == 1 && len > 8192) return launch_sort_compact<32>(keys, segs, len, sorted, perm, perm32, s);
    return warps == 16 ? launch_sort_compact<16>(keys, segs, len, sorted, perm, perm32, s)
                       : launch_sort_compact<8>(keys, segs, len, sorted, perm, perm32, s);
  }
  if (len <= SORT_SMEM_MAX) {
    const size_t smem = 2 * (size_t)len * sizeof(uint2);
    if (smem > 32 * 1024)  // static + dynamic beyond 48 KB needs the opt-in (static is < 16 KB here)
      SHWD_CUDA_CHECK(cudaFuncSetAttribute(segmented_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    segmented_sort_kernel<true><<<segs, SORT_THREADS, smem, s>>>(keys, len, sorted, perm, perm32, nullptr);
  } else {
    const size_t need = shwd_segmented_sort_workspace_bytes(segs, len);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 7)) return SHWD_ERR_WORKSPACE;
    segmented_sort_kernel<false><<<segs, SORT_THREADS, 0, s>>>(keys, len, sorted, perm, perm32, static_cast<uint2*>(workspace));
  }
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_segmented_sort(const float* keys, int segs, int len, float* sorted, int64_t* perm, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  return launch_segmented_sort(keys, segs, len, sorted, perm, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int shwd_segmented_sort_i32(const float* keys, int segs, int len, float* sorted, int32_t* perm, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  return launch_segmented_sort(keys, segs, len, sorted, nullptr, perm, workspace, workspace_bytes, stream);
}

extern "C" size_t shwd_circular_w1_workspace_bytes(int S, int n, int m) {
  (void)S;
  (void)n;
  (void)m;
  return 0;  // everything lives in shared memory / registers (n + m <= 10240)
}

static int circular_w1_dispatch(const float* us, const float* vs, const int32_t* pu, const int32_t* pv, int S, int n, int m, float* w,
                                float* gus, float* gvs, void* stream) {
  if (!us || !vs || !w || S < 0 || n <= 0 || m <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (S == 0) return SHWD_OK;
  const size_t nm = (size_t)n + m;
  if (nm > (size_t)CW1_THREADS * CW1_PER_THREAD_MAX) return SHWD_ERR_UNSUPPORTED;  // n + m <= 32768
  const int c = (int)((nm + CW1_THREADS - 1) / CW1_THREADS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c <= 4) return launch_circular_w1<4>(us, vs, pu, pv, S, n, m, w, gus, gvs, st);
  if (c <= 8) return launch_circular_w1<8>(us, vs, pu, pv, S, n, m, w, gus, gvs, st);
  if (c <= 12) return launch_circular_w1<12>(us, vs, pu, pv, S, n, m, w, gus, gvs, st);
  if (c <= 16) return launch_circular_w1<16>(us, vs, pu, pv, S, n, m, w, gus, gvs, st);
  if (c <= CW1_PER_THREAD) return launch_circular_w1<CW1_PER_THREAD>(us, vs, pu, pv, S, n, m, w, gus, gvs, st);
  if (c <= 32) return launch_circular_w1<32>(us, vs, pu, pv, S, n, m, w, gus, gvs, st);
  return launch_circular_w1<CW1_PER_THREAD_MAX>(us, vs, pu, pv, S, n, m, w, gus, gvs, st);
}

extern "C" int shwd_circular_w1(const float* us, const float* vs, int S, int n, int m, float* w, float* gus, float* gvs,
                                void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  return circular_w1_dispatch(us, vs, nullptr, nullptr, S, n, m, w, gus, gvs, stream);
}

extern "C" int shwd_circular_w1_scatter(const float* us, const float* vs, const int32_t* perm_u, const int32_t* perm_v, int S, int n,
                                        int m, float* w, float* gku, float* gkv, void* stream) {
  if (!perm_u || !perm_v || !gku || !gkv) return SHWD_ERR_INVALID_ARGUMENT;
  return circular_w1_dispatch(us, vs, perm_u, perm_v, S, n, m, w, gku, gkv, stream);
}

extern "C" int shwd_euclid_sw(const float* xs, const float* ys, int S, int n, float p, float* acc, float* gxs, float* gys,
                              void* stream) {
  if (!xs || !ys || !acc || S < 0 || n <= 0 || !(p > 0.f)) return SHWD_ERR_INVALID_ARGUMENT;
  if (S == 0) return SHWD_OK;
  euclid_sw_kernel<<<S, SORT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(xs, ys, nullptr, nullptr, n, p, acc, gxs, gys);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_euclid_sw_scatter(const float* xs, const float* ys, const int32_t* perm_x, const int32_t* perm_y, int S, int n,
                                      float p, float* acc, float* gkx, float* gky, void* stream) {
  if (!xs || !ys || !perm_x || !perm_y || !acc || !gkx || !gky || S < 0 || n <= 0 || !(p > 0.f)) return SHWD_ERR_INVALID_ARGUMENT;
  if (S == 0) return SHWD_OK;
  euclid_sw_kernel<<<S, SORT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(xs, ys, perm_x, perm_y, n, p, acc, gkx, gky);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_unsort(const float* gsorted, const int64_t* perm, int segs, int len, float* gkeys, void* stream) {
  if (!gsorted || !perm || !gkeys || segs < 0 || len <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  const size_t total = (size_t)segs * len;
  if (total == 0) return SHWD_OK;
  unsort_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(gsorted, perm, total, len, gkeys);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

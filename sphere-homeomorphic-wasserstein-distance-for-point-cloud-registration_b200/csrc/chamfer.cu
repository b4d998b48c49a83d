// Chamfer distance: tiled nearest-neighbour min-reduction, both directions in one launch; deterministic backward.
//
// Replaces pytorch3d.loss.chamfer_distance as the reference calls it (Point_Cloud_Resistration/train_CD.py:123,161;
// Comparison_Wasserstein_with_Chamfer_distance/main_rotation.py:203): squared L2, K = 1.
//   d_xy[b,i] = min_j |x_i - y_j|^2 (first minimum wins), idx_xy[b,i] = argmin;  d_yx / idx_yx likewise.
// The squared distance is formed as ((dx*dx + dy*dy) + dz*dz) with separately rounded products so it is bit-equal to
// torch's ((x-y)**2).sum(-1) -- the argmin indices are then bit-exact against the oracle.
//
// Forward: one thread owns one query point; the other cloud is streamed through shared memory in SoA form (three
// LDS.128 feed four candidates), 1024 points per tile.  The eight float32 operations of a squared distance run as
// packed f32x2 instructions on two candidates at once (same roundings, half the issue slots), and the running argmin
// is updated through a min tree: `min(d0..d3) < best` is false for almost every group of four (a query's best improves
// O(log n) times), so the compare / select chain of the first version -- three ALU instructions per candidate on top of
// the eight FP32 ones, which made the loop issue-bound -- only runs on those rare groups, in index order with a strict
// '<' (the first minimum still wins).  HBM traffic is the algorithmic minimum (12 B/point in, 8 B/point out); the loop
// is bound by the FMA pipe: 8 lane-ops per (query, candidate).
// Backward: the gather  sum_{j : nn_r(j) == i}  is found per WARP: a warp owns 32 consecutive queries, its lanes test 32
// candidates at a time for "nn_r(j) in my warp's range" (one subtract + compare + ballot per 32 candidates instead of a
// compare per (query, candidate)), and the few hits are applied by the owning lane in ascending j -- the same order, and
// the same bits, as a sequential scan.
#include "common.cuh"
#include "f32x2.cuh"

namespace shwd {

constexpr int CH_THREADS = 256;
constexpr int CH_TILE = 1024;

__device__ __forceinline__ float sqdist(float ax, float ay, float az, float sx, float sy, float sz) {
  float dx = ax - sx, dy = ay - sy, dz = az - sz;
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// the same on two candidates: ((dx*dx + dy*dy) + dz*dz), every product and sum rounded separately
__device__ __forceinline__ f2 sqdist2(f2 ax, f2 ay, f2 az, f2 sx, f2 sy, f2 sz) {
  const f2 dx = sub2(ax, sx), dy = sub2(ay, sy), dz = sub2(az, sz);
  // The products stay SCALAR (__fmul_rn): ptxas contracts mul.rn.f32x2 + add.rn.f32x2 (and fma(d, d, -0) + add) into
  // FFMA2, which drops the product's rounding -- the last bit of the distance and, on near ties, the argmin would differ
  // from torch.  Differences and sums are packed.
  const f2 px = mk2(__fmul_rn(lo2(dx), lo2(dx)), __fmul_rn(hi2(dx), hi2(dx)));
  const f2 py = mk2(__fmul_rn(lo2(dy), lo2(dy)), __fmul_rn(hi2(dy), hi2(dy)));
  const f2 pz = mk2(__fmul_rn(lo2(dz), lo2(dz)), __fmul_rn(hi2(dz), hi2(dz)));
  return add2(add2(px, py), pz);
}

// grid.x = query blocks of direction 0 (x queries) followed by direction 1 (y queries); grid.y = B
// SPLIT > 1 (small grids: one pair, or few short clouds): a CTA owns CH_THREADS / SPLIT queries and SPLIT threads share a
// query, thread `sub` scanning the sub-th slice of every tile -- SPLIT times as many CTAs for the same work, so a single pair
// of 16384-point clouds fills the GPU (1024 CTAs instead of 128; a warp still reads one candidate address: broadcast).
// The SPLIT partial (distance, index) pairs of a query are combined by smallest distance, then smallest index -- each partial
// is already the first minimum of its slice set, so this is exactly the first minimum of the sequential scan.
template <int SPLIT>
__global__ void __launch_bounds__(CH_THREADS) chamfer_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int N,
                                                                 int M, int blocks_x, float* __restrict__ d_xy,
                                                                 int* __restrict__ idx_xy, float* __restrict__ d_yx,
                                                                 int* __restrict__ idx_yx) {
  __shared__ __align__(16) float tx[CH_TILE];
  __shared__ __align__(16) float ty[CH_TILE];
  __shared__ __align__(16) float tz[CH_TILE];
  const int b = blockIdx.y;
  const bool dir = blockIdx.x >= blocks_x;  // false: x queries against y
  const int qb = dir ? blockIdx.x - blocks_x : blockIdx.x;
  const float* q = (dir ? y + (size_t)b * M * 3 : x + (size_t)b * N * 3);
  const float* r = (dir ? x + (size_t)b * N * 3 : y + (size_t)b * M * 3);
  const int nq = dir ? M : N, nr = dir ? N : M;
  constexpr int QPB = CH_THREADS / SPLIT;  // queries per CTA
  constexpr int SLICE = CH_TILE / SPLIT;   // candidates of a tile one thread scans
  const int sub = threadIdx.x / QPB;
  const int i = qb * QPB + (threadIdx.x - sub * QPB);
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (i < nq) {
    ax = __ldg(q + 3 * i);
    ay = __ldg(q + 3 * i + 1);
    az = __ldg(q + 3 * i + 2);
  }
  const f2 ax2 = bc2(ax), ay2 = bc2(ay), az2 = bc2(az);
  float best = INFINITY;
  int bi = 0;
  for (int t0 = 0; t0 < nr; t0 += CH_TILE) {
    const int cnt = min(CH_TILE, nr - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < CH_TILE; j += CH_THREADS) {
      const bool in = j < cnt;  // padding never wins: distance = +inf (or NaN)
      tx[j] = in ? __ldg(r + 3 * (t0 + j)) : INFINITY;
      ty[j] = in ? __ldg(r + 3 * (t0 + j) + 1) : INFINITY;
      tz[j] = in ? __ldg(r + 3 * (t0 + j) + 2) : INFINITY;
    }
    __syncthreads();
    const int lim = min((cnt + 3) & ~3, (sub + 1) * SLICE);
#pragma unroll 2
    for (int j = sub * SLICE; j < lim; j += 4) {
      const float4 X = *reinterpret_cast<const float4*>(tx + j), Y = *reinterpret_cast<const float4*>(ty + j);
      const float4 Z = *reinterpret_cast<const float4*>(tz + j);
      const f2 d01 = sqdist2(ax2, ay2, az2, mk2(X.x, X.y), mk2(Y.x, Y.y), mk2(Z.x, Z.y));
      const f2 d23 = sqdist2(ax2, ay2, az2, mk2(X.z, X.w), mk2(Y.z, Y.w), mk2(Z.z, Z.w));
      const float d0 = lo2(d01), d1 = hi2(d01), d2 = lo2(d23), d3 = hi2(d23);
      if (fminf(fminf(d0, d1), fminf(d2, d3)) < best) {
        // strict '<' in index order keeps the first minimum
        if (d0 < best) { best = d0; bi = t0 + j; }
        if (d1 < best) { best = d1; bi = t0 + j + 1; }
        if (d2 < best) { best = d2; bi = t0 + j + 2; }
        if (d3 < best) { best = d3; bi = t0 + j + 3; }
      }
    }
  }
  if (SPLIT > 1) {
    __shared__ float pbest[CH_THREADS];
    __shared__ int pbi[CH_THREADS];
    pbest[threadIdx.x] = best;
    pbi[threadIdx.x] = bi;
    __syncthreads();
    if (sub != 0) return;
#pragma unroll
    for (int r = 1; r < SPLIT; ++r) {
      const float ob = pbest[r * QPB + threadIdx.x];
      const int oi = pbi[r * QPB + threadIdx.x];
      if (ob < best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
  }
  if (i < nq) {
    if (dir) {
      d_yx[(size_t)b * M + i] = best;
      idx_yx[(size_t)b * M + i] = bi;
    } else {
      d_xy[(size_t)b * N + i] = best;
      idx_xy[(size_t)b * N + i] = bi;
    }
  }
}

// gq_i = 2 g_i (q_i - r_{nn(i)})  +  sum_{j : nn_r(j) == i} 2 h_j (q_i - r_j), the second sum in ascending j
// (deterministic; no float atomics).
// UNIFORM (the fused loss node): every d_xy[b, :] has the same upstream gradient g[b * g_stride] * cx (d_yx: ... * cy) -- the
// reductions of chamfer_distance folded into two scalars -- so no (B,N) / (B,M) gradient tensors exist.
template <bool UNIFORM>
__global__ void __launch_bounds__(CH_THREADS) chamfer_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int N,
                                                                 int M, int blocks_x, const int* __restrict__ idx_xy,
                                                                 const int* __restrict__ idx_yx, const float* __restrict__ gdx,
                                                                 const float* __restrict__ gdy, const float* __restrict__ g,
                                                                 int g_stride, float cx, float cy, float* __restrict__ gx,
                                                                 float* __restrict__ gy) {
  __shared__ float4 tile[CH_TILE];  // (r_j, 2 h_j)
  __shared__ int tidx[CH_TILE];
  const int b = blockIdx.y;
  const bool dir = blockIdx.x >= blocks_x;
  const int qb = dir ? blockIdx.x - blocks_x : blockIdx.x;
  const float* q = (dir ? y + (size_t)b * M * 3 : x + (size_t)b * N * 3);
  const float* r = (dir ? x + (size_t)b * N * 3 : y + (size_t)b * M * 3);
  const int nq = dir ? M : N, nr = dir ? N : M;
  const int* q_nn = dir ? idx_yx + (size_t)b * M : idx_xy + (size_t)b * N;  // nn of each query in r
  const int* r_nn = dir ? idx_xy + (size_t)b * N : idx_yx + (size_t)b * M;  // nn of each r point in q
  const float* gq = UNIFORM ? nullptr : (dir ? gdy + (size_t)b * M : gdx + (size_t)b * N);
  const float* gr = UNIFORM ? nullptr : (dir ? gdx + (size_t)b * N : gdy + (size_t)b * M);
  float uq = 0.f, ur = 0.f;  // UNIFORM: 2 * upstream gradient of every query / every r point
  if (UNIFORM) {
    const float gb = __ldg(g + (size_t)b * g_stride);
    uq = 2.f * (gb * (dir ? cy : cx));
    ur = 2.f * (gb * (dir ? cx : cy));
  }
  float* out = dir ? gy + (size_t)b * M * 3 : gx + (size_t)b * N * 3;
  const int i = qb * CH_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int wbase = i - lane;  // first query of this warp
  float ax = 0.f, ay = 0.f, az = 0.f, ox = 0.f, oy = 0.f, oz = 0.f;
  if (i < nq) {
    ax = __ldg(q + 3 * i);
    ay = __ldg(q + 3 * i + 1);
    az = __ldg(q + 3 * i + 2);
    const int j = __ldg(q_nn + i);
    const float g2 = UNIFORM ? uq : 2.f * __ldg(gq + i);
    ox = g2 * (ax - __ldg(r + 3 * j));
    oy = g2 * (ay - __ldg(r + 3 * j + 1));
    oz = g2 * (az - __ldg(r + 3 * j + 2));
  }
  for (int t0 = 0; t0 < nr; t0 += CH_TILE) {
    const int cnt = min(CH_TILE, nr - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < CH_TILE; j += CH_THREADS) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int k = -1;
      if (j < cnt) {
        v = make_float4(__ldg(r + 3 * (t0 + j)), __ldg(r + 3 * (t0 + j) + 1), __ldg(r + 3 * (t0 + j) + 2),
                        UNIFORM ? ur : 2.f * __ldg(gr + t0 + j));
        k = __ldg(r_nn + t0 + j);
      }
      tile[j] = v;
      tidx[j] = k;
    }
    __syncthreads();
    for (int j0 = 0; j0 < cnt; j0 += 32) {  // padding entries hold -1: never inside [wbase, wbase + 32)
      const int own = tidx[j0 + lane] - wbase;  // which lane of this warp candidate j0 + lane belongs to, if any
      unsigned hits = __ballot_sync(0xffffffffu, (unsigned)own < 32u);
      while (hits) {
        const int l = __ffs(hits) - 1;
        hits &= hits - 1;
        if (__shfl_sync(0xffffffffu, own, l) == lane) {
          const float4 s = tile[j0 + l];
          ox = fmaf(s.w, ax - s.x, ox);
          oy = fmaf(s.w, ay - s.y, oy);
          oz = fmaf(s.w, az - s.z, oz);
        }
      }
    }
  }
  if (i < nq) {
    out[3 * i] = ox;
    out[3 * i + 1] = oy;
    out[3 * i + 2] = oz;
  }
}

// The reductions of pytorch3d's chamfer_distance in one launch: CTA b sums d_xy[b, :] and d_yx[b, :] in a fixed order,
// loss_b = sx * sum_i d_xy + sy * sum_j d_yx; with a batch reduction the last CTA to finish (ticket) adds the loss_b in
// index order and scales -- deterministic, no float atomics.
__device__ __forceinline__ float block_sum_fixed(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < CH_THREADS / 32; ++w) t += red[w];
  return t;
}
__global__ void __launch_bounds__(CH_THREADS) chamfer_reduce_kernel(const float* __restrict__ d_xy, const float* __restrict__ d_yx,
                                                                    int B, int N, int M, float sx, float sy, float sb,
                                                                    int reduce_batch, float* __restrict__ per_pair,
                                                                    unsigned* __restrict__ ticket, float* __restrict__ loss) {
  __shared__ float red[CH_THREADS / 32];
  __shared__ bool last;
  const int b = blockIdx.x;
  float a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < N; i += CH_THREADS) a += __ldg(d_xy + (size_t)b * N + i);
  if (sy != 0.f)
    for (int j = threadIdx.x; j < M; j += CH_THREADS) c += __ldg(d_yx + (size_t)b * M + j);
  a = block_sum_fixed(a, red);
  c = block_sum_fixed(c, red);
  const float lb = (sy != 0.f) ? sx * a + sy * c : sx * a;
  if (!reduce_batch) {
    if (threadIdx.x == 0) loss[b] = lb;
    return;
  }
  if (threadIdx.x == 0) {
    per_pair[b] = lb;
    __threadfence();
    last = atomicAdd(ticket, 1u) == (unsigned)(B - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int k = threadIdx.x; k < B; k += CH_THREADS) t += __ldcg(per_pair + k);
  t = block_sum_fixed(t, red);
  if (threadIdx.x == 0) {
    loss[0] = t * sb;
    *ticket = 0u;  // re-armed for the next call on this workspace
  }
}

}  // namespace shwd

using namespace shwd;

extern "C" int shwd_chamfer_fwd(const float* x, const float* y, int B, int N, int M, float* d_xy, int* idx_xy, float* d_yx,
                                int* idx_yx, void* stream) {
  if (!x || !y || !d_xy || !idx_xy || !d_yx || !idx_yx || B < 0 || N <= 0 || M <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  const int bx = (N + CH_THREADS - 1) / CH_THREADS, by = (M + CH_THREADS - 1) / CH_THREADS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((long long)(bx + by) * B >= 592) {  // at least four 256-thread CTAs per SM: one thread per query fills the GPU
    chamfer_fwd_kernel<1><<<dim3(bx + by, B), CH_THREADS, 0, s>>>(x, y, N, M, bx, d_xy, idx_xy, d_yx, idx_yx);
  } else {
    constexpr int SPLIT = 8, QPB = CH_THREADS / SPLIT;
    const int sx = (N + QPB - 1) / QPB, sy = (M + QPB - 1) / QPB;
    chamfer_fwd_kernel<SPLIT><<<dim3(sx + sy, B), CH_THREADS, 0, s>>>(x, y, N, M, sx, d_xy, idx_xy, d_yx, idx_yx);
  }
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_chamfer_bwd(const float* x, const float* y, int B, int N, int M, const int* idx_xy, const int* idx_yx,
                                const float* gdx, const float* gdy, float* gx, float* gy, void* stream) {
  if (!x || !y || !idx_xy || !idx_yx || !gdx || !gdy || !gx || !gy || B < 0 || N <= 0 || M <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  const int bx = (N + CH_THREADS - 1) / CH_THREADS, by = (M + CH_THREADS - 1) / CH_THREADS;
  chamfer_bwd_kernel<false><<<dim3(bx + by, B), CH_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, y, N, M, bx, idx_xy, idx_yx, gdx,
                                                                                                  gdy, nullptr, 0, 0.f, 0.f, gx, gy);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// chamfer_distance's reductions (pytorch3d semantics as the reference calls it, train_CD.py:123,161; main_rotation.py:203) on
// the nearest-neighbour distances of shwd_chamfer_fwd:  loss_b = sx * sum_i d_xy[b,i] + sy * sum_j d_yx[b,j]  (sx = 1/N or 1,
// sy = 1/M, 1 or 0 for single_directional);  reduce_batch != 0: loss[0] = sb * sum_b loss_b (sb = 1/B or 1), else loss[b].
// ws: B floats + one zero-initialised unsigned (shwd_chamfer_reduce_workspace_bytes), re-armed by the kernel.
extern "C" size_t shwd_chamfer_reduce_workspace_bytes(int B) { return (size_t)(B > 0 ? B : 1) * sizeof(float) + 16; }
extern "C" int shwd_chamfer_reduce(const float* d_xy, const float* d_yx, int B, int N, int M, float sx, float sy, float sb,
                                   int reduce_batch, void* ws, float* loss, void* stream) {
  if (!d_xy || !d_yx || !loss || !ws || B < 0 || N <= 0 || M <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  unsigned* ticket = reinterpret_cast<unsigned*>(ws);
  float* per_pair = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 16);
  chamfer_reduce_kernel<<<B, CH_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(d_xy, d_yx, B, N, M, sx, sy, sb, reduce_batch,
                                                                               per_pair, ticket, loss);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// Backward of the fused loss: d loss / d d_xy[b, :] = g[b * g_stride] * cx, d loss / d d_yx[b, :] = g[b * g_stride] * cy
// (g: the upstream gradient on the device -- one scalar, or one per pair without a batch reduction).
extern "C" int shwd_chamfer_bwd_uniform(const float* x, const float* y, int B, int N, int M, const int* idx_xy, const int* idx_yx,
                                        const float* g, int g_stride, float cx, float cy, float* gx, float* gy, void* stream) {
  if (!x || !y || !idx_xy || !idx_yx || !g || !gx || !gy || B < 0 || N <= 0 || M <= 0 || (g_stride != 0 && g_stride != 1))
    return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  const int bx = (N + CH_THREADS - 1) / CH_THREADS, by = (M + CH_THREADS - 1) / CH_THREADS;
  chamfer_bwd_kernel<true><<<dim3(bx + by, B), CH_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, y, N, M, bx, idx_xy, idx_yx, nullptr,
                                                                                                 nullptr, g, g_stride, cx, cy, gx, gy);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// The learned sphere map phi, Planar variant (`--flow_name Planar`, train_W_COS.py:306,379): a stack of planar flows
//     z <- z + u^ tanh(lin + b),   u^ = u + (log(1 + exp(<w,u>)) - 1 - <w,u>) w / |w|^2
// fused into one forward launch and one backward launch (+ a one-CTA parameter reduction).
//
// Replaces   Norm_Flow_structure.forward ("Planar")   Point_Cloud_Resistration/losses/s2_wasserstein.py:140-143,160-163
//            flows.Planar.forward (act "tanh")         losses/normflows_ishikawa/flows/planar.py:49-60
// which run as ~12 eager torch kernels per flow layer and direction; the log-determinant is discarded by the caller
// (`x, _ = flow(x)`) and is not reproduced.
//
// `lin` is what the reference's `torch.sum(w * z, list(range(1, w.dim())), keepdim=True)` yields with w of shape (1, 3):
// a sum over DIM 1 of the input.
//   * (N, 3) input ("points" layout, clouds == 0):  lin_n = <w, z_n>: the textbook planar flow, one thread per point, the
//     whole stack in registers (12 B read + 12 B written per point).
//   * (B, N, 3) input ("clouds" layout, what the loss wrappers feed, s2_wasserstein.py:243-246): dim 1 is the POINT axis,
//     so lin_bc = w_c sum_n z_bnc -- every layer is a per-cloud, per-coordinate translation d_lbc = u^_lc tanh(w_lc S_lbc +
//     b_l) of the whole cloud, and the column sums follow S_(l+1) = S_l + N d_l.  One CTA per cloud (a cluster of 8 CTAs with
//     a distributed-shared-memory exchange of the partial sums when the clouds are large and few): a fixed-order column sum,
//     a scalar recurrence over the layers on one thread per coordinate, then z + d_1 + ... + d_L applied in the reference's
//     order.  NOTE the recurrence amplifies a perturbation of S_l by 1 + N u^ w (1 - tanh^2) per layer: where that is large
//     (N = 1024, unsaturated layers) the reference's own float32 result is decided by the rounding of its column sum.
//
// Parameters per flow layer (PL_PER_LAYER = 7 raw floats): [u 3 | w 3 | b 1].  The backward returns gradients w.r.t. the
// raw parameters: per-CTA partial sums of (d/du^, direct d/dw, d/db) in the workspace, summed in a fixed order by a second
// one-CTA kernel that also applies the chain through u^ (no float atomics: bit-reproducible).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace shwd {

constexpr int PL_PER_LAYER = 7;
constexpr int PL_MAX_LAYERS = 8;
constexpr int PL_CLOUD_THREADS = 384;  // a multiple of 3: a thread's strided elements all lie in one coordinate column
constexpr int PL_CLUSTER = 8;          // CTAs per cloud on the cluster path (large clouds, few of them)
constexpr int PL_POINT_THREADS = 256;
constexpr int PL_POINT_WARPS = PL_POINT_THREADS / 32;

struct PlanarLayer {
  float uh[3], w[3], b;
};

// u^ of one layer from its raw parameters, in the reference's operation order (planar.py:55-57)
__device__ __forceinline__ PlanarLayer pl_layer(const float* __restrict__ prm) {
  PlanarLayer L;
  const float u0 = __ldg(prm), u1 = __ldg(prm + 1), u2 = __ldg(prm + 2);
  L.w[0] = __ldg(prm + 3);
  L.w[1] = __ldg(prm + 4);
  L.w[2] = __ldg(prm + 5);
  L.b = __ldg(prm + 6);
  const float inner = L.w[0] * u0 + L.w[1] * u1 + L.w[2] * u2;
  const float ww = L.w[0] * L.w[0] + L.w[1] * L.w[1] + L.w[2] * L.w[2];
  const float k = logf(1.f + expf(inner)) - 1.f - inner;
  L.uh[0] = u0 + k * L.w[0] / ww;
  L.uh[1] = u1 + k * L.w[1] / ww;
  L.uh[2] = u2 + k * L.w[2] / ww;
  return L;
}

// Column sums of the points [lo, hi) of one (N, 3) cloud by a CTA of PL_CLOUD_THREADS threads, fixed order: strided
// per-thread partial sums (four loads in flight), a halving tree whose strides stay multiples of 3, then -- on the cluster
// path -- the CTAs' sums added in rank order through distributed shared memory.  Result in red[0..2] on every CTA.
// Accumulated in float64 (3 N additions per cloud: free next to the loads): the recurrence over the layers amplifies an error
// of the sum by up to ~N per layer, so the kernel adds none of its own to the rounding the reference's float32 sum carries.
__device__ __forceinline__ void pl_column_sums(const float* __restrict__ xb, int lo, int hi, double* red, double* slot) {
  constexpr int T = PL_CLOUD_THREADS;
  const float* p = xb + 3 * (size_t)lo;
  const int n3 = 3 * (hi - lo);
  double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
  int e = threadIdx.x;
  for (; e + 3 * T < n3; e += 4 * T) {
    const float v0 = __ldg(p + e), v1 = __ldg(p + e + T), v2 = __ldg(p + e + 2 * T), v3 = __ldg(p + e + 3 * T);
    a0 += (double)v0;
    a1 += (double)v1;
    a2 += (double)v2;
    a3 += (double)v3;
  }
  for (; e < n3; e += T) a0 += (double)__ldg(p + e);
  red[threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
#pragma unroll
  for (int s = T / 2; s >= 3; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned nb = cluster.num_blocks();
  if (nb > 1) {
    if (threadIdx.x < 3) slot[threadIdx.x] = red[threadIdx.x];
    cluster.sync();
    if (threadIdx.x < 3) {
      double v = 0.;
      for (unsigned r = 0; r < nb; ++r) v += cluster.map_shared_rank(slot, r)[threadIdx.x];
      red[threadIdx.x] = v;
    }
    cluster.sync();  // nobody leaves while a peer still reads its slot
  }
}
// this CTA's contiguous share [lo, hi) of the cloud's N points (the whole cloud without a cluster)
__device__ __forceinline__ void pl_cloud_range(int N, int& b, int& lo, int& hi) {
  cg::cluster_group cluster = cg::this_cluster();
  const int nb = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
  b = blockIdx.x / nb;
  const int per = (N + nb - 1) / nb;
  lo = min(r * per, N);
  hi = min(lo + per, N);
}

// ---- clouds layout ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PL_CLOUD_THREADS) planar_clouds_fwd_kernel(const float* __restrict__ x, int N,
                                                                             const float* __restrict__ params, int n_layers,
                                                                             float* __restrict__ y, double* __restrict__ colsum) {
  __shared__ double red[PL_CLOUD_THREADS];
  __shared__ double slot[3];
  __shared__ float delta[PL_MAX_LAYERS][3];
  int b, lo, hi;
  pl_cloud_range(N, b, lo, hi);
  const size_t base = (size_t)b * N * 3;
  pl_column_sums(x + base, lo, hi, red, slot);
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    double S = red[c];  // the column sum stays in float64 along the recurrence; each layer rounds it once, as its input
    if (lo == 0) colsum[3 * (size_t)b + c] = S;
    for (int l = 0; l < n_layers; ++l) {
      const PlanarLayer P = pl_layer(params + l * PL_PER_LAYER);
      const float d = P.uh[c] * tanhf(P.w[c] * (float)S + P.b);
      delta[l][c] = d;
      S += (double)N * (double)d;
    }
  }
  __syncthreads();
  const int c = threadIdx.x % 3;
  float dl[PL_MAX_LAYERS];
#pragma unroll
  for (int l = 0; l < PL_MAX_LAYERS; ++l) dl[l] = l < n_layers ? delta[l][c] : 0.f;
  const float* xi = x + base + 3 * (size_t)lo;
  float* yo = y + base + 3 * (size_t)lo;
  const int n3 = 3 * (hi - lo);
#pragma unroll 4
  for (int e = threadIdx.x; e < n3; e += PL_CLOUD_THREADS) {
    float v = __ldg(xi + e);
#pragma unroll
    for (int l = 0; l < PL_MAX_LAYERS; ++l)
      if (l < n_layers) v += dl[l];  // ((z + d_1) + d_2) + ...: the reference's order
    yo[e] = v;
  }
}

// gx = gy + dL/dS_0 (per column); partial[cloud][l*7 + (0..2: d/du^, 3..5: direct d/dw, 6: d/db)]
__global__ void __launch_bounds__(PL_CLOUD_THREADS) planar_clouds_bwd_kernel(const float* __restrict__ gy,
                                                                             const double* __restrict__ colsum, int N,
                                                                             const float* __restrict__ params, int n_layers,
                                                                             float* __restrict__ gx, float* __restrict__ partial) {
  __shared__ double red[PL_CLOUD_THREADS];
  __shared__ double slot[3];
  __shared__ float gS0[3];
  __shared__ float ga_s[PL_MAX_LAYERS][3];
  int b, lo, hi;
  pl_cloud_range(N, b, lo, hi);
  const size_t base = (size_t)b * N * 3;
  pl_column_sums(gy + base, lo, hi, red, slot);
  float* out = partial + (size_t)b * n_layers * PL_PER_LAYER;
  const bool writer = lo == 0;  // rank 0 of the cluster: the chain is evaluated by every CTA, its sums written once
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    const float GS = (float)red[c];
    float S[PL_MAX_LAYERS], t[PL_MAX_LAYERS];
    double s = colsum[3 * (size_t)b + c];
    for (int l = 0; l < n_layers; ++l) {  // the forward recurrence again (same operations, same bits)
      const PlanarLayer P = pl_layer(params + l * PL_PER_LAYER);
      S[l] = (float)s;
      t[l] = tanhf(P.w[c] * S[l] + P.b);
      s += (double)N * (double)(P.uh[c] * t[l]);
    }
    float gS = 0.f;  // dL/dS_(l+1)
    for (int l = n_layers - 1; l >= 0; --l) {
      const PlanarLayer P = pl_layer(params + l * PL_PER_LAYER);
      const float gd = GS + (float)N * gS;  // d_l reaches every output point of the column and S_(l+1)
      const float ga = gd * P.uh[c] * (1.f - t[l] * t[l]);
      if (writer) {
        out[l * PL_PER_LAYER + c] = gd * t[l];
        out[l * PL_PER_LAYER + 3 + c] = ga * S[l];
      }
      ga_s[l][c] = ga;
      gS += ga * P.w[c];
    }
    gS0[c] = gS;
  }
  __syncthreads();
  if (writer && threadIdx.x < n_layers)
    out[threadIdx.x * PL_PER_LAYER + 6] = (ga_s[threadIdx.x][0] + ga_s[threadIdx.x][1]) + ga_s[threadIdx.x][2];
  const float g0 = gS0[threadIdx.x % 3];
  const float* gi = gy + base + 3 * (size_t)lo;
  float* go = gx + base + 3 * (size_t)lo;
  const int n3 = 3 * (hi - lo);
#pragma unroll 4
  for (int e = threadIdx.x; e < n3; e += PL_CLOUD_THREADS) go[e] = __ldg(gi + e) + g0;
}

// ---- points layout ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PL_POINT_THREADS) planar_points_fwd_kernel(const float* __restrict__ x, int npts,
                                                                             const float* __restrict__ params, int n_layers,
                                                                             float* __restrict__ y) {
  __shared__ PlanarLayer Ls[PL_MAX_LAYERS];
  if (threadIdx.x < n_layers) Ls[threadIdx.x] = pl_layer(params + threadIdx.x * PL_PER_LAYER);
  __syncthreads();
  for (int n = blockIdx.x * PL_POINT_THREADS + threadIdx.x; n < npts; n += gridDim.x * PL_POINT_THREADS) {
    float z0 = __ldg(x + 3 * (size_t)n), z1 = __ldg(x + 3 * (size_t)n + 1), z2 = __ldg(x + 3 * (size_t)n + 2);
    for (int l = 0; l < n_layers; ++l) {
      const PlanarLayer& P = Ls[l];
      const float t = tanhf((P.w[0] * z0 + P.w[1] * z1 + P.w[2] * z2) + P.b);
      z0 += P.uh[0] * t;
      z1 += P.uh[1] * t;
      z2 += P.uh[2] * t;
    }
    y[3 * (size_t)n] = z0;
    y[3 * (size_t)n + 1] = z1;
    y[3 * (size_t)n + 2] = z2;
  }
}

template <int NL>
__global__ void __launch_bounds__(PL_POINT_THREADS) planar_points_bwd_kernel(const float* __restrict__ x,
                                                                             const float* __restrict__ gy, int npts,
                                                                             const float* __restrict__ params,
                                                                             float* __restrict__ gx, float* __restrict__ partial) {
  __shared__ PlanarLayer Ls[NL];
  __shared__ float wsum[PL_POINT_WARPS][NL * PL_PER_LAYER];
  if (threadIdx.x < NL) Ls[threadIdx.x] = pl_layer(params + threadIdx.x * PL_PER_LAYER);
  __syncthreads();
  float acc[NL * PL_PER_LAYER];
#pragma unroll
  for (int i = 0; i < NL * PL_PER_LAYER; ++i) acc[i] = 0.f;
  for (int n = blockIdx.x * PL_POINT_THREADS + threadIdx.x; n < npts; n += gridDim.x * PL_POINT_THREADS) {
    float z[NL + 1][3], t[NL];
    z[0][0] = __ldg(x + 3 * (size_t)n);
    z[0][1] = __ldg(x + 3 * (size_t)n + 1);
    z[0][2] = __ldg(x + 3 * (size_t)n + 2);
    float g0 = __ldg(gy + 3 * (size_t)n), g1 = __ldg(gy + 3 * (size_t)n + 1), g2 = __ldg(gy + 3 * (size_t)n + 2);
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const PlanarLayer& P = Ls[l];
      t[l] = tanhf((P.w[0] * z[l][0] + P.w[1] * z[l][1] + P.w[2] * z[l][2]) + P.b);
      z[l + 1][0] = z[l][0] + P.uh[0] * t[l];
      z[l + 1][1] = z[l][1] + P.uh[1] * t[l];
      z[l + 1][2] = z[l][2] + P.uh[2] * t[l];
    }
#pragma unroll
    for (int l = NL - 1; l >= 0; --l) {
      const PlanarLayer& P = Ls[l];
      const float ga = (g0 * P.uh[0] + g1 * P.uh[1] + g2 * P.uh[2]) * (1.f - t[l] * t[l]);
      float* a = acc + l * PL_PER_LAYER;
      a[0] += g0 * t[l];
      a[1] += g1 * t[l];
      a[2] += g2 * t[l];
      a[3] += ga * z[l][0];
      a[4] += ga * z[l][1];
      a[5] += ga * z[l][2];
      a[6] += ga;
      g0 += ga * P.w[0];
      g1 += ga * P.w[1];
      g2 += ga * P.w[2];
    }
    gx[3 * (size_t)n] = g0;
    gx[3 * (size_t)n + 1] = g1;
    gx[3 * (size_t)n + 2] = g2;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NL * PL_PER_LAYER; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) wsum[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NL * PL_PER_LAYER) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < PL_POINT_WARPS; ++w) v += wsum[w][threadIdx.x];
    partial[(size_t)blockIdx.x * NL * PL_PER_LAYER + threadIdx.x] = v;
  }
}

// Fixed-order sum of the per-CTA partials, then the chain through u^ = u + k(<w,u>) w / |w|^2:
//   A = <g_u^, w>;  d/du = g_u^ + A k' w / |w|^2;  d/dw = direct + g_u^ k / |w|^2 + A (k' u / |w|^2 - 2 k w / |w|^4);
//   k = softplus(<w,u>) - 1 - <w,u>,  k' = sigmoid(<w,u>) - 1.
constexpr int PL_RED_LANES = 16;
__global__ void __launch_bounds__(PL_MAX_LAYERS* PL_PER_LAYER* PL_RED_LANES) planar_reduce_kernel(
    const float* __restrict__ partial, int rows, const float* __restrict__ params, int n_layers, float* __restrict__ gparams) {
  __shared__ float part[PL_RED_LANES][PL_MAX_LAYERS * PL_PER_LAYER];
  __shared__ float tot[PL_MAX_LAYERS * PL_PER_LAYER];
  const int np = n_layers * PL_PER_LAYER;
  const int col = threadIdx.x % (PL_MAX_LAYERS * PL_PER_LAYER), lane = threadIdx.x / (PL_MAX_LAYERS * PL_PER_LAYER);
  if (col < np) {
    float v = 0.f;
    for (int r = lane; r < rows; r += PL_RED_LANES) v += __ldg(partial + (size_t)r * np + col);
    part[lane][col] = v;
  }
  __syncthreads();
  if (threadIdx.x < np) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < PL_RED_LANES; ++j) v += part[j][threadIdx.x];
    tot[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x < n_layers) {
    const float* prm = params + threadIdx.x * PL_PER_LAYER;
    const float* g = tot + threadIdx.x * PL_PER_LAYER;
    float* o = gparams + threadIdx.x * PL_PER_LAYER;
    const float u[3] = {__ldg(prm), __ldg(prm + 1), __ldg(prm + 2)}, w[3] = {__ldg(prm + 3), __ldg(prm + 4), __ldg(prm + 5)};
    const float inner = w[0] * u[0] + w[1] * u[1] + w[2] * u[2];
    const float ww = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const float ex = expf(inner);
    const float k = logf(1.f + ex) - 1.f - inner;
    const float kp = -1.f / (1.f + ex);  // sigmoid(inner) - 1
    const float A = g[0] * w[0] + g[1] * w[1] + g[2] * w[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = g[c] + A * kp * w[c] / ww;
      o[3 + c] = g[3 + c] + g[c] * k / ww + A * (kp * u[c] / ww - 2.f * k * w[c] / (ww * ww));
    }
    o[6] = g[6];
  }
}

static int pl_point_blocks(int npts) {
  const int want = (npts + PL_POINT_THREADS - 1) / PL_POINT_THREADS;
  const int cap = 4 * sm_count();
  return want < cap ? want : cap;
}

// a cluster of PL_CLUSTER CTAs per cloud when one CTA per cloud would leave most of the GPU idle on a large cloud
static bool pl_use_cluster(int clouds, int N) { return N >= 8192 && clouds * PL_CLUSTER <= 2 * sm_count(); }
template <typename... KArgs, typename... Args>
static cudaError_t pl_launch_clouds(void (*kernel)(KArgs...), int clouds, bool cluster, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster ? clouds * PL_CLUSTER : clouds);
  cfg.blockDim = dim3(PL_CLOUD_THREADS);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster ? PL_CLUSTER : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <int NL>
static void pl_launch_points_bwd(int blocks, cudaStream_t s, const float* x, const float* gy, int npts, const float* params,
                                 float* gx, float* partial) {
  planar_points_bwd_kernel<NL><<<blocks, PL_POINT_THREADS, 0, s>>>(x, gy, npts, params, gx, partial);
}

}  // namespace shwd

using namespace shwd;

extern "C" int shwd_planar_max_layers(void) { return PL_MAX_LAYERS; }
extern "C" int shwd_planar_params_per_layer(void) { return PL_PER_LAYER; }

extern "C" size_t shwd_planar_workspace_bytes(int clouds, int npts, int n_layers) {
  if (clouds < 0 || npts <= 0 || n_layers <= 0) return 0;
  const size_t rows = clouds > 0 ? (size_t)clouds : (size_t)pl_point_blocks(npts);
  return rows * (size_t)n_layers * PL_PER_LAYER * sizeof(float);
}

extern "C" int shwd_planar_fwd(const float* x, int clouds, int npts, const float* params, int n_layers, float* y, double* colsum,
                               void* stream) {
  if (!x || !params || !y || clouds < 0 || npts < 0 || n_layers <= 0 || n_layers > PL_MAX_LAYERS || (clouds > 0 && !colsum))
    return SHWD_ERR_INVALID_ARGUMENT;
  if (npts == 0) return SHWD_OK;
  if (npts > 0x7fffffff / 3) return SHWD_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (clouds > 0)
    SHWD_CUDA_CHECK(pl_launch_clouds(planar_clouds_fwd_kernel, clouds, pl_use_cluster(clouds, npts), s, x, npts, params, n_layers, y,
                                     colsum));
  else
    planar_points_fwd_kernel<<<pl_point_blocks(npts), PL_POINT_THREADS, 0, s>>>(x, npts, params, n_layers, y);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_planar_bwd(const float* x, const float* gy, const double* colsum, int clouds, int npts, const float* params,
                               int n_layers, float* gx, float* gparams, void* workspace, size_t workspace_bytes, void* stream) {
  if (!gy || !params || !gx || !gparams || clouds < 0 || npts < 0 || n_layers <= 0 || n_layers > PL_MAX_LAYERS ||
      (clouds > 0 ? !colsum : !x))
    return SHWD_ERR_INVALID_ARGUMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int np = n_layers * PL_PER_LAYER;
  if (npts == 0) {
    SHWD_CUDA_CHECK(cudaMemsetAsync(gparams, 0, np * sizeof(float), s));
    return SHWD_OK;
  }
  if (npts > 0x7fffffff / 3) return SHWD_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < shwd_planar_workspace_bytes(clouds, npts, n_layers) ||
      (reinterpret_cast<uintptr_t>(workspace) & 3))
    return SHWD_ERR_WORKSPACE;
  float* partial = static_cast<float*>(workspace);
  int rows;
  if (clouds > 0) {
    rows = clouds;
    SHWD_CUDA_CHECK(pl_launch_clouds(planar_clouds_bwd_kernel, clouds, pl_use_cluster(clouds, npts), s, gy, colsum, npts, params,
                                     n_layers, gx, partial));
  } else {
    rows = pl_point_blocks(npts);
    switch (n_layers) {
      case 1: pl_launch_points_bwd<1>(rows, s, x, gy, npts, params, gx, partial); break;
      case 2: pl_launch_points_bwd<2>(rows, s, x, gy, npts, params, gx, partial); break;
      case 3: pl_launch_points_bwd<3>(rows, s, x, gy, npts, params, gx, partial); break;
      case 4: pl_launch_points_bwd<4>(rows, s, x, gy, npts, params, gx, partial); break;
      case 5: pl_launch_points_bwd<5>(rows, s, x, gy, npts, params, gx, partial); break;
      case 6: pl_launch_points_bwd<6>(rows, s, x, gy, npts, params, gx, partial); break;
      case 7: pl_launch_points_bwd<7>(rows, s, x, gy, npts, params, gx, partial); break;
      default: pl_launch_points_bwd<8>(rows, s, x, gy, npts, params, gx, partial); break;
    }
  }
  SHWD_CUDA_CHECK(cudaGetLastError());
  planar_reduce_kernel<<<1, PL_MAX_LAYERS * PL_PER_LAYER * PL_RED_LANES, 0, s>>>(partial, rows, params, n_layers, gparams);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

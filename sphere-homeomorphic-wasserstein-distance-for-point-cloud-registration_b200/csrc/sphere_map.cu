// Sphere map: centroid subtraction + projection on the unit sphere, forward and backward.
//
// Replaces   x - mean(x, dim=1)                      Point_Cloud_Resistration/train_W_COS.py:167-168
//            x / max(||x||_2, 1e-8)                  inside F.cosine_similarity, losses/s2_wasserstein.py:122
//            sum_{n} | ||x_n|| - 1 |                 regularization_of_normalizing_flow, s2_wasserstein.py:224-232
//
// Layout: input (B,N,3) float32 AoS (12 B/point) is read as three float4 per four points (fully used 128-bit
// loads, each warp touching one contiguous 1536 B span); output is the packed float4 record
// (x^0, x^1, x^2, 1/max(||x_c||,1e-8)) that the OT kernels stage into shared memory unchanged.
// One CTA per cloud; the centroid is a warp-shuffle + shared-memory tree reduction.  HBM-bound:
// 12 B read (x2 when centring: the second pass hits L1/L2) + 16 B written per point.
// Few large clouds (a single pair of 16384 ... 65536 points: cfg4 / the top of cfg5) would leave the GPU to one or two SMs:
// those launches use one thread-block CLUSTER of 8 CTAs per cloud -- every CTA owns a contiguous eighth of the points, the
// per-CTA partial sums (centroid, regulariser, gradient mean) are exchanged through distributed shared memory and added in
// rank order by every CTA (deterministic, no scratch memory, no second launch).
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace shwd {

constexpr int SM_THREADS = 256;
constexpr int SM_CLUSTER = 8;  // CTAs per cloud on the cluster path

// Sum of one float4 per CTA over the CTAs of this cluster, in rank order (a plain copy when the launch has no cluster).
__device__ __forceinline__ float4 cluster_sum4(float4 mine, float4* slot) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned nb = cluster.num_blocks();
  if (nb == 1) return mine;
  __syncthreads();  // (slot may still be read by a previous call's tail in this CTA)
  if (threadIdx.x == 0) *slot = mine;
  cluster.sync();
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  for (unsigned r = 0; r < nb; ++r) {
    const float4 v = *cluster.map_shared_rank(slot, r);
    t.x += v.x;
    t.y += v.y;
    t.z += v.z;
    t.w += v.w;
  }
  cluster.sync();  // nobody leaves (or overwrites its slot) while a peer still reads it
  return t;
}
// this CTA's contiguous share [lo, hi) of `count` items
__device__ __forceinline__ void cluster_range(int count, int& lo, int& hi) {
  cg::cluster_group cluster = cg::this_cluster();
  const int nb = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
  const int per = (count + nb - 1) / nb;
  lo = min(r * per, count);
  hi = min(lo + per, count);
}

struct P4 {
  float3 p[4];
};

__device__ __forceinline__ P4 load4(const float* base, int quad) {
  const float4* v = reinterpret_cast<const float4*>(base) + 3 * (size_t)quad;
  float4 a = __ldg(v), b = __ldg(v + 1), c = __ldg(v + 2);
  P4 r;
  r.p[0] = make_float3(a.x, a.y, a.z);
  r.p[1] = make_float3(a.w, b.x, b.y);
  r.p[2] = make_float3(b.z, b.w, c.x);
  r.p[3] = make_float3(c.y, c.z, c.w);
  return r;
}
__device__ __forceinline__ void store4(float* base, int quad, const P4& r) {
  float4* v = reinterpret_cast<float4*>(base) + 3 * (size_t)quad;
  v[0] = make_float4(r.p[0].x, r.p[0].y, r.p[0].z, r.p[1].x);
  v[1] = make_float4(r.p[1].y, r.p[1].z, r.p[2].x, r.p[2].y);
  v[2] = make_float4(r.p[2].z, r.p[3].x, r.p[3].y, r.p[3].z);
}

__device__ __forceinline__ float3 block_sum3(float3 v, float* sm /* 3*32 floats */) {
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  v.z = warp_sum(v.z);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) {
    sm[w] = v.x;
    sm[32 + w] = v.y;
    sm[64 + w] = v.z;
  }
  __syncthreads();
  int nw = blockDim.x >> 5;
  float3 r;
  r.x = (l < nw) ? sm[l] : 0.f;
  r.y = (l < nw) ? sm[32 + l] : 0.f;
  r.z = (l < nw) ? sm[64 + l] : 0.f;
  r.x = warp_sum(r.x);
  r.y = warp_sum(r.y);
  r.z = warp_sum(r.z);
  return r;
}

__device__ __forceinline__ float norm3(float3 v) {
  // same association as ATen's vector_norm over a length-3 inner dim on CPU
  return sqrtf(fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x)));
}

template <bool VEC>
__global__ void __launch_bounds__(SM_THREADS) sphere_map_fwd_kernel(const float* __restrict__ x, float4* __restrict__ xh4,
                                                                    float* __restrict__ reg_out, int N, int flags) {
  __shared__ float red[96];
  __shared__ float4 cslot;
  const int b = blockIdx.x / (int)cg::this_cluster().num_blocks();
  const float* xb = x + (size_t)b * N * 3;
  float4* ob = xh4 + (size_t)b * N;
  const bool center = flags & SHWD_MAP_CENTER, normalize = flags & SHWD_MAP_NORMALIZE;
  float3 mean = make_float3(0.f, 0.f, 0.f);
  float reg = 0.f;
  int lo, hi;  // this CTA's quads (VEC) or points
  cluster_range(VEC ? N / 4 : N, lo, hi);
  if (center || reg_out) {
    float3 s = make_float3(0.f, 0.f, 0.f);
    if (VEC) {
      for (int q = lo + threadIdx.x; q < hi; q += SM_THREADS) {
        P4 r = load4(xb, q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s.x += r.p[e].x;
          s.y += r.p[e].y;
          s.z += r.p[e].z;
          reg += fabsf(norm3(r.p[e]) - 1.f);
        }
      }
    } else {
      for (int n = lo + threadIdx.x; n < hi; n += SM_THREADS) {
        float3 p = make_float3(__ldg(xb + 3 * n), __ldg(xb + 3 * n + 1), __ldg(xb + 3 * n + 2));
        s.x += p.x;
        s.y += p.y;
        s.z += p.z;
        reg += fabsf(norm3(p) - 1.f);
      }
    }
    float3 tot = block_sum3(s, red);
    float rsum = 0.f;
    if (reg_out) rsum = block_sum3(make_float3(reg, 0.f, 0.f), red).x;
    const float4 all = cluster_sum4(make_float4(tot.x, tot.y, tot.z, rsum), &cslot);
    if (center) mean = make_float3(all.x / N, all.y / N, all.z / N);
    if (reg_out && threadIdx.x == 0 && cg::this_cluster().block_rank() == 0) reg_out[b] = all.w;
  }
  auto map = [&](float3 p) -> float4 {
    p.x -= mean.x;
    p.y -= mean.y;
    p.z -= mean.z;
    float inv = 1.f;
    if (normalize) {
      float d = fmaxf(norm3(p), 1e-8f);
      inv = 1.f / d;
      p.x = p.x / d;  // true division, as torch does (x / max(norm, eps))
      p.y = p.y / d;
      p.z = p.z / d;
    }
    return make_float4(p.x, p.y, p.z, inv);
  };
  if (VEC) {
    for (int q = lo + threadIdx.x; q < hi; q += SM_THREADS) {
      P4 r = load4(xb, q);
#pragma unroll
      for (int e = 0; e < 4; ++e) ob[4 * q + e] = map(r.p[e]);
    }
  } else {
    for (int n = lo + threadIdx.x; n < hi; n += SM_THREADS)
      ob[n] = map(make_float3(__ldg(xb + 3 * n), __ldg(xb + 3 * n + 1), __ldg(xb + 3 * n + 2)));
  }
}

// Backward: g4 = dL/d(x^) -> dL/dx through the normalisation Jacobian (I - x^ x^T)/||x_c|| and the centring
// Jacobian (I - 11^T/N), plus the regulariser's sub-gradient sign(||x||-1) x/||x||.
template <bool VEC>
__global__ void __launch_bounds__(SM_THREADS) sphere_map_bwd_kernel(const float* __restrict__ x, const float4* __restrict__ xh4,
                                                                    const float4* __restrict__ g4, const float* __restrict__ greg,
                                                                    float* __restrict__ gx, int N, int flags) {
  __shared__ float red[96];
  __shared__ float4 cslot;
  const int b = blockIdx.x / (int)cg::this_cluster().num_blocks();
  int lo, hi;  // this CTA's quads (VEC) or points
  cluster_range(VEC ? N / 4 : N, lo, hi);
  const int plo = VEC ? 4 * lo : lo, phi = VEC ? 4 * hi : hi;
  const float* xb = x + (size_t)b * N * 3;
  const float4* hb = xh4 + (size_t)b * N;
  const float4* gb = g4 ? g4 + (size_t)b * N : nullptr;
  float* ob = gx + (size_t)b * N * 3;
  const bool center = flags & SHWD_MAP_CENTER, normalize = flags & SHWD_MAP_NORMALIZE;
  const float gr = greg ? greg[b] : 0.f;

  auto gc_of = [&](int n) -> float3 {  // gradient w.r.t. the centred point
    float3 r = make_float3(0.f, 0.f, 0.f);
    if (gb) {
      float4 g = gb[n];
      r = make_float3(g.x, g.y, g.z);
      if (normalize) {
        float4 h = hb[n];
        if (h.w < 1e8f) {  // regular branch: ||x_c|| > 1e-8
          float d = h.x * g.x + h.y * g.y + h.z * g.z;
          r = make_float3((g.x - h.x * d) * h.w, (g.y - h.y * d) * h.w, (g.z - h.z * d) * h.w);
        } else {  // clamp active: x^ = x_c / 1e-8
          r = make_float3(g.x * h.w, g.y * h.w, g.z * h.w);
        }
      }
    }
    return r;
  };
  auto greg_of = [&](float3 p) -> float3 {  // d/dp | ||p|| - 1 |
    float nr = norm3(p);
    float s = (nr > 1.f) ? 1.f : ((nr < 1.f) ? -1.f : 0.f);
    float k = (nr > 0.f) ? gr * s / nr : 0.f;
    return make_float3(k * p.x, k * p.y, k * p.z);
  };

  float3 mean = make_float3(0.f, 0.f, 0.f);
  if (center && gb) {
    float3 s = make_float3(0.f, 0.f, 0.f);
    for (int n = plo + threadIdx.x; n < phi; n += SM_THREADS) {
      float3 g = gc_of(n);
      s.x += g.x;
      s.y += g.y;
      s.z += g.z;
    }
    float3 tot = block_sum3(s, red);
    const float4 all = cluster_sum4(make_float4(tot.x, tot.y, tot.z, 0.f), &cslot);
    mean = make_float3(all.x / N, all.y / N, all.z / N);
  }
  if (VEC) {
    for (int q = lo + threadIdx.x; q < hi; q += SM_THREADS) {
      P4 out;
      P4 raw;
      if (greg) raw = load4(xb, q);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float3 g = gc_of(4 * q + e);
        g.x -= mean.x;
        g.y -= mean.y;
        g.z -= mean.z;
        if (greg) {
          float3 t = greg_of(raw.p[e]);
          g.x += t.x;
          g.y += t.y;
          g.z += t.z;
        }
        out.p[e] = g;
      }
      store4(ob, q, out);
    }
  } else {
    for (int n = lo + threadIdx.x; n < hi; n += SM_THREADS) {
      float3 g = gc_of(n);
      g.x -= mean.x;
      g.y -= mean.y;
      g.z -= mean.z;
      if (greg) {
        float3 t = greg_of(make_float3(__ldg(xb + 3 * n), __ldg(xb + 3 * n + 1), __ldg(xb + 3 * n + 2)));
        g.x += t.x;
        g.y += t.y;
        g.z += t.z;
      }
      ob[3 * n] = g.x;
      ob[3 * n + 1] = g.y;
      ob[3 * n + 2] = g.z;
    }
  }
}

}  // namespace shwd

using namespace shwd;

static bool vec_ok(const void* p, int N) { return (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0); }
// a cluster of SM_CLUSTER CTAs per cloud when one CTA per cloud would leave most of the GPU idle on a large cloud
static bool use_cluster(int B, int N) { return N >= 8192 && B * SM_CLUSTER <= 2 * sm_count(); }
template <typename... KArgs, typename... Args>
static cudaError_t launch_map(void (*kernel)(KArgs...), int B, bool cluster, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster ? B * SM_CLUSTER : B);
  cfg.blockDim = dim3(SM_THREADS);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster ? SM_CLUSTER : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

extern "C" int shwd_sphere_map_fwd(const float* x, float* xh4, float* reg_out, int B, int N, int flags, void* stream) {
  if (!x || !xh4 || B < 0 || N <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool cl = use_cluster(B, N);
  if (vec_ok(x, N))
    SHWD_CUDA_CHECK(launch_map(sphere_map_fwd_kernel<true>, B, cl, s, x, reinterpret_cast<float4*>(xh4), reg_out, N, flags));
  else
    SHWD_CUDA_CHECK(launch_map(sphere_map_fwd_kernel<false>, B, cl, s, x, reinterpret_cast<float4*>(xh4), reg_out, N, flags));
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_sphere_map_bwd(const float* x, const float* xh4, const float* g4, const float* greg, float* gx, int B,
                                   int N, int flags, void* stream) {
  if (!x || !xh4 || !gx || B < 0 || N <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool cl = use_cluster(B, N);
  if (vec_ok(x, N) && vec_ok(gx, N))
    SHWD_CUDA_CHECK(launch_map(sphere_map_bwd_kernel<true>, B, cl, s, x, reinterpret_cast<const float4*>(xh4),
                               reinterpret_cast<const float4*>(g4), greg, gx, N, flags));
  else
    SHWD_CUDA_CHECK(launch_map(sphere_map_bwd_kernel<false>, B, cl, s, x, reinterpret_cast<const float4*>(xh4),
                               reinterpret_cast<const float4*>(g4), greg, gx, N, flags));
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

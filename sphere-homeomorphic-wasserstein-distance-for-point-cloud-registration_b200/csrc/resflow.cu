// The learned sphere map phi of the SHWD loss: a stack of Residual flows x <- x + LipschitzMLP(x),
// fused into one forward and one backward launch.
//
// Replaces   Norm_Flow_structure.forward ("Residual")   Point_Cloud_Resistration/losses/s2_wasserstein.py:144-163
//            flows.Residual / iResBlock forward map      losses/normflows_ishikawa/flows/residual.py:63-68,118-124
//            nets.LipschitzMLP                           losses/normflows_ishikawa/nets/lipschitz.py:14-68
//            Swish                                       nets/lipschitz.py:642-648   x * sigmoid(x * softplus(beta)) / 1.1
// The reference runs it as ~45 eager torch kernels per flow layer and direction (7 x [Swish, Linear]), twice per
// training step (inner ascent + outer evaluation); the log-determinant estimator it also computes is discarded
// (`x, _ = flow(x)`) and is not reproduced.
//
// The kernels take the RAW parameters (weights, biases, Swish betas) plus the frozen power-iteration vectors u_k, v_k and
// form the effective ones themselves, per CTA: W_k / f_k with f_k = max(1, (u_k^T W_k v_k) / coeff)
// (InducedNormLinear.compute_weight(update=False), nets/lipschitz.py:223-274 -- the reference never updates u, v) and
// s_k = softplus(beta_k).  The backward returns gradients w.r.t. the raw parameters (the chain through f_k and the
// softplus is applied in the final reduction kernel), so a training step needs no eager parameter preparation at all.
//
// Layout per flow layer (RF_PER_LAYER floats):  [W0 8x3 | b0 8 | W1..W5 8x8 | b 8 each | W6 3x8 | b6 3 | beta0..beta6];
// u/v buffer per flow layer (RF_UV_PER_LAYER floats): [u0 8 | v0 3 | (u 8 | v 8) x5 | u6 3 | v6 8].
// One thread per point; activations never leave registers.  Parameter gradients are reduced deterministically:
// warp shuffle -> per-warp shared accumulators -> per-CTA partials in the workspace -> a second kernel sums the
// partials in a fixed order (no float atomics).
#include "common.cuh"

namespace shwd {

constexpr int RF_THREADS = 128;
constexpr int RF_WARPS = RF_THREADS / 32;
constexpr int RF_HID = 8, RF_DIM = 3, RF_LIN = 7;
constexpr int RF_W0 = 0;                                   // 8x3
constexpr int RF_B0 = RF_W0 + RF_HID * RF_DIM;             // 8
constexpr int RF_WH = RF_B0 + RF_HID;                      // 5 x (8x8 + 8)
constexpr int RF_WL = RF_WH + 5 * (RF_HID * RF_HID + RF_HID);  // 3x8
constexpr int RF_BL = RF_WL + RF_DIM * RF_HID;             // 3
constexpr int RF_S = RF_BL + RF_DIM;                       // 7 swish scales
constexpr int RF_PER_LAYER = RF_S + RF_LIN;                // 426
constexpr int RF_MAX_LAYERS = 8;
constexpr int RF_UV_PER_LAYER = (RF_HID + RF_DIM) + 5 * (2 * RF_HID) + (RF_DIM + RF_HID);  // 102

// Geometry of linear layer k (0..6) inside a flow layer's parameter / uv blocks.
__device__ __forceinline__ void rf_lin_geom(int k, int& w_off, int& b_off, int& out, int& in, int& u_off, int& v_off) {
  if (k == 0) {
    w_off = RF_W0; b_off = RF_B0; out = RF_HID; in = RF_DIM; u_off = 0; v_off = RF_HID;
  } else if (k == 6) {
    w_off = RF_WL; b_off = RF_BL; out = RF_DIM; in = RF_HID; u_off = (RF_HID + RF_DIM) + 5 * 2 * RF_HID; v_off = u_off + RF_DIM;
  } else {
    w_off = RF_WH + (k - 1) * (RF_HID * RF_HID + RF_HID); b_off = w_off + RF_HID * RF_HID; out = RF_HID; in = RF_HID;
    u_off = (RF_HID + RF_DIM) + (k - 1) * 2 * RF_HID; v_off = u_off + RF_HID;
  }
}
// sigma_k = u^T W v and f_k = max(1, sigma_k / coeff) of linear layer (f, k)
__device__ __forceinline__ float rf_sigma(const float* __restrict__ raw, const float* __restrict__ uv, int f, int k) {
  int w_off, b_off, out, in, u_off, v_off;
  rf_lin_geom(k, w_off, b_off, out, in, u_off, v_off);
  const float* W = raw + f * RF_PER_LAYER + w_off;
  const float* u = uv + f * RF_UV_PER_LAYER + u_off;
  const float* v = uv + f * RF_UV_PER_LAYER + v_off;
  float sg = 0.f;
  for (int i = 0; i < out; ++i) {
    float r = 0.f;
    for (int j = 0; j < in; ++j) r = fmaf(__ldg(W + i * in + j), __ldg(v + j), r);
    sg = fmaf(__ldg(u + i), r, sg);
  }
  return sg;
}
__device__ __forceinline__ float rf_softplus(float b) { return b > 20.f ? b : log1pf(expf(b)); }  // torch threshold 20
// which linear layer a parameter slot belongs to (-1: bias / beta slots)
__device__ __forceinline__ int rf_weight_layer(int i) {
  if (i < RF_B0) return 0;
  if (i < RF_WH) return -1;
  if (i < RF_WL) {
    const int r = (i - RF_WH) % (RF_HID * RF_HID + RF_HID);
    return r < RF_HID * RF_HID ? 1 + (i - RF_WH) / (RF_HID * RF_HID + RF_HID) : -1;
  }
  if (i < RF_BL) return 6;
  return -1;
}
// Effective parameters into shared memory: sP (n_layers * RF_PER_LAYER), sF (n_layers * 7 factors).  All threads call.
__device__ __forceinline__ void rf_load_params(const float* __restrict__ raw, const float* __restrict__ uv, int n_layers, float coeff,
                                               float* sP, float* sF) {
  for (int t = threadIdx.x; t < n_layers * RF_LIN; t += blockDim.x)
    sF[t] = fmaxf(rf_sigma(raw, uv, t / RF_LIN, t % RF_LIN) / coeff, 1.f);
  __syncthreads();
  for (int i = threadIdx.x; i < n_layers * RF_PER_LAYER; i += blockDim.x) {
    const int f = i / RF_PER_LAYER, r = i % RF_PER_LAYER;
    const float p = __ldg(raw + i);
    const int k = rf_weight_layer(r);
    sP[i] = (k >= 0) ? p / sF[f * RF_LIN + k] : (r >= RF_S ? rf_softplus(p) : p);
  }
  __syncthreads();
}

__device__ __forceinline__ float rf_sigmoid(float z) { return 1.f / (1.f + expf(-z)); }
// swish(v) = v * sigmoid(v s) / 1.1
__device__ __forceinline__ float rf_swish(float v, float s) { return v * rf_sigmoid(v * s) / 1.1f; }

// One LipschitzMLP application: out(3) = net(x(3)).  If ACT, the inputs of every linear layer (after Swish) and the
// pre-Swish values are kept for the backward.
struct RfAct {
  float pre0[RF_DIM];            // input of Swish 0 (= x)
  float pre[6][RF_HID];          // inputs of Swish 1..6 (= outputs of linear 0..5)
};

__device__ __forceinline__ void rf_mlp(const float* __restrict__ P, const float (&x)[RF_DIM], float (&out)[RF_DIM], RfAct* act) {
  float a3[RF_DIM];
#pragma unroll
  for (int j = 0; j < RF_DIM; ++j) {
    if (act) act->pre0[j] = x[j];
    a3[j] = rf_swish(x[j], P[RF_S + 0]);
  }
  float h[RF_HID];
#pragma unroll
  for (int i = 0; i < RF_HID; ++i) {
    float v = P[RF_B0 + i];
#pragma unroll
    for (int j = 0; j < RF_DIM; ++j) v = fmaf(P[RF_W0 + i * RF_DIM + j], a3[j], v);
    h[i] = v;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float* W = P + RF_WH + k * (RF_HID * RF_HID + RF_HID);
    const float* b = W + RF_HID * RF_HID;
    float a[RF_HID], o[RF_HID];
#pragma unroll
    for (int j = 0; j < RF_HID; ++j) {
      if (act) act->pre[k][j] = h[j];
      a[j] = rf_swish(h[j], P[RF_S + 1 + k]);
    }
#pragma unroll
    for (int i = 0; i < RF_HID; ++i) {
      float v = b[i];
#pragma unroll
      for (int j = 0; j < RF_HID; ++j) v = fmaf(W[i * RF_HID + j], a[j], v);
      o[i] = v;
    }
#pragma unroll
    for (int i = 0; i < RF_HID; ++i) h[i] = o[i];
  }
  float a[RF_HID];
#pragma unroll
  for (int j = 0; j < RF_HID; ++j) {
    if (act) act->pre[5][j] = h[j];
    a[j] = rf_swish(h[j], P[RF_S + 6]);
  }
#pragma unroll
  for (int i = 0; i < RF_DIM; ++i) {
    float v = P[RF_BL + i];
#pragma unroll
    for (int j = 0; j < RF_HID; ++j) v = fmaf(P[RF_WL + i * RF_HID + j], a[j], v);
    out[i] = v;
  }
}

__global__ void __launch_bounds__(RF_THREADS) resflow_fwd_kernel(const float* __restrict__ x, int npts, const float* __restrict__ params,
                                                                 const float* __restrict__ uv, int n_layers, float coeff,
                                                                 float* __restrict__ y) {
  extern __shared__ float sP[];
  __shared__ float sF[RF_MAX_LAYERS * RF_LIN];
  rf_load_params(params, uv, n_layers, coeff, sP, sF);
  const int n = blockIdx.x * RF_THREADS + threadIdx.x;
  if (n >= npts) return;
  float v[RF_DIM] = {__ldg(x + 3 * (size_t)n), __ldg(x + 3 * (size_t)n + 1), __ldg(x + 3 * (size_t)n + 2)};
  for (int f = 0; f < n_layers; ++f) {
    float o[RF_DIM];
    rf_mlp(sP + f * RF_PER_LAYER, v, o, nullptr);
#pragma unroll
    for (int j = 0; j < RF_DIM; ++j) v[j] += o[j];
  }
  y[3 * (size_t)n] = v[0];
  y[3 * (size_t)n + 1] = v[1];
  y[3 * (size_t)n + 2] = v[2];
}

// d swish / d v and d swish / d s at (v, s)
__device__ __forceinline__ void rf_dswish(float v, float s, float& dv, float& ds) {
  const float sg = rf_sigmoid(v * s);
  const float t = sg * (1.f - sg);
  dv = (sg + v * s * t) / 1.1f;
  ds = v * v * t / 1.1f;
}

// acc[idx] += sum over the warp of val (lane 0 adds into the warp's shared accumulator)
__device__ __forceinline__ void rf_acc(float* wacc, int idx, float val) {
  val = warp_sum(val);
  if ((threadIdx.x & 31) == 0) wacc[idx] += val;
}

__global__ void __launch_bounds__(RF_THREADS) resflow_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, int npts,
                                                                 const float* __restrict__ params, const float* __restrict__ uv,
                                                                 int n_layers, float coeff, float* __restrict__ gx,
                                                                 float* __restrict__ partial) {
  extern __shared__ float smem[];
  __shared__ float sF[RF_MAX_LAYERS * RF_LIN];
  const int np = n_layers * RF_PER_LAYER;
  float* sP = smem;                    // effective parameters
  float* sA = smem + np;               // RF_WARPS x np per-warp accumulators
  for (int i = threadIdx.x; i < RF_WARPS * np; i += RF_THREADS) sA[i] = 0.f;
  rf_load_params(params, uv, n_layers, coeff, sP, sF);
  float* wacc = sA + (threadIdx.x >> 5) * np;
  const int n = blockIdx.x * RF_THREADS + threadIdx.x;
  const bool live = n < npts;
  // forward: keep the input of every flow layer
  float xin[RF_MAX_LAYERS][RF_DIM];
  float v[RF_DIM] = {0.f, 0.f, 0.f};
  if (live) {
    v[0] = __ldg(x + 3 * (size_t)n);
    v[1] = __ldg(x + 3 * (size_t)n + 1);
    v[2] = __ldg(x + 3 * (size_t)n + 2);
  }
#pragma unroll
  for (int f = 0; f < RF_MAX_LAYERS; ++f) {
    if (f < n_layers) {
#pragma unroll
      for (int j = 0; j < RF_DIM; ++j) xin[f][j] = v[j];
      float o[RF_DIM];
      rf_mlp(sP + f * RF_PER_LAYER, v, o, nullptr);
#pragma unroll
      for (int j = 0; j < RF_DIM; ++j) v[j] += o[j];
    }
  }
  float g[RF_DIM] = {0.f, 0.f, 0.f};  // d loss / d (output of the current flow layer); dead lanes contribute zeros
  if (live) {
    g[0] = __ldg(gy + 3 * (size_t)n);
    g[1] = __ldg(gy + 3 * (size_t)n + 1);
    g[2] = __ldg(gy + 3 * (size_t)n + 2);
  }
#pragma unroll
  for (int fr = 0; fr < RF_MAX_LAYERS; ++fr) {
    const int f = n_layers - 1 - fr;
    if (f < 0) continue;
    const float* P = sP + f * RF_PER_LAYER;
    float* A = wacc + f * RF_PER_LAYER;
    RfAct act;
    float o[RF_DIM];
    float xi[RF_DIM];
#pragma unroll
    for (int j = 0; j < RF_DIM; ++j) xi[j] = xin[f][j];
    rf_mlp(P, xi, o, &act);
    // y = x + net(x): grad flows to x directly (g) and through the net (go = g)
    // ---- last linear (3 x 8) and Swish 6
    float gh[RF_HID];
    {
      float a[RF_HID], dv[RF_HID], ds[RF_HID];
#pragma unroll
      for (int j = 0; j < RF_HID; ++j) {
        a[j] = rf_swish(act.pre[5][j], P[RF_S + 6]);
        rf_dswish(act.pre[5][j], P[RF_S + 6], dv[j], ds[j]);
      }
      float sacc = 0.f;
#pragma unroll
      for (int j = 0; j < RF_HID; ++j) {
        float ga = 0.f;
#pragma unroll
        for (int i = 0; i < RF_DIM; ++i) {
          ga = fmaf(P[RF_WL + i * RF_HID + j], g[i], ga);
          rf_acc(A, RF_WL + i * RF_HID + j, g[i] * a[j]);
        }
        gh[j] = ga * dv[j];
        sacc = fmaf(ga, ds[j], sacc);
      }
#pragma unroll
      for (int i = 0; i < RF_DIM; ++i) rf_acc(A, RF_BL + i, g[i]);
      rf_acc(A, RF_S + 6, sacc);
    }
    // ---- hidden linears 5..1 (8 x 8) and their Swish
#pragma unroll
    for (int k = 4; k >= 0; --k) {
      const float* W = P + RF_WH + k * (RF_HID * RF_HID + RF_HID);
      const int wo = RF_WH + k * (RF_HID * RF_HID + RF_HID);
      float a[RF_HID], dv[RF_HID], ds[RF_HID], gn[RF_HID];
#pragma unroll
      for (int j = 0; j < RF_HID; ++j) {
        a[j] = rf_swish(act.pre[k][j], P[RF_S + 1 + k]);
        rf_dswish(act.pre[k][j], P[RF_S + 1 + k], dv[j], ds[j]);
      }
      float sacc = 0.f;
#pragma unroll
      for (int j = 0; j < RF_HID; ++j) {
        float ga = 0.f;
#pragma unroll
        for (int i = 0; i < RF_HID; ++i) {
          ga = fmaf(W[i * RF_HID + j], gh[i], ga);
          rf_acc(A, wo + i * RF_HID + j, gh[i] * a[j]);
        }
        gn[j] = ga * dv[j];
        sacc = fmaf(ga, ds[j], sacc);
      }
#pragma unroll
      for (int i = 0; i < RF_HID; ++i) rf_acc(A, wo + RF_HID * RF_HID + i, gh[i]);
      rf_acc(A, RF_S + 1 + k, sacc);
#pragma unroll
      for (int j = 0; j < RF_HID; ++j) gh[j] = gn[j];
    }
    // ---- first linear (8 x 3) and Swish 0
    {
      float a[RF_DIM], dv[RF_DIM], ds[RF_DIM];
#pragma unroll
      for (int j = 0; j < RF_DIM; ++j) {
        a[j] = rf_swish(act.pre0[j], P[RF_S + 0]);
        rf_dswish(act.pre0[j], P[RF_S + 0], dv[j], ds[j]);
      }
      float sacc = 0.f;
      float gxn[RF_DIM];
#pragma unroll
      for (int j = 0; j < RF_DIM; ++j) {
        float ga = 0.f;
#pragma unroll
        for (int i = 0; i < RF_HID; ++i) {
          ga = fmaf(P[RF_W0 + i * RF_DIM + j], gh[i], ga);
          rf_acc(A, RF_W0 + i * RF_DIM + j, gh[i] * a[j]);
        }
        gxn[j] = ga * dv[j];
        sacc = fmaf(ga, ds[j], sacc);
      }
#pragma unroll
      for (int i = 0; i < RF_HID; ++i) rf_acc(A, RF_B0 + i, gh[i]);
      rf_acc(A, RF_S + 0, sacc);
#pragma unroll
      for (int j = 0; j < RF_DIM; ++j) g[j] += gxn[j];
    }
  }
  if (live) {
    gx[3 * (size_t)n] = g[0];
    gx[3 * (size_t)n + 1] = g[1];
    gx[3 * (size_t)n + 2] = g[2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < np; i += RF_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < RF_WARPS; ++w) t += sA[w * np + i];
    partial[(size_t)blockIdx.x * np + i] = t;
  }
}

// geff[i] = sum over the CTA partials, fixed order (gradient w.r.t. the EFFECTIVE parameters)
__global__ void resflow_reduce_kernel(const float* __restrict__ partial, int nblocks, int np, float* __restrict__ geff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  float t = 0.f;
  for (int b = 0; b < nblocks; ++b) t += partial[(size_t)b * np + i];
  geff[i] = t;
}

// Chain to the raw parameters (one CTA):  W_eff = W / f, f = max(1, sigma / c), sigma = u^T W v
//   dL/dW = G / f - [sigma > c] (<G, W> / (f^2 c)) u v^T ;   dL/dbeta = G_s * sigmoid(beta) (softplus') ;  biases unchanged.
__global__ void resflow_chain_kernel(const float* __restrict__ geff, const float* __restrict__ raw, const float* __restrict__ uv,
                                     int n_layers, float coeff, float* __restrict__ graw) {
  __shared__ float sF[RF_MAX_LAYERS * RF_LIN], sD[RF_MAX_LAYERS * RF_LIN];
  for (int t = threadIdx.x; t < n_layers * RF_LIN; t += blockDim.x) {
    const int f = t / RF_LIN, k = t % RF_LIN;
    int w_off, b_off, out, in, u_off, v_off;
    rf_lin_geom(k, w_off, b_off, out, in, u_off, v_off);
    const float sg = rf_sigma(raw, uv, f, k);
    const float fac = fmaxf(sg / coeff, 1.f);
    float dot = 0.f;
    for (int e = 0; e < out * in; ++e) dot = fmaf(geff[f * RF_PER_LAYER + w_off + e], __ldg(raw + f * RF_PER_LAYER + w_off + e), dot);
    sF[t] = fac;
    sD[t] = (sg / coeff > 1.f) ? dot / (fac * fac * coeff) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_layers * RF_PER_LAYER; i += blockDim.x) {
    const int f = i / RF_PER_LAYER, r = i % RF_PER_LAYER;
    const int k = rf_weight_layer(r);
    const float g = geff[i];
    float o;
    if (k >= 0) {
      int w_off, b_off, out, in, u_off, v_off;
      rf_lin_geom(k, w_off, b_off, out, in, u_off, v_off);
      const int e = r - w_off, row = e / in, col = e % in;
      o = g / sF[f * RF_LIN + k] - sD[f * RF_LIN + k] * __ldg(uv + f * RF_UV_PER_LAYER + u_off + row) * __ldg(uv + f * RF_UV_PER_LAYER + v_off + col);
    } else if (r >= RF_S) {
      const float b = __ldg(raw + i);
      o = g * (b > 20.f ? 1.f : 1.f / (1.f + expf(-b)));
    } else {
      o = g;
    }
    graw[i] = o;
  }
}

}  // namespace shwd

using namespace shwd;

extern "C" int shwd_resflow_params_per_layer(void) { return RF_PER_LAYER; }
extern "C" int shwd_resflow_uv_per_layer(void) { return RF_UV_PER_LAYER; }

extern "C" size_t shwd_resflow_workspace_bytes(int npts, int n_layers) {
  if (npts <= 0 || n_layers <= 0) return 0;
  const size_t blocks = ((size_t)npts + RF_THREADS - 1) / RF_THREADS;
  return (blocks + 1) * (size_t)n_layers * RF_PER_LAYER * sizeof(float);
}

extern "C" int shwd_resflow_fwd(const float* x, int npts, const float* params, const float* uv, int n_layers, float coeff,
                                float* y, void* stream) {
  if (!x || !params || !uv || !y || npts < 0 || n_layers <= 0 || n_layers > RF_MAX_LAYERS || !(coeff > 0.f))
    return SHWD_ERR_INVALID_ARGUMENT;
  if (npts == 0) return SHWD_OK;
  const int blocks = (npts + RF_THREADS - 1) / RF_THREADS;
  resflow_fwd_kernel<<<blocks, RF_THREADS, n_layers * RF_PER_LAYER * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      x, npts, params, uv, n_layers, coeff, y);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_resflow_bwd(const float* x, const float* gy, int npts, const float* params, const float* uv, int n_layers,
                                float coeff, float* gx, float* gparams, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !gy || !params || !uv || !gx || !gparams || npts < 0 || n_layers <= 0 || n_layers > RF_MAX_LAYERS || !(coeff > 0.f))
    return SHWD_ERR_INVALID_ARGUMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int np = n_layers * RF_PER_LAYER;
  if (npts == 0) {
    SHWD_CUDA_CHECK(cudaMemsetAsync(gparams, 0, np * sizeof(float), s));
    return SHWD_OK;
  }
  if (!workspace || workspace_bytes < shwd_resflow_workspace_bytes(npts, n_layers) ||
      (reinterpret_cast<uintptr_t>(workspace) & 3))
    return SHWD_ERR_WORKSPACE;
  const int blocks = (npts + RF_THREADS - 1) / RF_THREADS;
  float* partial = static_cast<float*>(workspace);
  float* geff = partial + (size_t)blocks * np;
  const size_t smem = (size_t)(1 + RF_WARPS) * np * sizeof(float);
  if (smem > 32 * 1024)  // static + dynamic beyond 48 KB needs the opt-in (static is < 16 KB here)
    SHWD_CUDA_CHECK(cudaFuncSetAttribute(resflow_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  resflow_bwd_kernel<<<blocks, RF_THREADS, smem, s>>>(x, gy, npts, params, uv, n_layers, coeff, gx, partial);
  SHWD_CUDA_CHECK(cudaGetLastError());
  resflow_reduce_kernel<<<(np + 127) / 128, 128, 0, s>>>(partial, blocks, np, geff);
  SHWD_CUDA_CHECK(cudaGetLastError());
  resflow_chain_kernel<<<1, 256, 0, s>>>(geff, params, uv, n_layers, coeff, gparams);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

// Measurement helpers: FP32-FMA and MUFU(ex2) issue-rate microbenchmarks.
//
// the OT sweeps are bound by the FP32 / SFU pipes, whose peaks are not in MEASURED_PEAKS.json, so
// bench.py measures them live with these kernels (CUDA events around the launch) and reports the roofline fraction
// of the Sinkhorn kernels against the MEASURED FP32 lane-op rate (an FFMA counts as one lane-op per lane).
#include "common.cuh"

namespace shwd {

constexpr int PK_THREADS = 512;
constexpr int PK_CHAINS = 8;
constexpr int PK_INNER = 64;

// 8 independent FFMA chains per thread; PK_INNER * 8 FFMA per outer iteration.
__global__ void __launch_bounds__(PK_THREADS) peak_fp32_kernel(float* out, int iters, float a, float b) {
  float v[PK_CHAINS];
#pragma unroll
  for (int c = 0; c < PK_CHAINS; ++c) v[c] = (float)(threadIdx.x + c) * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < PK_INNER; ++k) {
#pragma unroll
      for (int c = 0; c < PK_CHAINS; ++c) v[c] = fmaf(v[c], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < PK_CHAINS; ++c) s += v[c];
  out[(size_t)blockIdx.x * PK_THREADS + threadIdx.x] = s;
}

// 8 independent ex2.approx chains per thread.
__global__ void __launch_bounds__(PK_THREADS) peak_mufu_kernel(float* out, int iters) {
  float v[PK_CHAINS];
#pragma unroll
  for (int c = 0; c < PK_CHAINS; ++c) v[c] = (float)(threadIdx.x + c) * 1e-4f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < PK_INNER; ++k) {
#pragma unroll
      for (int c = 0; c < PK_CHAINS; ++c) v[c] = ex2_approx(v[c]);  // converges to the fixed point of 2^x - harmless
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < PK_CHAINS; ++c) s += v[c];
  out[(size_t)blockIdx.x * PK_THREADS + threadIdx.x] = s;
}

}  // namespace shwd

using namespace shwd;

static int peak_grid() { return sm_count() * 4; }  // 4 x 512 threads = 64 warps per SM

extern "C" int shwd_peak_fp32(float* out, int iters, double* ops, void* stream) {
  if (!out || iters <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  const int grid = peak_grid();
  peak_fp32_kernel<<<grid, PK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(out, iters, 0.999f, 1e-3f);
  SHWD_CUDA_CHECK(cudaGetLastError());
  if (ops) *ops = (double)grid * PK_THREADS * (double)iters * PK_INNER * PK_CHAINS;
  return SHWD_OK;
}

extern "C" int shwd_peak_mufu(float* out, int iters, double* ops, void* stream) {
  if (!out || iters <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  const int grid = peak_grid();
  peak_mufu_kernel<<<grid, PK_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(out, iters);
  SHWD_CUDA_CHECK(cudaGetLastError());
  if (ops) *ops = (double)grid * PK_THREADS * (double)iters * PK_INNER * PK_CHAINS;
  return SHWD_OK;
}

// Data side of a registration training step: the random rigid transform the reference applies to every
// source cloud on the CPU, one item at a time inside DataLoader workers,
//     Dataset_Transformation.__call__ / create_pose_7d / quaternion_rotate / qrot
//         Point_Cloud_Resistration/data_utils/Data_set_maker.py:40-52, 173-230
// as one HBM-bound launch over the whole batch resident on the device:
//     q_b = pose[b, 0:4] / max(|pose[b, 0:4]|, 1e-12)                       (create_pose_7d: F.normalize)
//     out[b, n] = v + 2 (w (q x v) + q x (q x v)) + t_b,  v = src[b, n]      (qrot + translation, same operation order)
//     R_b = columns q-rotate(e_k)  -> igt_rotation = quaternion_rotate(eye(3), igt).permute(1, 0)   (:224)
// 12 B read + 12 B written per point; the (optional) Gaussian sensor noise of add_noise (:13-22) is fused in with a
// counter-based generator (a 64-bit mix of (seed, element index) -> Box-Muller), so a noisy, transformed batch costs one
// pass.
#include "common.cuh"

namespace shwd {

__device__ __forceinline__ float3 rg_cross(float3 a, float3 b) {
  return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 rg_qrot(float w, float3 q, float3 v) {
  const float3 uv = rg_cross(q, v);
  const float3 uuv = rg_cross(q, uv);
  return make_float3(v.x + 2.f * (w * uv.x + uuv.x), v.y + 2.f * (w * uv.y + uuv.y), v.z + 2.f * (w * uv.z + uuv.z));
}
// splitmix64: a counter-based generator (independent streams per element, reproducible for a given seed)
__device__ __forceinline__ unsigned long long rg_mix(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float rg_normal(unsigned long long seed, unsigned long long idx) {
  const unsigned long long r = rg_mix(seed ^ rg_mix(idx));
  const float u1 = ((unsigned)(r >> 40) + 1u) * (1.f / 16777217.f);        // (0, 1)
  const float u2 = (unsigned)((r >> 8) & 0xFFFFFFu) * (1.f / 16777216.f);  // [0, 1)
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

__global__ void __launch_bounds__(256) rigid_transform_kernel(const float* __restrict__ src, const float* __restrict__ pose, int N,
                                                              float noise_std, unsigned long long seed, float* __restrict__ out,
                                                              float* __restrict__ R) {
  const int b = blockIdx.y;
  const float* P = pose + 7 * (size_t)b;
  float w = __ldg(P), qx = __ldg(P + 1), qy = __ldg(P + 2), qz = __ldg(P + 3);
  const float nr = fmaxf(sqrtf(w * w + qx * qx + qy * qy + qz * qz), 1e-12f);  // F.normalize
  w /= nr;
  qx /= nr;
  qy /= nr;
  qz /= nr;
  const float3 q = make_float3(qx, qy, qz);
  const float3 t = make_float3(__ldg(P + 4), __ldg(P + 5), __ldg(P + 6));
  if (R && blockIdx.x == 0 && threadIdx.x < 3) {
    const int k = threadIdx.x;  // column k of R = rotated e_k  (row k of quaternion_rotate(eye), transposed)
    const float3 e = rg_qrot(w, q, make_float3(k == 0, k == 1, k == 2));
    R[9 * (size_t)b + 0 * 3 + k] = e.x;
    R[9 * (size_t)b + 1 * 3 + k] = e.y;
    R[9 * (size_t)b + 2 * 3 + k] = e.z;
  }
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const size_t o = ((size_t)b * N + n) * 3;
  float3 v = make_float3(__ldg(src + o), __ldg(src + o + 1), __ldg(src + o + 2));
  if (noise_std > 0.f) {  // add_noise (:13-22) is applied to the source BEFORE the transform
    v.x += noise_std * rg_normal(seed, 3ull * ((unsigned long long)b * N + n));
    v.y += noise_std * rg_normal(seed, 3ull * ((unsigned long long)b * N + n) + 1);
    v.z += noise_std * rg_normal(seed, 3ull * ((unsigned long long)b * N + n) + 2);
  }
  const float3 r = rg_qrot(w, q, v);
  out[o] = r.x + t.x;
  out[o + 1] = r.y + t.y;
  out[o + 2] = r.z + t.z;
}

}  // namespace shwd

using namespace shwd;

extern "C" int shwd_rigid_transform(const float* src, const float* pose7, int B, int N, float noise_std, unsigned long long seed,
                                    float* out, float* rotation, void* stream) {
  if (!src || !pose7 || !out || B < 0 || N <= 0 || noise_std < 0.f) return SHWD_ERR_INVALID_ARGUMENT;
  if (B == 0) return SHWD_OK;
  if (B > 65535) return SHWD_ERR_UNSUPPORTED;
  dim3 grid((N + 255) / 256, B);
  rigid_transform_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, pose7, N, noise_std, seed, out, rotation);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

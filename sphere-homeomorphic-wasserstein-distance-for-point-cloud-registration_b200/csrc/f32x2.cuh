// Packed 2 x float32 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2 via PTX fma.rn.f32x2 & co).
//
// One packed instruction performs two independent IEEE float32 operations -- bit-identical to the scalar FFMA/FADD/FMUL
// -- at the same lane throughput but HALF the issue slots (tools/microbench/ffma2_bench.cu: 37.05 vs 36.25 Tlane-FMA/s
// pure, 32.5 vs 30.2 T/s when mixed with the OT sweeps' MUFU ratio).  The OT sweeps are issue-bound with scalar code
// (ncu: issue slots busy 65 %, 24.5 inst per element), so packing moves them to the FMA-pipe / XU-pipe co-limit.
// ptxas folds the neg/abs helpers below into operand modifiers (FADD2 R, -|R|.F32x2, 1) and broadcasts scalars
// (R.F32), so there are no extra MOVs.
#pragma once
#include "common.cuh"

namespace shwd {

struct f2 {
  unsigned long long v;
};

__device__ __forceinline__ f2 mk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 bc2(float a) { return mk2(a, a); }
__device__ __forceinline__ float lo2(f2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  return x;
}
__device__ __forceinline__ float hi2(f2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  return y;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 neg2(f2 a) { return mk2(-lo2(a), -hi2(a)); }
__device__ __forceinline__ f2 abs2(f2 a) { return mk2(fabsf(lo2(a)), fabsf(hi2(a))); }
__device__ __forceinline__ f2 max2(f2 a, f2 b) { return mk2(fmaxf(lo2(a), lo2(b)), fmaxf(hi2(a), hi2(b))); }
__device__ __forceinline__ f2 ex2_2(f2 a) { return mk2(ex2_approx(lo2(a)), ex2_approx(hi2(a))); }
__device__ __forceinline__ f2 f2_from(float2 v) { return mk2(v.x, v.y); }

}  // namespace shwd

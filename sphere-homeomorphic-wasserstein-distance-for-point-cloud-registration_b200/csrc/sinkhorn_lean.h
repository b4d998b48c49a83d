// Interface between sinkhorn.cu (C ABI entry points, path selection) and sinkhorn_lean.cu (the dedicated-CTA kernels
// for small problems).  Include after sinkhorn_core.cuh.
#pragma once
#include "sinkhorn_core.cuh"

namespace shwd {

struct LeanGeom {
  int Q;            // CTAs per pair
  int grid;         // CTAs of the launch
  int whole_pairs;  // B exceeds the SM count: every CTA takes whole pairs in turn (Q = 1)
  int rw_shift;     // log2(owners per warp row): 3, 4 or 5
  int R[2];         // owners per CTA band (row / col owners), multiples of the warp row
  int T[2];         // packed records of the streamed cloud per half-step type (type 0 streams y, type 1 streams x)
  int nc[2];        // cached intermediates: record pairs per lane and type held in shared memory (sinkhorn_lean.cu)
  int cr[2];        // ... owner groups per warp the table is laid out for (1 or 2)
  int smem;         // dynamic shared memory of the launch
};

bool lean_plan(int B, int N, int M, LeanGeom* out);
bool lean_selected(int B, int N, int M, int fast, int hist_levels, float thresh);
int launch_lean_fwd(int fast, const SinkParams& prm, const LeanGeom& gm, cudaStream_t s);
int launch_lean_bwd(int fast, const SinkParams& prm, const LeanGeom& gm, cudaStream_t s);

}  // namespace shwd

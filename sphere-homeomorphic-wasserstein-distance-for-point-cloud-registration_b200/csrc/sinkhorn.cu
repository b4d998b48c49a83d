// Entropic OT (log-domain Sinkhorn) on an on-the-fly cost: forward, and reverse-mode through every unrolled
// iteration.  The N x M cost / kernel / plan matrices are never materialised.
//
// Replaces   cost matrices          Point_Cloud_Resistration/losses/s2_wasserstein.py:52-63, 112-123
//            Sinkhorn recurrence    Comparison_Wasserstein_with_Chamfer_distance/losses/sinkhorn.py:24-58
//                                   Point_Cloud_Resistration/losses/Sinkhorn.py:25-60, Sinkhorn_fixed.py:32-67
//            its autograd           (torch reverse-mode over the unrolled loop)
//
// Formulation (scaled log2 domain, k = log2(e)/eps, alpha = k u, beta = k v, la2 = log2(1/N + 1e-8), lb2 likewise):
//     alpha^l_i = la2 - log2 sum_j 2^(beta^{l-1}_j - k C_ij)          ("row" half-step, owners = x points)
//     beta^l_j  = lb2 - log2 sum_i 2^(alpha^l_i   - k C_ij)           ("col" half-step, owners = y points)
// which is the reference recurrence  u <- eps (log a - LSE_j M) + u  with the +u/-u pair cancelled analytically.
// cost = sum_ij P_ij C_ij with P_ij = 2^(alpha^L_i + beta^L_j - k C_ij).
//
// Execution model: ONE persistent cooperative launch per direction.  Every half-step is a "sweep": each owner point
// streams all points of the other cloud (staged in shared memory as float4 (x,y,z,potential) records, broadcast
// LDS.128) and recomputes dot -> acos -> exp2 in registers on the FP32/SFU pipes (K = 3: tensor cores are useless here).
// Owners are dealt to the CTAs in 32-owner groups over the flattened (pair, group) space, so the 148 SMs stay
// balanced for any (B, N).  Inside a CTA the 16 warps split the streamed range; partial results are merged through
// shared memory in a fixed order (bit-reproducible, no float atomics on the data path).  A half-step of pair b may
// start once every group of its previous half-step is done: per-pair counters in global memory
// (release: __syncthreads + __threadfence + atomicAdd; acquire: ld.acquire.gpu spin + __syncthreads).  Pairs never
// wait on each other.  Iterates are kept in a write-once history (B, L+1, N): no WAR hazards, and the backward
// replays them in reverse.
//
// Backward sweeps (derivation checked against torch autograd to 3e-8 in f64): with S^u,l_ij = 2^(alpha^l_i +
// beta^{l-1}_j - kC_ij - la2) and S^v,l_ij = 2^(alpha^l_i + beta^l_j - kC_ij - lb2),
//     row sweep(l):  abar^l_i      = [l=L: g ln2 r_i] - sum_j bbar^l_j S^v,l_ij
//     col sweep(l):  bbar^{l-1}_j  =                  - sum_i abar^l_i S^u,l_ij
// and each sweep also adds the cost-gradient terms it can form for its OWNER side without any cross-thread reduction:
// its own term plus the adjacent other-type term that shares the streamed vector (row sweep(l): S^u,l+1; col sweep(l):
// S^v,l), i.e. two exp2 per element instead of a 4-value cross-lane reduction per element.
#include "sinkhorn_core.cuh"
#include "sinkhorn_lean.h"

namespace shwd {

// ================================================================================================================
// Forward: 2L half-steps, then [early stop: pick L*], then the two final sweeps (row/col sums of P*C) and the cost.
// ================================================================================================================
template <int FAST>
__global__ void __launch_bounds__(SK_THREADS, SK_CTAS_PER_SM) sinkhorn_fwd_kernel(const SinkParams prm) {
  extern __shared__ float4 smem4[];
  float4* sS = smem4;                      // 2 x CHUNK_PAD staged records (one set per segment)
  float4* part = smem4 + 2 * CHUNK_PAD;
  float2* sAdj = reinterpret_cast<float2*>(part + SK_WARPS * GMAX * 32);  // 2 x CHUNK_PAD
  float4* sOwn = reinterpret_cast<float4*>(sAdj + 2 * CHUNK_PAD);          // 3 x GMAX*32 staged owner records
  float4* sOwnC = sOwn + 3 * GMAX * 32;                                    // 2 x GMAX*32 resident owner coordinates
  float* sOldP = reinterpret_cast<float*>(sOwnC + 2 * GMAX * 32);          // 2 x GMAX*32 resident owner potentials
  __shared__ int s_ls;
  __shared__ SweepIO s_ios[2];      // item descriptors: shared, not local memory (every __threadfence of the signal
  __shared__ ResidentType s_RT[2];  // path invalidates L1, and spilled descriptors would be re-fetched from L2 each time)

  const int gr = (prm.N + 31) / 32, gc = (prm.M + 31) / 32;  // groups per pair, row / col owners
  PROF_INIT();
  const int HL = prm.hist_levels;
  const int L = prm.iters;
  auto slot = [&](int l) { return HL > 1 ? l : 0; };

  // beta^0 = 0 (kept in the history so the backward can stream it)
  if (HL > 1) {
    for (long long i = (long long)blockIdx.x * SK_THREADS + threadIdx.x; i < (long long)prm.B * prm.M;
         i += (long long)gridDim.x * SK_THREADS) {
      int b = (int)(i / prm.M), j = (int)(i % prm.M);
      prm.beta[((size_t)b * HL) * prm.M + j] = 0.f;
    }
  }

  ResidentType (&RT)[2] = s_RT;
  resident_setup<FAST>(prm, gr, gc, sS, sOwnC, sOldP, RT);

  // data-flow synchronisation of the LSE half-steps (see SPIN_SENTINEL); the counters then only see the last two of them
  const bool spin = SHWD_SPIN && is_packed_cost(FAST) && HL > 1 && prm.spin_ready;
  const int lse_done = spin ? (gr + gc) : L * (gr + gc);  // per-pair counter value once every LSE half-step is published

  for (int h = 0; h < 2 * L; ++h) {
    const int type = h & 1;            // 0: alpha (row owners), 1: beta (col owners)
    const int l = (h >> 1) + 1;        // level being produced
    const int gpp = type ? gc : gr;
    int gbeg, gend;
    cta_range((long long)prm.B * gpp, gbeg, gend);
    const int target_unit_r = ((h + 1) >> 1), target_unit_c = (h >> 1);  // #row / #col half-steps before h
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
      io.lconst = type ? prm.lb2 : prm.la2;
      io.err_out = nullptr;
      io.old_pot = nullptr;
      if (type == 0) {
        io.own = prm.X + (size_t)b * prm.N;
        io.n_own = prm.N;
        io.str = prm.Y + (size_t)b * prm.M;
        io.n_str = prm.M;
        io.str_pot = (l == 1) ? nullptr : prm.beta + ((size_t)b * HL + slot(l - 1)) * prm.M;
        io.out_pot = prm.alpha + ((size_t)b * HL + slot(l)) * prm.N;
        io.out_pot_lo = prm.alpha_lo + ((size_t)b * HL + slot(l)) * prm.N;
        io.old_pot = (l == 1) ? nullptr : prm.alpha + ((size_t)b * HL + slot(l - 1)) * prm.N;
        if (prm.thresh > 0.f) io.err_out = prm.err + (size_t)(l - 1) * prm.B + b;
      } else {
        io.own = prm.Y + (size_t)b * prm.M;
        io.n_own = prm.M;
        io.str = prm.X + (size_t)b * prm.N;
        io.n_str = prm.N;
        io.str_pot = prm.alpha + ((size_t)b * HL + slot(l)) * prm.N;
        io.old_pot = (l == 1) ? nullptr : prm.beta + ((size_t)b * HL + slot(l - 1)) * prm.M;  // beta^0 = 0
        io.out_pot = prm.beta + ((size_t)b * HL + slot(l)) * prm.M;
        io.out_pot_lo = prm.beta_lo + ((size_t)b * HL + slot(l)) * prm.M;
      }
      }
      ++nseg;
      }
      __syncthreads();  // the item descriptors (written by thread 0) are visible to the CTA
      PROF_MARK(5);
        // (with the early-stop rule the running-maximum LSE is kept: its result does not depend on the previous iterate, so
      //  the iteration settles on a bitwise fixed point exactly where the reference's does -- sinkhorn.py:42-44)
      WaitSpec ws = {prm.done, segb[0], segb[1], target_unit_r * gr + target_unit_c * gc, prm.status,
                     (SHWD_OFFSET_LSE && h >= 2 && is_packed_cost(FAST) && !(prm.thresh > 0.f)) ? 1 : 0, spin ? 1 : 0};
      sweep<FAST, MODE_LSE, false>(prm.cp, ios, seg0, seg1, nseg, sS, sAdj, part, sOwn, ws, RT[type].ok ? &RT[type] : nullptr);
      // spin mode: only the last alpha / beta half-steps publish through the counters (for the final sweeps and the cost)
      if (!spin || h >= 2 * L - 2) signal_done2(prm.done, segb, seg0, seg1, nseg);
      PROF_MARK(4);
    }
  }

  fwd_tail<FAST>(prm, lse_done, sS, sAdj, part, sOwn, s_ios, s_ls);
}

// ================================================================================================================
// Backward: 2L*+1 reverse sweeps.  phase ph < 2L*: l = L* - ph/2, type = ph & 1;  phase 2L*: row sweep(0).
// ================================================================================================================
template <int FAST>
__global__ void __launch_bounds__(SK_THREADS, SK_CTAS_PER_SM) sinkhorn_bwd_kernel(const SinkParams prm) {
  extern __shared__ float4 smem4[];
  float4* sS = smem4;                      // 2 x CHUNK_PAD staged records (one set per segment)
  float4* part = smem4 + 2 * CHUNK_PAD;
  float2* sAdj = reinterpret_cast<float2*>(part + SK_WARPS * GMAX * 32);  // 2 x CHUNK_PAD
  float4* sOwn = reinterpret_cast<float4*>(sAdj + 2 * CHUNK_PAD);          // 3 x GMAX*32 staged owner records
  float4* sOwnC = sOwn + 3 * GMAX * 32;                                    // 2 x GMAX*32 resident owner coordinates
  float* sOldP = reinterpret_cast<float*>(sOwnC + 2 * GMAX * 32);          // (unused by the backward)
  __shared__ SweepIO s_ios[2];
  __shared__ ResidentType s_RT[2];
  ResidentType (&RT)[2] = s_RT;
  if (threadIdx.x == 0) RT[0].ok = RT[1].ok = 0;

  const int gr = (prm.N + 31) / 32, gc = (prm.M + 31) / 32;
  const int Ls = *prm.iters_run;
  PROF_INIT();

  const bool spin = false;  // two parity planes + counters (see sinkhorn_core.cuh)
  for (int ph = 0; ph <= 2 * Ls; ++ph) {
    if (ph == 2) resident_setup<FAST>(prm, gr, gc, sS, sOwnC, sOldP, RT);  // the two FINAL sweeps (ph 0, 1) stage through sS
    bwd_phase<FAST>(prm, ph, Ls, spin, sS, sAdj, part, sOwn, s_ios, RT);
  }
}

// A persistent launch whose inter-CTA wait timed out sets the workspace's status word and carries on with garbage.  Every
// launcher ends with this kernel: one word read when all is well, NaN over the launch's outputs otherwise -- decided on
// the device, so a failed launch can never pass for a result (and the NaN propagates through shwd_sphere_map_bwd).
__global__ void poison_on_failure_kernel(const int* status, float* a, size_t na, float* b, size_t nb) {
  if (*status == 0) return;
  const float nan = __int_as_float(0x7fc00000);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += (size_t)gridDim.x * blockDim.x) a[i] = nan;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += (size_t)gridDim.x * blockDim.x) b[i] = nan;
}

// Dense plan / cost for the reference's (cost, P, C) return value (opt-in, small problems only).
template <int FAST>
__global__ void plan_dense_kernel(const float4* X, const float4* Y, int N, int M, CostParams cp, const float* alpha,
                                  const float* beta, int sn, int sm, float inv_k, float* P, float* C) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= M) return;
  float4 x = X[(size_t)b * N + i], y = Y[(size_t)b * M + j];
  typename Cost<FAST>::E ce = Cost<FAST>::eval(cp, x.x, x.y, x.z, y.x, y.y, y.z);
  size_t o = ((size_t)b * N + i) * M + j;
  if (C) C[o] = Cost<FAST>::kc(cp, ce) * inv_k;
  // the plan exactly as the final-cost sweeps evaluate it: 2^(fl(M(beta_j) + alpha_i))
  if (P) P[o] = ex2_approx(__fadd_rn(Cost<FAST>::m(cp, ce, beta[(size_t)b * sm + j]), alpha[(size_t)b * sn + i]));
}

}  // namespace shwd

using namespace shwd;

extern "C" size_t shwd_sinkhorn_workspace_bytes(int B, int N, int M, int iters) {
  if (B <= 0 || N <= 0 || M <= 0 || iters <= 0) return 0;
  // (the lean backward polls level-indexed write-once adjoint planes; reserve them wherever that path may be taken)
  return carve(nullptr, B, N, M, iters, lean_plan(B, N, M, nullptr) ? iters + 1 : 2).total;
}

extern "C" int shwd_sinkhorn_status_offset(void) { return 0; }

static int check_common(const void* x4, const void* y4, int B, int N, int M, float p, float eps, int iters) {
  if (!x4 || !y4 || B <= 0 || N <= 0 || M <= 0 || iters <= 0) return SHWD_ERR_INVALID_ARGUMENT;
  if (!(eps > 0.f) || !(p > 0.f)) return SHWD_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(x4) & 15) || (reinterpret_cast<uintptr_t>(y4) & 15)) return SHWD_ERR_INVALID_ARGUMENT;
  return SHWD_OK;
}

extern "C" int shwd_sinkhorn_fwd(const float* x4, const float* y4, int B, int N, int M, int cost_kind, float p, float n_power,
                                 float eps, int iters, float early_stop_thresh, int hist_levels, float* alpha_hist,
                                 float* beta_hist, float* row_pc, float* col_pc, float* cost, int* iters_run,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x4, y4, B, N, M, p, eps, iters);
  if (rc) return rc;
  if (cost_kind < 0 || cost_kind > 3 || !alpha_hist || !beta_hist || !row_pc || !col_pc || !cost || !iters_run)
    return SHWD_ERR_INVALID_ARGUMENT;
  if (hist_levels != 1 && hist_levels != iters + 1) return SHWD_ERR_INVALID_ARGUMENT;
  if (early_stop_thresh > 0.f && hist_levels == 1) return SHWD_ERR_INVALID_ARGUMENT;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return SHWD_ERR_WORKSPACE;
  Workspace w = carve(workspace, B, N, M, iters, 2);
  if (workspace_bytes < w.total) return SHWD_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SHWD_CUDA_CHECK(cudaMemsetAsync(workspace, 0, w.head_bytes, s));

  const int fast = pick_fast(cost_kind, p, n_power);
  SinkParams prm = {};
  prm.X = reinterpret_cast<const float4*>(x4);
  prm.Y = reinterpret_cast<const float4*>(y4);
  prm.B = B;
  prm.N = N;
  prm.M = M;
  prm.cp = make_cost(cost_kind, p, n_power, eps, fast);
  prm.iters = iters;
  prm.hist_levels = hist_levels;
  prm.adj_planes = 2;
  prm.alpha = alpha_hist;
  prm.beta = beta_hist;
  prm.alpha_lo = alpha_hist + (size_t)B * hist_levels * N;
  prm.beta_lo = beta_hist + (size_t)B * hist_levels * M;
  prm.row_pc = row_pc;
  prm.col_pc = col_pc;
  prm.cost = cost;
  prm.iters_run = iters_run;
  prm.thresh = early_stop_thresh;
  fill_marginals(prm, N, M, eps);
  prm.done = w.done;
  prm.status = w.status;
  prm.err = w.err;
  prm.spin_ready = 0;
  if (SHWD_SPIN && is_packed_cost(fast) && hist_levels > 1) {
    // write-once potential planes (not the residual planes): sentinel-filled, see SPIN_SENTINEL
    SHWD_CUDA_CHECK(cudaMemsetAsync(alpha_hist, 0xFF, sizeof(float) * (size_t)B * hist_levels * N, s));
    SHWD_CUDA_CHECK(cudaMemsetAsync(beta_hist, 0xFF, sizeof(float) * (size_t)B * hist_levels * M, s));
    prm.spin_ready = 1;
  }
  LeanGeom gm;
  const size_t smem = sinkhorn_smem_bytes();
  const int maxg = B * (((N > M ? N : M) + 31) / 32);
  if (prm.spin_ready && lean_selected(B, N, M, fast, hist_levels, early_stop_thresh) && lean_plan(B, N, M, &gm)) {
    rc = launch_lean_fwd(fast, prm, gm, s);
  } else {
    switch (fast) {
      case FAST_GEO2: rc = launch_persistent(sinkhorn_fwd_kernel<FAST_GEO2>, prm, smem, maxg, s); break;
      case FAST_SQE2: rc = launch_persistent(sinkhorn_fwd_kernel<FAST_SQE2>, prm, smem, maxg, s); break;
      case FAST_GEO1: rc = launch_persistent(sinkhorn_fwd_kernel<FAST_GEO1>, prm, smem, maxg, s); break;
      case FAST_SQE1: rc = launch_persistent(sinkhorn_fwd_kernel<FAST_SQE1>, prm, smem, maxg, s); break;
      case FAST_EUC2: rc = launch_persistent(sinkhorn_fwd_kernel<FAST_EUC2>, prm, smem, maxg, s); break;
      case FAST_OMC2: rc = launch_persistent(sinkhorn_fwd_kernel<FAST_OMC2>, prm, smem, maxg, s); break;
      default: rc = launch_persistent(sinkhorn_fwd_kernel<GENERIC>, prm, smem, maxg, s);
    }
  }
  if (rc) return rc;
  poison_on_failure_kernel<<<1, 256, 0, s>>>(w.status, cost, (size_t)B, nullptr, 0);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_sinkhorn_bwd(const float* x4, const float* y4, int B, int N, int M, int cost_kind, float p, float n_power,
                                 float eps, int iters, const float* alpha_hist, const float* beta_hist, const float* row_pc,
                                 const float* col_pc, const int* iters_run, const float* grad_cost, float* g4x, float* g4y,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x4, y4, B, N, M, p, eps, iters);
  if (rc) return rc;
  if (cost_kind < 0 || cost_kind > 3 || !alpha_hist || !beta_hist || !row_pc || !col_pc || !iters_run || !grad_cost || !g4x ||
      !g4y)
    return SHWD_ERR_INVALID_ARGUMENT;
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) return SHWD_ERR_WORKSPACE;
  const int fast = pick_fast(cost_kind, p, n_power);
  LeanGeom gm;
  bool lean = lean_selected(B, N, M, fast, iters + 1, 0.f) && lean_plan(B, N, M, &gm);
  Workspace w = carve(workspace, B, N, M, iters, lean ? iters + 1 : 2);
  if (lean && workspace_bytes < w.total) {  // sized for the general kernel only: take that one
    lean = false;
    w = carve(workspace, B, N, M, iters, 2);
  }
  if (workspace_bytes < w.total) return SHWD_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SHWD_CUDA_CHECK(cudaMemsetAsync(workspace, 0, w.head_bytes, s));

  SinkParams prm = {};
  prm.X = reinterpret_cast<const float4*>(x4);
  prm.Y = reinterpret_cast<const float4*>(y4);
  prm.B = B;
  prm.N = N;
  prm.M = M;
  prm.cp = make_cost(cost_kind, p, n_power, eps, fast);
  prm.iters = iters;
  prm.hist_levels = iters + 1;
  prm.alpha = const_cast<float*>(alpha_hist);
  prm.beta = const_cast<float*>(beta_hist);
  prm.alpha_lo = prm.alpha + (size_t)B * (iters + 1) * N;
  prm.beta_lo = prm.beta + (size_t)B * (iters + 1) * M;
  prm.row_pc = const_cast<float*>(row_pc);
  prm.col_pc = const_cast<float*>(col_pc);
  prm.iters_run = const_cast<int*>(iters_run);
  prm.grad_cost = grad_cost;
  prm.g4x = reinterpret_cast<float4*>(g4x);
  prm.g4y = reinterpret_cast<float4*>(g4y);
  fill_marginals(prm, N, M, eps);
  prm.done = w.done;
  prm.status = w.status;
  prm.err = w.err;
  prm.abar = w.abar;
  prm.bbar = w.bbar;
  prm.spin_ready = 0;
  prm.adj_planes = lean ? iters + 1 : 2;
  if (lean) {
    // write-once adjoint planes, sentinel-filled (see SPIN_SENTINEL)
    SHWD_CUDA_CHECK(cudaMemsetAsync(w.abar, 0xFF, sizeof(float) * (size_t)(iters + 1) * B * N, s));
    SHWD_CUDA_CHECK(cudaMemsetAsync(w.bbar, 0xFF, sizeof(float) * (size_t)(iters + 1) * B * M, s));
    prm.spin_ready = 1;
    rc = launch_lean_bwd(fast, prm, gm, s);
  } else {
    const size_t smem = sinkhorn_smem_bytes();
    const int maxg = B * (((N > M ? N : M) + 31) / 32);
    switch (fast) {
      case FAST_GEO2: rc = launch_persistent(sinkhorn_bwd_kernel<FAST_GEO2>, prm, smem, maxg, s); break;
      case FAST_SQE2: rc = launch_persistent(sinkhorn_bwd_kernel<FAST_SQE2>, prm, smem, maxg, s); break;
      case FAST_GEO1: rc = launch_persistent(sinkhorn_bwd_kernel<FAST_GEO1>, prm, smem, maxg, s); break;
      case FAST_SQE1: rc = launch_persistent(sinkhorn_bwd_kernel<FAST_SQE1>, prm, smem, maxg, s); break;
      case FAST_EUC2: rc = launch_persistent(sinkhorn_bwd_kernel<FAST_EUC2>, prm, smem, maxg, s); break;
      case FAST_OMC2: rc = launch_persistent(sinkhorn_bwd_kernel<FAST_OMC2>, prm, smem, maxg, s); break;
      default: rc = launch_persistent(sinkhorn_bwd_kernel<GENERIC>, prm, smem, maxg, s);
    }
  }
  if (rc) return rc;
  poison_on_failure_kernel<<<64, 256, 0, s>>>(w.status, g4x, (size_t)B * N * 4, g4y, (size_t)B * M * 4);
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

extern "C" int shwd_sinkhorn_plan_dense(const float* x4, const float* y4, int B, int N, int M, int cost_kind, float p,
                                        float n_power, float eps, const float* alpha, const float* beta, int level_stride_n,
                                        int level_stride_m, float* P, float* C, void* stream) {
  if (!x4 || !y4 || B <= 0 || N <= 0 || M <= 0 || !(eps > 0.f)) return SHWD_ERR_INVALID_ARGUMENT;
  if (P && (!alpha || !beta)) return SHWD_ERR_INVALID_ARGUMENT;
  if (N > 65535 || B > 65535) return SHWD_ERR_UNSUPPORTED;
  const int fast = pick_fast(cost_kind, p, n_power);
  CostParams cp = make_cost(cost_kind, p, n_power, eps, fast);
  const float inv_k = (float)((double)eps / 1.4426950408889634);
  dim3 grid((M + 127) / 128, N, B), block(128);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float4* X = reinterpret_cast<const float4*>(x4);
  const float4* Y = reinterpret_cast<const float4*>(y4);
  switch (fast) {
    case FAST_GEO2:
      plan_dense_kernel<FAST_GEO2><<<grid, block, 0, s>>>(X, Y, N, M, cp, alpha, beta, level_stride_n, level_stride_m, inv_k, P, C);
      break;
    case FAST_SQE2:
      plan_dense_kernel<FAST_SQE2><<<grid, block, 0, s>>>(X, Y, N, M, cp, alpha, beta, level_stride_n, level_stride_m, inv_k, P, C);
      break;
    case FAST_GEO1:
      plan_dense_kernel<FAST_GEO1><<<grid, block, 0, s>>>(X, Y, N, M, cp, alpha, beta, level_stride_n, level_stride_m, inv_k, P, C);
      break;
    case FAST_SQE1:
      plan_dense_kernel<FAST_SQE1><<<grid, block, 0, s>>>(X, Y, N, M, cp, alpha, beta, level_stride_n, level_stride_m, inv_k, P, C);
      break;
    case FAST_EUC2:
      plan_dense_kernel<FAST_EUC2><<<grid, block, 0, s>>>(X, Y, N, M, cp, alpha, beta, level_stride_n, level_stride_m, inv_k, P, C);
      break;
    case FAST_OMC2:
      plan_dense_kernel<FAST_OMC2><<<grid, block, 0, s>>>(X, Y, N, M, cp, alpha, beta, level_stride_n, level_stride_m, inv_k, P, C);
      break;
    default:
      plan_dense_kernel<GENERIC><<<grid, block, 0, s>>>(X, Y, N, M, cp, alpha, beta, level_stride_n, level_stride_m, inv_k, P, C);
  }
  SHWD_CUDA_CHECK(cudaGetLastError());
  return SHWD_OK;
}

#ifdef SHWD_PROFILE
extern "C" int shwd_prof_read(unsigned long long* out8, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out8, shwd::g_prof, sizeof(unsigned long long) * 8);
  if (e != cudaSuccess) return SHWD_ERR_CUDA;
  if (reset) {
    unsigned long long z[8] = {0};
    e = cudaMemcpyToSymbol(shwd::g_prof, z, sizeof(z));
    if (e != cudaSuccess) return SHWD_ERR_CUDA;
  }
  return SHWD_OK;
}
#endif

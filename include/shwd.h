/*
 * shwd.h -- C ABI of libshwd_b200.so: the sphere-homeomorphic Wasserstein loss path on B200 (sm_100a).
 *
 * The reference (Satoshi0728/Sphere-Homeomorphic-Wasserstein-Distance-for-Point-Cloud-Registration) is pure
 * Python/torch and has no FFI; its boundary for this path is the Python call signature of the loss objects
 * (Point_Cloud_Resistration/losses/__init__.py:27-32, Comparison_Wasserstein_with_Chamfer_distance/losses/__init__.py:1-2).
 * Each entry point below names the reference code it replaces.  The Python mirror of the reference interface lives in
 * the package's losses/ directory and binds these symbols with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors), float32 unless stated;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation, no hidden allocation:
 *     scratch is passed in and sized by the *_workspace_bytes() helpers;
 *   - return value: 0 on success, negative code otherwise (shwd_error_string); no exceptions cross the ABI;
 *   - "packed points" are float4 records (x, y, z, w) -- 16-byte aligned -- produced by shwd_sphere_map_fwd.
 */
#ifndef SHWD_H_
#define SHWD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHWD_OK 0
#define SHWD_ERR_INVALID_ARGUMENT (-1)
#define SHWD_ERR_CUDA (-2)
#define SHWD_ERR_WORKSPACE (-3)
#define SHWD_ERR_UNSUPPORTED (-4)

/* cost kinds */
#define SHWD_COST_GEODESIC 0      /* acos(<x^,y^>)^p            s2_wasserstein.py:112-123                       */
#define SHWD_COST_SQEUCLID 1      /* sum_k |x_k-y_k|^p          s2_wasserstein.py:52-63, Sinkhorn.py:72-82       */
#define SHWD_COST_EUCLID 2        /* (sum_k |x_k-y_k|^p)^(1/p)  Sinkhorn_fixed.py:79-89                          */
#define SHWD_COST_ONE_MINUS_COS 3 /* (1-<x^,y^>)^p              max_spherical_w_cos_with_regulation.py:745       */

/* sphere-map flags */
#define SHWD_MAP_CENTER 1    /* x - mean_n(x)            train_W_COS.py:167-168                                   */
#define SHWD_MAP_NORMALIZE 2 /* x / max(||x||, 1e-8)     inside F.cosine_similarity, s2_wasserstein.py:122         */

int shwd_version(void);
const char* shwd_error_string(int code);
/* number of SMs / max co-resident CTAs the persistent kernels will use on the current device */
int shwd_device_sm_count(void);

/* ---- sphere map ------------------------------------------------------------------------------------------------
 * x (B,N,3) contiguous -> xh4 (B,N,4): (x^_0, x^_1, x^_2, 1/max(||x_c||,1e-8)) [w = 1 when not normalising].
 * Replaces train_W_COS.py:167-168 (centring) + the normalisation F.cosine_similarity applies at s2_wasserstein.py:122.
 * reg_out (nullable, (B) floats): sum_n | ||x_bn|| - 1 |   (regularization_of_normalizing_flow, s2_wasserstein.py:224-232),
 * computed on the un-centred, un-normalised input. */
int shwd_sphere_map_fwd(const float* x, float* xh4, float* reg_out, int B, int N, int flags, void* stream);
/* g4 (B,N,4) gradient w.r.t. the packed mapped points (w ignored) -> gx (B,N,3) gradient w.r.t. the raw points.
 * x is the raw input of the forward; xh4 its output.  greg (nullable, (B)): upstream gradient of reg_out. */
int shwd_sphere_map_bwd(const float* x, const float* xh4, const float* g4, const float* greg, float* gx, int B, int N,
                        int flags, void* stream);

/* ---- entropic OT on an on-the-fly cost (never materialises N x M) -----------------------------------------------
 * Replaces the cost matrices (s2_wasserstein.py:52-63,112-123) + the log-domain Sinkhorn recurrence
 * (Comparison_.../losses/sinkhorn.py:24-58; losses/Sinkhorn.py:25-60; losses/Sinkhorn_fixed.py:32-67).
 *   x4 (B,N,4), y4 (B,M,4) packed points (unit vectors for the two cosine kinds, raw coordinates otherwise)
 *   p            cost power; n_power: log_N_Sinkhorn's C^N (1 otherwise)
 *   eps, iters   entropic regularisation and iteration count L
 *   early_stop_thresh  <=0: run exactly L iterations (Sinkhorn.py); >0: use the iterate at which the batch-mean L1
 *                change of u first drops below it (sinkhorn.py:42-44) -- decided on the device, no host sync
 *   hist_levels  L+1 to keep every iterate (needed by the backward and by early stop) or 1 for forward-only
 *   alpha_hist (2,B,hist_levels,N), beta_hist (2,B,hist_levels,M): plane 0 = log2-domain scaled duals k*u, k*v
 *                (k = log2(e)/eps) of every iterate; plane 1 = the float32 rounding residual of each stored value
 *                (the backward normalises its softmax factors with the unrounded log-sum-exp)
 *   row_pc (B,N) = sum_j P_ij C_ij, col_pc (B,M) = sum_i P_ij C_ij, cost (B) = sum_ij P_ij C_ij
 *   iters_run (1 int, device): the iterate index L* actually used
 */
size_t shwd_sinkhorn_workspace_bytes(int B, int N, int M, int iters);
int shwd_sinkhorn_fwd(const float* x4, const float* y4, int B, int N, int M, int cost_kind, float p, float n_power,
                      float eps, int iters, float early_stop_thresh, int hist_levels, float* alpha_hist,
                      float* beta_hist, float* row_pc, float* col_pc, float* cost, int* iters_run, void* workspace,
                      size_t workspace_bytes, void* stream);
/* Reverse-mode through all L* unrolled iterations (what torch autograd does on the reference).
 * grad_cost (B): upstream gradient of cost.  g4x (B,N,4), g4y (B,M,4): gradient w.r.t. the packed points. */
int shwd_sinkhorn_bwd(const float* x4, const float* y4, int B, int N, int M, int cost_kind, float p, float n_power,
                      float eps, int iters, const float* alpha_hist, const float* beta_hist, const float* row_pc,
                      const float* col_pc, const int* iters_run, const float* grad_cost, float* g4x, float* g4y,
                      void* workspace, size_t workspace_bytes, void* stream);
/* status word of the last persistent launch that used `workspace` (0 ok, 1 = an inter-CTA wait timed out).  A failed
 * launch also overwrites its outputs (cost; g4x, g4y) with NaN on the device, so it can never pass for a result. */
int shwd_sinkhorn_status_offset(void);
/* Two kernel families serve shwd_sinkhorn_fwd / _bwd with identical arithmetic and history format: the flattened-deal
 * persistent kernels (any size; the benchmark shape B=32, N=1024) and the dedicated-CTA "lean" kernels for small
 * problems -- the reference's B=32, N=128..256 runs (train_RUNNER.py:124-127), single pairs of N~1000
 * (Flow_ellipsoid.ipynb cell 8), a strong-scaled batch's 4 pairs per GPU.  The lean forward needs the history
 * (hist_levels = iters+1) and no early stop.
 *   shwd_sinkhorn_lean_regime: 1 if a (B,N,M) problem would take the lean kernels -- callers then keep the history
 *                              even for forward-only calls;
 *   shwd_sinkhorn_set_path:    0 automatic (default), 1 always flattened-deal, 2 lean whenever eligible (A/B timing,
 *                              tests).  Process-wide; not thread-safe against concurrent launches. */
int shwd_sinkhorn_lean_regime(int B, int N, int M);
int shwd_sinkhorn_set_path(int mode);
/* Opt-in dense outputs for the reference's (cost, P, C) return (sinkhorn.py:60): P, C (B,N,M), either nullable. */
int shwd_sinkhorn_plan_dense(const float* x4, const float* y4, int B, int N, int M, int cost_kind, float p,
                             float n_power, float eps, const float* alpha, const float* beta, int level_stride_n,
                             int level_stride_m, float* P, float* C, void* stream);

/* ---- Chamfer distance (pytorch3d.loss.chamfer_distance semantics; call sites train_CD.py:123,161, main_rotation.py:203)
 * x (B,N,3), y (B,M,3) raw contiguous.  d_xy (B,N) = min_j |x_i-y_j|^2, idx_xy (B,N) argmin; d_yx/idx_yx likewise. */
int shwd_chamfer_fwd(const float* x, const float* y, int B, int N, int M, float* d_xy, int* idx_xy, float* d_yx,
                     int* idx_yx, void* stream);
/* gdx (B,N), gdy (B,M): upstream gradients of d_xy, d_yx.  gx (B,N,3), gy (B,M,3) are overwritten. */
int shwd_chamfer_bwd(const float* x, const float* y, int B, int N, int M, const int* idx_xy, const int* idx_yx,
                     const float* gdx, const float* gdy, float* gx, float* gy, void* stream);
/* The reductions of chamfer_distance (point_reduction / batch_reduction / single_directional of the call sites above) on
 * the distances of shwd_chamfer_fwd, one launch, fixed summation order:  loss_b = sx * sum_i d_xy[b,i] + sy * sum_j d_yx[b,j];
 * reduce_batch != 0: loss[0] = sb * sum_b loss_b, else loss[b] = loss_b.  ws: shwd_chamfer_reduce_workspace_bytes(B) bytes,
 * zeroed once by the caller (the kernel re-arms it). */
size_t shwd_chamfer_reduce_workspace_bytes(int B);
int shwd_chamfer_reduce(const float* d_xy, const float* d_yx, int B, int N, int M, float sx, float sy, float sb,
                        int reduce_batch, void* ws, float* loss, void* stream);
/* Backward of that loss without per-point gradient tensors: d loss / d d_xy[b,:] = g[b * g_stride] * cx and
 * d loss / d d_yx[b,:] = g[b * g_stride] * cy, g on the device (g_stride 0: one scalar, 1: one per pair). */
int shwd_chamfer_bwd_uniform(const float* x, const float* y, int B, int N, int M, const int* idx_xy, const int* idx_yx,
                             const float* g, int g_stride, float cx, float cy, float* gx, float* gy, void* stream);

/* ---- sliced paths ------------------------------------------------------------------------------------------------
 * Circle projection (sliced_cost, max_spherical_sliced_w.py:270-279): x (B,N,3), U (P,3,2) -> keys (B,P,N) in [0,1]. */
int shwd_project_circle(const float* x, const float* U, int B, int N, int P, float* keys, void* stream);
int shwd_project_circle_bwd(const float* x, const float* U, int B, int N, int P, const float* gkeys, float* gx,
                            void* stream);
/* Projection backward: rows with N % 4 == 0 (and 16-byte aligned buffers, a grid of at least one CTA per SM) are read four
 * points per lane (512 contiguous bytes per warp and row) -- same bits as the one-point-per-lane kernel.
 * shwd_project_bwd_set_wide: 1 on (default), 0 off (A/B timing, tests).  Process-wide. */
int shwd_project_bwd_set_wide(int on);
/* Line projection (Flow_ellipsoid.ipynb:208-220): x (B,N,3), theta (P,3) -> keys (B,P,N). */
int shwd_project_line(const float* x, const float* theta, int B, int N, int P, float* keys, void* stream);
int shwd_project_line_bwd(const float* theta, int B, int N, int P, const float* gkeys, float* gx, void* stream);
/* Both backward projections with the chain rule of the per-pair slice MEAN folded in: gx[b] = (gw[b] / P) * (...), gw (B,)
 * on the device -- no elementwise scaling pass after the launch (ops.SlicedLossFn.backward). */
int shwd_project_circle_bwd_scaled(const float* x, const float* U, int B, int N, int P, const float* gkeys, const float* gw,
                                   float* gx, void* stream);
int shwd_project_line_bwd_scaled(const float* theta, int B, int N, int P, const float* gkeys, const float* gw, float* gx,
                                 void* stream);
/* Stable segmented sort (torch.sort(stable=True) order, NaN last, -0 == +0): keys (segs,len) -> sorted (segs,len),
 * perm (segs,len) int64.  Replaces torch.sort at max_spherical_sliced_w.py:163-164,224-225,232,235.
 * Rows of up to 16384 keys are sorted in shared memory (workspace 0 bytes); longer rows need the global scratch. */
size_t shwd_segmented_sort_workspace_bytes(int segs, int len);
int shwd_segmented_sort(const float* keys, int segs, int len, float* sorted, int64_t* perm, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Same sort, permutation as int32 (the fused sliced losses keep it on the device only: half the bytes of torch's int64). */
int shwd_segmented_sort_i32(const float* keys, int segs, int len, float* sorted, int32_t* perm, void* workspace,
                            size_t workspace_bytes, void* stream);
/* The sliced losses' own sort: the sort CTA of slice (b, p) computes its keys from the cloud x (B,N,3) and frame p itself
 * (mode 1: circle coordinates through frames (P,3,2), max_spherical_sliced_w.py:270-279; mode 2: line projections through
 * directions (P,3), Flow_ellipsoid.ipynb:214-216) -- no (B,P,N) key array is written or read -- and emits the sorted
 * values (B*P,N) (rebuilt from the sort keys: -0.0 comes back as +0.0) and the int32 stable-sort permutation.  Same keys,
 * same permutation as shwd_project_* followed by shwd_segmented_sort_i32.  N <= shwd_sort_projected_max_points(). */
int shwd_sort_projected_max_points(void);
/* Rows of up to 8192 keys whose keys are spread out (finite, at most 48 per bucket of a 4096-bucket monotone map of the
 * value) are sorted by one bucket pass + an in-bucket rank instead of the radix passes; same (key, index) order, i.e. the
 * same bits.  shwd_sort_set_method: 0 automatic (default), 1 radix passes only (A/B timing, tests).  Process-wide. */
int shwd_sort_set_method(int method);
int shwd_sort_projected(const float* x, const float* frames, int B, int N, int P, int mode, float* sorted, int32_t* perm,
                        void* stream);
/* The same three entry points with ONE FRAME SET PER PAIR, U / frames (B,P,3,2): the batched variant
 * max_spherical_sliced_w_fast.py:258-319 (`Z = randn(B,P,d,2)`, `Us[b]` applied to pair b), circle keys only (mode 1). */
int shwd_project_circle_pp(const float* x, const float* U, int B, int N, int P, float* keys, void* stream);
int shwd_project_circle_bwd_scaled_pp(const float* x, const float* U, int B, int N, int P, const float* gkeys, const float* gw,
                                      float* gx, void* stream);
int shwd_sort_projected_pp(const float* x, const float* frames, int B, int N, int P, int mode, float* sorted, int32_t* perm,
                           void* stream);
/* Circular W1 by level median on sorted circle coordinates (emd1D_circle, max_spherical_sliced_w.py:230-247):
 * us (S,n), vs (S,m) sorted ascending, n + m <= 32768 -> w (S); gus/gvs (nullable) receive dW/d(sorted values). */
size_t shwd_circular_w1_workspace_bytes(int S, int n, int m);
int shwd_circular_w1(const float* us, const float* vs, int S, int n, int m, float* w, float* gus, float* gvs,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Circular W_p^p, p != 1, by bisection on the rotation (binary_search_circle + dCost + Cost,
 * max_spherical_sliced_w.py:25-207), all rounds and the final cost in one launch:
 * us (S,n), vs (S,m) sorted ascending circle coordinates in [0,1] -> w (S) = Cost(theta*), theta (S, nullable) = the
 * rotation found; gus/gvs (nullable) receive d w / d(sorted values) with theta detached (:207).  tm/tp: initial
 * bracket (-1, 1 in the reference), tol: stopping width (eps / max(Lm, Lp) = 1e-7).  n, m <= 32768, n + m <= 56320.
 * No workspace is needed any more (the uniform CDFs are evaluated in registers): _workspace_bytes returns 0 and the
 * workspace arguments are ignored; both are kept so that existing callers keep linking. */
size_t shwd_circular_wp_workspace_bytes(int S, int n, int m);
int shwd_circular_wp(const float* us, const float* vs, int S, int n, int m, float p, float tm, float tp, float tol,
                     float* w, float* gus, float* gvs, float* theta, void* workspace, size_t workspace_bytes,
                     void* stream);
/* Equal cloud sizes take closed-form searches while the rotation sits on the 1/n grid (powers of two) or safely off it (any
 * length) -- same bits, ~4x fewer instructions per round; circular_wp.cu.  shwd_circular_wp_set_dyadic: 1 on (default), 0 off
 * (A/B timing, tests).  Process-wide. */
int shwd_circular_wp_set_dyadic(int on);
/* The same with the reference's u_weights / v_weights (binary_search_circle, max_spherical_sliced_w.py:117,156-170): ucdf
 * (S,n) / vcdf (S,m) are the per-slice CDF tables cumsum(weights[..., sorter], -1) the reference forms (non-decreasing);
 * the closed-form searches of the uniform kernel become bisections on the tables.  gcu (S,n) / gcv (S,m) (nullable; need
 * gus and gvs) receive d w / d ucdf, d w / d vcdf -- the path by which the reference's autograd reaches the weights (the
 * merged CDF axis of the final Cost, :93-95).  n + m <= 28160. */
int shwd_circular_wp_weighted(const float* us, const float* vs, const float* ucdf, const float* vcdf, int S, int n, int m,
                              float p, float tm, float tp, float tol, float* w, float* gus, float* gvs, float* gcu, float* gcv,
                              float* theta, void* stream);
/* Euclidean sliced W on sorted projections (Flow_ellipsoid.ipynb:217-219): xs, ys (S,n) sorted -> acc (S) =
 * sum_n |xs-ys|^p ; gxs/gys (nullable) receive d acc / d(sorted values). */
int shwd_euclid_sw(const float* xs, const float* ys, int S, int n, float p, float* acc, float* gxs, float* gys,
                   void* stream);
/* The three 1-D reductions with the sort's backward folded in: perm_u/perm_v (S,n)/(S,m) int32 from
 * shwd_segmented_sort_i32; gku/gkv receive d w / d(UNSORTED keys), i.e. gku[s][perm_u[s][k]] = d w_s / d us[s][k] -- what
 * autograd's scatter through torch.sort (max_spherical_sliced_w.py:163-164, 224-225; Flow_ellipsoid.ipynb:217) produces,
 * without the sorted-order gradient ever reaching HBM. */
int shwd_circular_w1_scatter(const float* us, const float* vs, const int32_t* perm_u, const int32_t* perm_v, int S, int n,
                             int m, float* w, float* gku, float* gkv, void* stream);
int shwd_circular_wp_scatter(const float* us, const float* vs, const int32_t* perm_u, const int32_t* perm_v, int S, int n,
                             int m, float p, float tm, float tp, float tol, float* w, float* gku, float* gkv, float* theta,
                             void* workspace, size_t workspace_bytes, void* stream);
int shwd_euclid_sw_scatter(const float* xs, const float* ys, const int32_t* perm_x, const int32_t* perm_y, int S, int n,
                           float p, float* acc, float* gkx, float* gky, void* stream);
/* Scatter gradients of sorted values back through the permutation: gkeys[seg][perm[seg][k]] = gsorted[seg][k]. */
int shwd_unsort(const float* gsorted, const int64_t* perm, int segs, int len, float* gkeys, void* stream);

/* ---- the learned sphere map phi: a stack of Residual flows x <- x + LipschitzMLP(x), fused --------------------------
 * Replaces Norm_Flow_structure.forward ("Residual", s2_wasserstein.py:144-163; normflows_ishikawa/flows/residual.py:63-68,
 * nets/lipschitz.py:14-68,223-274,642-648) and its autograd.  x, y, gy, gx: (npts,3).
 * params: n_layers x shwd_resflow_params_per_layer() RAW parameters per flow layer, laid out
 *   [W0 8x3 | b0 8 | (W 8x8 | b 8) x5 | W6 3x8 | b6 3 | beta0..beta6]      (row-major out x in; beta = Swish parameter)
 * uv: n_layers x shwd_resflow_uv_per_layer() frozen power-iteration vectors [u0 8 | v0 3 | (u 8 | v 8) x5 | u6 3 | v6 8].
 * The kernels form W_k / max(1, (u_k^T W_k v_k)/coeff) and softplus(beta_k) themselves; gparams receives the gradient
 * w.r.t. the RAW parameters (same layout), reduced in a fixed order (bit-reproducible).  n_layers <= 8. */
int shwd_resflow_params_per_layer(void);
int shwd_resflow_uv_per_layer(void);
size_t shwd_resflow_workspace_bytes(int npts, int n_layers);
int shwd_resflow_fwd(const float* x, int npts, const float* params, const float* uv, int n_layers, float coeff, float* y,
                     void* stream);
int shwd_resflow_bwd(const float* x, const float* gy, int npts, const float* params, const float* uv, int n_layers,
                     float coeff, float* gx, float* gparams, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the learned sphere map phi, Planar variant: a stack of planar flows z <- z + u^ tanh(lin + b), fused -------------
 * Replaces Norm_Flow_structure.forward ("Planar", s2_wasserstein.py:140-143,160-163; normflows_ishikawa/flows/planar.py:49-60:
 * u^ = u + (log(1 + exp(<w,u>)) - 1 - <w,u>) w / |w|^2, act = tanh) and its autograd; the log-determinant is discarded by
 * the caller and not computed.  params: n_layers x shwd_planar_params_per_layer() RAW parameters [u 3 | w 3 | b 1] per flow
 * layer, n_layers <= shwd_planar_max_layers().  `lin` is the reference's sum over dim 1 of w * z:
 *   clouds == 0: x, y, gy, gx are (npts,3) and lin_n = <w, z_n>                       (the un-batched branch);
 *   clouds  > 0: they are (clouds,npts,3) and lin_bc = w_c sum_n z_bnc               (what batched clouds yield in the
 *                reference: dim 1 is the point axis); colsum (clouds,3), float64, receives the per-cloud column sums
 *                of x in the forward and is handed back to the backward (x itself is not read again there).
 * gparams: gradient w.r.t. the raw parameters (same layout), reduced in a fixed order (bit-reproducible). */
int shwd_planar_max_layers(void);
int shwd_planar_params_per_layer(void);
size_t shwd_planar_workspace_bytes(int clouds, int npts, int n_layers);
int shwd_planar_fwd(const float* x, int clouds, int npts, const float* params, int n_layers, float* y, double* colsum,
                    void* stream);
int shwd_planar_bwd(const float* x, const float* gy, const double* colsum, int clouds, int npts, const float* params,
                    int n_layers, float* gx, float* gparams, void* workspace, size_t workspace_bytes, void* stream);

/* ---- exact optimal assignment (uniform weights, n == m): the exact solve behind ot.emd2 on the W_COS path ------------
 * Replaces the per-pair CPU solve `ot.emd2(a_i, b_i, C_i)` (s2_wasserstein.py:39-50, 99-110; main_rotation.py:63-79) by a
 * float64 forward auction with epsilon-scaling on the on-the-fly cost, one CTA per pair.  x4, y4: packed points (B,N,4)
 * (normalised by shwd_sphere_map_fwd for the cosine cost kinds).  sigma (B,N) int32: sigma[i] = point of y matched to
 * point i of x; emd2_b = (1/N) sum_i C(x_i, y_sigma[i]).  prices (B,N) float64 (nullable): the dual prices; rounds (B)
 * (nullable): bidding rounds used; status (1): set to 1 if a pair hit the round limit.  N <= shwd_exact_assignment_max_points(). */
int shwd_exact_assignment_max_points(void);
int shwd_exact_assignment(const float* x4, const float* y4, int B, int N, int cost_kind, float p, float n_power, int* sigma,
                          double* prices, int* rounds, int* status, void* stream);
/* The same exact solve on an explicit cost matrix C (B,N,N) float32, costs >= 0: the drop-in for ot.emd2(a, b, M) with
 * uniform weights as called at Comparison_Wasserstein_with_Chamfer_distance/main_rotation.py:63-79 (POT_loss) and in the
 * notebooks' W2 metric (Flow_ellipsoid.ipynb cell 8). */
int shwd_exact_assignment_dense(const float* C, int B, int N, int* sigma, double* prices, int* rounds, int* status,
                                void* stream);

/* ---- data side: random rigid transform (+ optional sensor noise) of a device-resident batch ---------------------------
 * Replaces Dataset_Transformation.__call__ / create_pose_7d / qrot (data_utils/Data_set_maker.py:40-52,173-230), run by the
 * reference per item on the CPU.  src, out (B,N,3); pose7 (B,7) = (un-normalised quaternion w,x,y,z | translation);
 * rotation (B,3,3), nullable = igt_rotation.  noise_std > 0 adds N(0, noise_std^2) to src first (add_noise, :13-22), from a
 * counter-based generator keyed by (seed, element). */
int shwd_rigid_transform(const float* src, const float* pose7, int B, int N, float noise_std, unsigned long long seed,
                         float* out, float* rotation, void* stream);

/* ---- measurement helpers (bench.py): FP32-FMA and MUFU issue-rate microbenchmarks -------------------------------
 * out (grid*block floats) scratch; returns the number of lane-ops each launch performs in *ops. */
int shwd_peak_fp32(float* out, int iters, double* ops, void* stream);
int shwd_peak_mufu(float* out, int iters, double* ops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHWD_H_ */

"""torch.autograd.Function wrappers over the C ABI (include/shwd.h).  PyTorch supplies device memory, the current
stream and autograd plumbing; all arithmetic happens in libshwd_b200.so.  CUDA tensors only -- no CPU fallback."""
import torch

from . import _lib
from ._lib import COST_KINDS, MAP_CENTER, MAP_NORMALIZE


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _as_cloud(x, name):
    """(B,N,3) or (N,3), any dtype/strides -> contiguous float32 (B,N,3) on CUDA, plus the un-batched flag."""
    if not isinstance(x, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not x.is_cuda:
        raise RuntimeError("%s must live on a CUDA device: the B200 loss path has no CPU fallback" % name)
    unbatched = x.dim() == 2
    if unbatched:
        x = x.unsqueeze(0)
    if x.dim() != 3 or x.shape[-1] != 3:
        raise ValueError("%s must have shape (B,N,3) or (N,3), got %s" % (name, tuple(x.shape)))
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous(), unbatched


# ---------------------------------------------------------------------------------------------------------------------
class SphereMapFn(torch.autograd.Function):
    """x (B,N,3) -> packed (B,N,4) = (x^, 1/||x_c||) and the per-cloud regulariser sum_n | ||x_n|| - 1 |."""

    @staticmethod
    def forward(ctx, x, flags, want_reg):
        B, N, _ = x.shape
        xh4 = torch.empty(B, N, 4, device=x.device, dtype=torch.float32)
        reg = torch.empty(B, device=x.device, dtype=torch.float32) if want_reg else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().shwd_sphere_map_fwd(_ptr(x), _ptr(xh4), _ptr(reg), B, N, flags, _stream()), "shwd_sphere_map_fwd")
        ctx.save_for_backward(x, xh4)
        ctx.flags = flags
        ctx.want_reg = want_reg
        if want_reg:
            return xh4, reg
        return xh4, x.new_zeros(())

    @staticmethod
    def backward(ctx, g4, greg):
        x, xh4 = ctx.saved_tensors
        B, N, _ = x.shape
        gx = torch.empty_like(x)
        g4 = g4.contiguous() if g4 is not None else None
        greg = greg.contiguous() if (ctx.want_reg and greg is not None) else None
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().shwd_sphere_map_bwd(_ptr(x), _ptr(xh4), _ptr(g4), _ptr(greg), _ptr(gx), B, N, ctx.flags, _stream()),
                       "shwd_sphere_map_bwd")
        return gx, None, None


def sphere_map(x, center=True, normalize=True):
    """Centre (train_W_COS.py:167-168) and project on the unit sphere (x / max(||x||,1e-8), as F.cosine_similarity does
    at s2_wasserstein.py:122).  Returns a tensor shaped like ``x``."""
    xc, unb = _as_cloud(x, "x")
    flags = (MAP_CENTER if center else 0) | (MAP_NORMALIZE if normalize else 0)
    xh4, _ = SphereMapFn.apply(xc, flags, False)
    out = xh4[..., :3]
    return out[0] if unb else out


def flow_regularization(x):
    """sum_{b,n} | ||x_bn|| - 1 |  (regularization_of_normalizing_flow, s2_wasserstein.py:224-232)."""
    xc, _ = _as_cloud(x, "x")
    _, reg = SphereMapFn.apply(xc, 0, True)
    return reg.sum()


# ---------------------------------------------------------------------------------------------------------------------
class EntropicOTFn(torch.autograd.Function):
    """cost_b = sum_ij P_ij C_ij after L log-domain Sinkhorn iterations on an on-the-fly cost; reverse mode through all
    iterations.  Inputs are raw clouds; for the cosine cost kinds the sphere map (normalisation, optional centring) is
    applied inside so its Jacobian is part of the same backward."""

    @staticmethod
    def forward(ctx, x, y, kind, p, n_power, eps, iters, thresh, center, need_grad):
        lib = _lib.lib()
        B, N, _ = x.shape
        M = y.shape[1]
        dev = x.device
        cosine = kind in (_lib.COST_GEODESIC, _lib.COST_ONE_MINUS_COS)
        flags = (MAP_NORMALIZE if cosine else 0) | (MAP_CENTER if center else 0)
        f32 = dict(device=dev, dtype=torch.float32)
        x4 = torch.empty(B, N, 4, **f32)
        y4 = torch.empty(B, M, 4, **f32)
        # the lean kernels for small problems exchange potentials through the write-once history planes: keep them
        keep = bool(need_grad) or thresh > 0 or bool(lib.shwd_sinkhorn_lean_regime(B, N, M))
        HL = iters + 1 if keep else 1
        alpha = torch.empty(2, B, HL, N, **f32)  # plane 0: iterates k*u; plane 1: float32 rounding residuals
        beta = torch.empty(2, B, HL, M, **f32)
        row_pc = torch.empty(B, N, **f32)
        col_pc = torch.empty(B, M, **f32)
        cost = torch.empty(B, **f32)
        iters_run = torch.empty(1, device=dev, dtype=torch.int32)
        wsb = lib.shwd_sinkhorn_workspace_bytes(B, N, M, iters)
        ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            s = _stream()
            _lib.check(lib.shwd_sphere_map_fwd(_ptr(x), _ptr(x4), None, B, N, flags, s), "shwd_sphere_map_fwd")
            _lib.check(lib.shwd_sphere_map_fwd(_ptr(y), _ptr(y4), None, B, M, flags, s), "shwd_sphere_map_fwd")
            _lib.check(lib.shwd_sinkhorn_fwd(_ptr(x4), _ptr(y4), B, N, M, kind, p, n_power, eps, iters, thresh, HL, _ptr(alpha),
                                             _ptr(beta), _ptr(row_pc), _ptr(col_pc), _ptr(cost), _ptr(iters_run), _ptr(ws), wsb, s),
                       "shwd_sinkhorn_fwd")
        # (a launch whose inter-CTA wait timed out NaNs `cost` itself, on the device: csrc/sinkhorn.cu poison_on_failure_kernel)
        ctx.save_for_backward(x, y, x4, y4, alpha, beta, row_pc, col_pc, iters_run)
        ctx.cfg = (kind, p, n_power, eps, iters, flags, keep)
        ctx.mark_non_differentiable(alpha, beta, iters_run, ws)
        return cost, alpha, beta, iters_run, ws

    @staticmethod
    def backward(ctx, gcost, _ga, _gb, _gi, _gw):
        lib = _lib.lib()
        x, y, x4, y4, alpha, beta, row_pc, col_pc, iters_run = ctx.saved_tensors
        kind, p, n_power, eps, iters, flags, keep = ctx.cfg
        if not keep:
            raise RuntimeError("EntropicOTFn: forward ran without history (need_grad=False); cannot backpropagate")
        B, N, _ = x.shape
        M = y.shape[1]
        dev = x.device
        gcost = gcost.contiguous().float()
        g4x = torch.empty(B, N, 4, device=dev, dtype=torch.float32)
        g4y = torch.empty(B, M, 4, device=dev, dtype=torch.float32)
        gx = torch.empty_like(x)
        gy = torch.empty_like(y)
        wsb = lib.shwd_sinkhorn_workspace_bytes(B, N, M, iters)
        ws = torch.empty(wsb, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            s = _stream()
            _lib.check(lib.shwd_sinkhorn_bwd(_ptr(x4), _ptr(y4), B, N, M, kind, p, n_power, eps, iters, _ptr(alpha), _ptr(beta),
                                             _ptr(row_pc), _ptr(col_pc), _ptr(iters_run), _ptr(gcost), _ptr(g4x), _ptr(g4y), _ptr(ws),
                                             wsb, s), "shwd_sinkhorn_bwd")
            _lib.check(lib.shwd_sphere_map_bwd(_ptr(x), _ptr(x4), _ptr(g4x), None, _ptr(gx), B, N, flags, s), "shwd_sphere_map_bwd")
            _lib.check(lib.shwd_sphere_map_bwd(_ptr(y), _ptr(y4), _ptr(g4y), None, _ptr(gy), B, M, flags, s), "shwd_sphere_map_bwd")
        return gx, gy, None, None, None, None, None, None, None, None  # (NaN on a failed launch, see above)


class EntropicOTResult:
    """Per-pair costs plus lazily materialised extras (duals, dense plan/cost)."""

    def __init__(self, cost, alpha, beta, iters_run, ws, x, y, cfg):
        self.cost, self._alpha, self._beta, self._iters_run, self._ws = cost, alpha, beta, iters_run, ws
        self._x, self._y, self._cfg = x, y, cfg

    def status(self):
        """0 if every inter-CTA wait of the forward launch completed (synchronises)."""
        return int(self._ws[:4].view(torch.int32).item())

    def iterations(self):
        return int(self._iters_run.item())

    def duals(self):
        """(u, v) of the iterate used, in the reference's units (alpha / k)."""
        kind, p, n_power, eps, iters, center = self._cfg
        lvl = self.iterations() if self._alpha.shape[2] > 1 else 0
        inv_k = eps / 1.4426950408889634
        return self._alpha[0, :, lvl] * inv_k, self._beta[0, :, lvl] * inv_k

    def dense(self, want_plan=True, want_cost=True):
        """(P, C) as (B,N,M) tensors -- the reference's extra return values (sinkhorn.py:60).  Small problems only."""
        lib = _lib.lib()
        kind, p, n_power, eps, iters, center = self._cfg
        x, y = self._x, self._y
        B, N, _ = x.shape
        M = y.shape[1]
        cosine = kind in (_lib.COST_GEODESIC, _lib.COST_ONE_MINUS_COS)
        flags = (MAP_NORMALIZE if cosine else 0) | (MAP_CENTER if center else 0)
        x4 = torch.empty(B, N, 4, device=x.device, dtype=torch.float32)
        y4 = torch.empty(B, M, 4, device=x.device, dtype=torch.float32)
        P = torch.empty(B, N, M, device=x.device, dtype=torch.float32) if want_plan else None
        C = torch.empty(B, N, M, device=x.device, dtype=torch.float32) if want_cost else None
        HL = self._alpha.shape[2]
        lvl = self.iterations() if HL > 1 else 0
        a = self._alpha[0, :, lvl]
        b = self._beta[0, :, lvl]
        with torch.cuda.device(x.device):
            s = _stream()
            _lib.check(lib.shwd_sphere_map_fwd(_ptr(x), _ptr(x4), None, B, N, flags, s), "shwd_sphere_map_fwd")
            _lib.check(lib.shwd_sphere_map_fwd(_ptr(y), _ptr(y4), None, B, M, flags, s), "shwd_sphere_map_fwd")
            _lib.check(lib.shwd_sinkhorn_plan_dense(_ptr(x4), _ptr(y4), B, N, M, kind, p, n_power, eps, a.data_ptr(), b.data_ptr(),
                                                    HL * N, HL * M, _ptr(P), _ptr(C), s), "shwd_sinkhorn_plan_dense")
        return P, C


def entropic_ot(x, y, kind="geodesic", p=2.0, eps=0.01, iters=100, n_power=1.0, early_stop_thresh=0.0, center=False):
    """Per-pair entropic OT cost (B,) between clouds x (B,N,3) and y (B,M,3) [or un-batched (N,3),(M,3) -> (1,)].

    Follows the reference recurrence (Comparison_.../losses/sinkhorn.py:24-58) on the cost matrix ``kind``
    (s2_wasserstein.py:52-63,112-123) without ever forming an N x M tensor.  Differentiable w.r.t. x and y through all
    iterations.  Returns an :class:`EntropicOTResult`."""
    xc, _ = _as_cloud(x, "x")
    yc, _ = _as_cloud(y, "y")
    if xc.shape[0] != yc.shape[0]:
        raise ValueError("batch sizes differ: %d vs %d" % (xc.shape[0], yc.shape[0]))
    if xc.device != yc.device:
        raise RuntimeError("x and y must be on the same CUDA device")
    k = COST_KINDS[kind] if isinstance(kind, str) else int(kind)
    need_grad = torch.is_grad_enabled() and (xc.requires_grad or yc.requires_grad)
    cost, alpha, beta, iters_run, ws = EntropicOTFn.apply(xc, yc, k, float(p), float(n_power), float(eps), int(iters),
                                                          float(early_stop_thresh), bool(center), need_grad)
    return EntropicOTResult(cost, alpha, beta, iters_run, ws, xc.detach(), yc.detach(),
                            (k, float(p), float(n_power), float(eps), int(iters), bool(center)))


# ---------------------------------------------------------------------------------------------------------------------
class ChamferFn(torch.autograd.Function):
    """(x (B,N,3), y (B,M,3)) -> (d_xy (B,N), d_yx (B,M)): nearest-neighbour squared distances, both directions."""

    @staticmethod
    def forward(ctx, x, y):
        x = x.contiguous()
        y = y.contiguous()
        B, N, _ = x.shape
        M = y.shape[1]
        dev = x.device
        d_xy = torch.empty(B, N, device=dev, dtype=torch.float32)
        d_yx = torch.empty(B, M, device=dev, dtype=torch.float32)
        i_xy = torch.empty(B, N, device=dev, dtype=torch.int32)
        i_yx = torch.empty(B, M, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().shwd_chamfer_fwd(_ptr(x), _ptr(y), B, N, M, _ptr(d_xy), _ptr(i_xy), _ptr(d_yx), _ptr(i_yx), _stream()),
                       "shwd_chamfer_fwd")
        ctx.save_for_backward(x, y, i_xy, i_yx)
        ctx.mark_non_differentiable(i_xy, i_yx)
        return d_xy, d_yx, i_xy, i_yx

    @staticmethod
    def backward(ctx, gdx, gdy, _a, _b):
        x, y, i_xy, i_yx = ctx.saved_tensors
        B, N, _ = x.shape
        M = y.shape[1]
        gdx = torch.zeros(B, N, device=x.device) if gdx is None else gdx.contiguous().float()
        gdy = torch.zeros(B, M, device=x.device) if gdy is None else gdy.contiguous().float()
        gx = torch.empty_like(x)
        gy = torch.empty_like(y)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().shwd_chamfer_bwd(_ptr(x), _ptr(y), B, N, M, _ptr(i_xy), _ptr(i_yx), _ptr(gdx), _ptr(gdy), _ptr(gx),
                                                   _ptr(gy), _stream()), "shwd_chamfer_bwd")
        return gx, gy


class ChamferLossFn(torch.autograd.Function):
    """chamfer_distance as ONE autograd node: nearest neighbours, the point / batch reductions (fixed order) and, backward,
    one launch that reads the upstream scalar on the device -- no (B,N) gradient tensors, no ATen reduce / elementwise
    kernels.  ``apply(x, y, sx, sy, sb, reduce_batch)`` -> loss () or (B,)  (pytorch3d semantics: train_CD.py:123,161)."""

    @staticmethod
    def forward(ctx, x, y, sx, sy, sb, reduce_batch):
        lib = _lib.lib()
        x = x.contiguous()
        y = y.contiguous()
        B, N, _ = x.shape
        M = y.shape[1]
        dev = x.device
        d_xy = torch.empty(B, N, device=dev, dtype=torch.float32)
        d_yx = torch.empty(B, M, device=dev, dtype=torch.float32)
        i_xy = torch.empty(B, N, device=dev, dtype=torch.int32)
        i_yx = torch.empty(B, M, device=dev, dtype=torch.int32)
        loss = torch.zeros((1,) if reduce_batch else (B,), device=dev, dtype=torch.float32)
        ws = torch.zeros(lib.shwd_chamfer_reduce_workspace_bytes(B), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            s = _stream()
            _lib.check(lib.shwd_chamfer_fwd(_ptr(x), _ptr(y), B, N, M, _ptr(d_xy), _ptr(i_xy), _ptr(d_yx), _ptr(i_yx), s), "shwd_chamfer_fwd")
            _lib.check(lib.shwd_chamfer_reduce(_ptr(d_xy), _ptr(d_yx), B, N, M, float(sx), float(sy), float(sb), int(reduce_batch),
                                               _ptr(ws), _ptr(loss), s), "shwd_chamfer_reduce")
        ctx.save_for_backward(x, y, i_xy, i_yx)
        ctx.scales = (float(sx), float(sy), float(sb), bool(reduce_batch))
        return loss.reshape(()) if reduce_batch else loss

    @staticmethod
    def backward(ctx, g):
        x, y, i_xy, i_yx = ctx.saved_tensors
        sx, sy, sb, reduce_batch = ctx.scales
        B, N, _ = x.shape
        M = y.shape[1]
        g = g.contiguous().float()
        gx = torch.empty_like(x)
        gy = torch.empty_like(y)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().shwd_chamfer_bwd_uniform(_ptr(x), _ptr(y), B, N, M, _ptr(i_xy), _ptr(i_yx), _ptr(g),
                                                           0 if reduce_batch else 1, sx * sb, sy * sb, _ptr(gx), _ptr(gy), _stream()),
                       "shwd_chamfer_bwd_uniform")
        return gx, gy, None, None, None, None


def chamfer_loss(x, y, point_mean=True, batch_reduction="mean", single_directional=False):
    """The fused chamfer_distance value (see ChamferLossFn)."""
    xc, _ = _as_cloud(x, "x")
    yc, _ = _as_cloud(y, "y")
    if xc.shape[0] != yc.shape[0]:
        raise ValueError("batch sizes differ: %d vs %d" % (xc.shape[0], yc.shape[0]))
    B, N, M = xc.shape[0], xc.shape[1], yc.shape[1]
    sx = 1.0 / N if point_mean else 1.0
    sy = 0.0 if single_directional else (1.0 / M if point_mean else 1.0)
    sb = 1.0 / max(B, 1) if batch_reduction == "mean" else 1.0
    return ChamferLossFn.apply(xc, yc, sx, sy, sb, batch_reduction is not None)


def chamfer_nn(x, y):
    """Nearest-neighbour squared distances and indices in both directions: (d_xy, d_yx, idx_xy, idx_yx)."""
    xc, _ = _as_cloud(x, "x")
    yc, _ = _as_cloud(y, "y")
    if xc.shape[0] != yc.shape[0]:
        raise ValueError("batch sizes differ: %d vs %d" % (xc.shape[0], yc.shape[0]))
    return ChamferFn.apply(xc, yc)


# ---------------------------------------------------------------------------------------------------------------------
class ProjectCircleFn(torch.autograd.Function):
    """x (B,N,3), U (P,3,2) -> circle coordinates (B,P,N) in [0,1]  (sliced_cost, max_spherical_sliced_w.py:270-279)."""

    @staticmethod
    def forward(ctx, x, U):
        x = x.contiguous()
        U = U.contiguous()  # torch.linalg.qr returns column-major batches
        B, N, _ = x.shape
        P = U.shape[0]
        keys = torch.empty(B, P, N, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().shwd_project_circle(_ptr(x), _ptr(U), B, N, P, _ptr(keys), _stream()), "shwd_project_circle")
        ctx.save_for_backward(x, U)
        return keys

    @staticmethod
    def backward(ctx, gk):
        x, U = ctx.saved_tensors
        B, N, _ = x.shape
        gx = torch.empty_like(x)
        gk = gk.contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().shwd_project_circle_bwd(_ptr(x), _ptr(U), B, N, U.shape[0], _ptr(gk), _ptr(gx), _stream()),
                       "shwd_project_circle_bwd")
        return gx, None


class ProjectLineFn(torch.autograd.Function):
    """x (B,N,3), theta (P,3) -> projections (B,P,N)  (Flow_ellipsoid.ipynb:214-216)."""

    @staticmethod
    def forward(ctx, x, theta):
        x = x.contiguous()
        theta = theta.contiguous()
        B, N, _ = x.shape
        P = theta.shape[0]
        keys = torch.empty(B, P, N, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().shwd_project_line(_ptr(x), _ptr(theta), B, N, P, _ptr(keys), _stream()), "shwd_project_line")
        ctx.save_for_backward(theta)
        ctx.shape = (B, N)
        return keys

    @staticmethod
    def backward(ctx, gk):
        (theta,) = ctx.saved_tensors
        B, N = ctx.shape
        gx = torch.empty(B, N, 3, device=gk.device, dtype=torch.float32)
        gk = gk.contiguous()
        with torch.cuda.device(gk.device):
            _lib.check(_lib.lib().shwd_project_line_bwd(_ptr(theta), B, N, theta.shape[0], _ptr(gk), _ptr(gx), _stream()),
                       "shwd_project_line_bwd")
        return gx, None


def segmented_sort_raw(keys):
    """keys (..., len) float32 CUDA -> (sorted, perm int64), perm identical to torch.sort(keys, -1, stable=True)."""
    lib = _lib.lib()
    if not keys.is_cuda:
        raise RuntimeError("keys must live on a CUDA device: no CPU fallback")
    k = keys.contiguous().float()
    length = k.shape[-1]
    segs = k.numel() // max(length, 1)
    out = torch.empty_like(k)
    perm = torch.empty(k.shape, device=k.device, dtype=torch.int64)
    wsb = lib.shwd_segmented_sort_workspace_bytes(segs, length)
    ws = torch.empty(max(wsb, 8), device=k.device, dtype=torch.uint8)
    with torch.cuda.device(k.device):
        _lib.check(lib.shwd_segmented_sort(_ptr(k), segs, length, _ptr(out), _ptr(perm), _ptr(ws), wsb, _stream()), "shwd_segmented_sort")
    return out, perm


class SegmentedSortFn(torch.autograd.Function):
    """Differentiable stable sort along the last dim: gradient of the sorted values scatters back through the
    permutation (what autograd does for torch.sort's values output)."""

    @staticmethod
    def forward(ctx, keys):
        out, perm = segmented_sort_raw(keys)
        ctx.save_for_backward(perm)
        ctx.mark_non_differentiable(perm)
        return out, perm

    @staticmethod
    def backward(ctx, gs, _gp):
        (perm,) = ctx.saved_tensors
        gs = gs.contiguous().float()
        length = perm.shape[-1]
        segs = perm.numel() // max(length, 1)
        gk = torch.empty_like(gs)
        with torch.cuda.device(gs.device):
            _lib.check(_lib.lib().shwd_unsort(_ptr(gs), _ptr(perm), segs, length, _ptr(gk), _stream()), "shwd_unsort")
        return gk


class CircularW1Fn(torch.autograd.Function):
    """Sorted circle coordinates us (S,n), vs (S,m) -> W1 per slice (S)  (emd1D_circle, max_spherical_sliced_w.py:230-247)."""

    @staticmethod
    def forward(ctx, us, vs):
        us = us.contiguous()
        vs = vs.contiguous()
        S, n = us.shape
        m = vs.shape[1]
        w = torch.empty(S, device=us.device, dtype=torch.float32)
        gus = torch.empty_like(us)
        gvs = torch.empty_like(vs)
        with torch.cuda.device(us.device):
            _lib.check(_lib.lib().shwd_circular_w1(_ptr(us), _ptr(vs), S, n, m, _ptr(w), _ptr(gus), _ptr(gvs), None, 0, _stream()),
                       "shwd_circular_w1")
        ctx.save_for_backward(gus, gvs)
        return w

    @staticmethod
    def backward(ctx, gw):
        gus, gvs = ctx.saved_tensors
        g = gw.contiguous().unsqueeze(1)
        return gus * g, gvs * g


class CircularWpFn(torch.autograd.Function):
    """Sorted circle coordinates us (S,n), vs (S,m) -> W_p^p per slice (S), p != 1: the whole bisection of
    binary_search_circle (max_spherical_sliced_w.py:117-207) in one launch; gradients flow through the sorted values
    only (theta is detached in the reference, :207)."""

    @staticmethod
    def forward(ctx, us, vs, p, tm, tp, tol):
        us = us.contiguous()
        vs = vs.contiguous()
        S, n = us.shape
        m = vs.shape[1]
        lib = _lib.lib()
        w = torch.empty(S, device=us.device, dtype=torch.float32)
        theta = torch.empty(S, device=us.device, dtype=torch.float32)
        gus = torch.empty_like(us)
        gvs = torch.empty_like(vs)
        wsb = lib.shwd_circular_wp_workspace_bytes(S, n, m)
        ws = torch.empty(max(wsb, 8), device=us.device, dtype=torch.uint8)
        with torch.cuda.device(us.device):
            _lib.check(lib.shwd_circular_wp(_ptr(us), _ptr(vs), S, n, m, float(p), float(tm), float(tp), float(tol), _ptr(w),
                                            _ptr(gus), _ptr(gvs), _ptr(theta), _ptr(ws), wsb, _stream()), "shwd_circular_wp")
        ctx.save_for_backward(gus, gvs)
        ctx.mark_non_differentiable(theta)
        return w, theta

    @staticmethod
    def backward(ctx, gw, _gt):
        gus, gvs = ctx.saved_tensors
        g = gw.contiguous().unsqueeze(1)
        return gus * g, gvs * g, None, None, None, None


class CircularWpWeightedFn(torch.autograd.Function):
    """binary_search_circle with u_weights / v_weights (max_spherical_sliced_w.py:117,156-170): sorted coordinates us (S,n),
    vs (S,m) and the per-slice CDF tables ucdf = cumsum(u_weights[..., sorter]), vcdf -> W_p^p per slice.  Gradients w.r.t. the
    sorted values and w.r.t. the CDF tables (the path by which the reference's autograd reaches the weights: the merged CDF
    axis of the final Cost, :93-95; theta is detached, :207)."""

    @staticmethod
    def forward(ctx, us, vs, ucdf, vcdf, p, tm, tp, tol):
        us, vs, ucdf, vcdf = (t.contiguous().float() for t in (us, vs, ucdf, vcdf))
        S, n = us.shape
        m = vs.shape[1]
        if ucdf.shape != us.shape or vcdf.shape != vs.shape:
            raise ValueError("CDF tables must have the shapes of the value rows")
        w = torch.empty(S, device=us.device, dtype=torch.float32)
        theta = torch.empty(S, device=us.device, dtype=torch.float32)
        gus, gvs, gcu, gcv = (torch.empty_like(t) for t in (us, vs, ucdf, vcdf))
        with torch.cuda.device(us.device):
            _lib.check(_lib.lib().shwd_circular_wp_weighted(_ptr(us), _ptr(vs), _ptr(ucdf), _ptr(vcdf), S, n, m, float(p), float(tm),
                                                            float(tp), float(tol), _ptr(w), _ptr(gus), _ptr(gvs), _ptr(gcu),
                                                            _ptr(gcv), _ptr(theta), _stream()), "shwd_circular_wp_weighted")
        ctx.save_for_backward(gus, gvs, gcu, gcv)
        ctx.mark_non_differentiable(theta)
        return w, theta

    @staticmethod
    def backward(ctx, gw, _gt):
        gus, gvs, gcu, gcv = ctx.saved_tensors
        g = gw.contiguous().unsqueeze(1)
        return gus * g, gvs * g, gcu * g, gcv * g, None, None, None, None


def weight_cdf(weights, perm, S, length, device):
    """cumsum(weights[..., sorter], -1) per slice (max_spherical_sliced_w.py:166-170), or the uniform 1/len weights when
    ``weights`` is None.  The reference accumulates on the CPU, where torch.cumsum carries a float64 accumulator and rounds
    every prefix to float32; the device scan is run in float64 for the same prefixes."""
    if weights is None:
        ws = torch.full((S, length), 1 / length, dtype=torch.float32, device=device)
    else:
        ws = weights.to(device=device, dtype=torch.float32)
        if perm is not None:  # (n,) weights: the reference's ``weights[..., sorter]``; per-slice rows are gathered row by row
            ws = ws[perm.long()] if ws.dim() == 1 else torch.gather(ws.expand(S, length), -1, perm.long())
        ws = ws.expand(S, length)
    return torch.cumsum(ws.double(), -1).float()


class EuclidSWFn(torch.autograd.Function):
    """Sorted projections xs, ys (S,n) -> sum_n |xs-ys|^p per slice (S)  (Flow_ellipsoid.ipynb:217-219)."""

    @staticmethod
    def forward(ctx, xs, ys, p):
        xs = xs.contiguous()
        ys = ys.contiguous()
        S, n = xs.shape
        acc = torch.empty(S, device=xs.device, dtype=torch.float32)
        gxs = torch.empty_like(xs)
        gys = torch.empty_like(ys)
        with torch.cuda.device(xs.device):
            _lib.check(_lib.lib().shwd_euclid_sw(_ptr(xs), _ptr(ys), S, n, float(p), _ptr(acc), _ptr(gxs), _ptr(gys), _stream()),
                       "shwd_euclid_sw")
        ctx.save_for_backward(gxs, gys)
        return acc

    @staticmethod
    def backward(ctx, ga):
        gxs, gys = ctx.saved_tensors
        g = ga.contiguous().unsqueeze(1)
        return gxs * g, gys * g, None


def exact_assignment(x, y, kind="sqeuclid", p=2.0, n_power=1.0, return_info=False):
    """Optimal assignment sigma (B,N) int64 between equally sized clouds under the cost ``kind`` (exact LP optimum for
    uniform weights -- what ``ot.emd2`` solves at s2_wasserstein.py:41-43), by the float64 auction kernel."""
    lib = _lib.lib()
    xc, _ = _as_cloud(x, "x")
    yc, _ = _as_cloud(y, "y")
    if xc.shape != yc.shape:
        raise ValueError("exact assignment needs equally sized clouds (uniform weights, n == m)")
    B, N, _ = xc.shape
    if N > lib.shwd_exact_assignment_max_points():
        raise ValueError("exact assignment supports at most %d points per cloud" % lib.shwd_exact_assignment_max_points())
    k = COST_KINDS[kind] if isinstance(kind, str) else int(kind)
    cosine = k in (_lib.COST_GEODESIC, _lib.COST_ONE_MINUS_COS)
    dev = xc.device
    xd, yd = xc.detach(), yc.detach()
    x4 = torch.empty(B, N, 4, device=dev, dtype=torch.float32)
    y4 = torch.empty(B, N, 4, device=dev, dtype=torch.float32)
    sigma = torch.empty(B, N, device=dev, dtype=torch.int32)
    prices = torch.empty(B, N, device=dev, dtype=torch.float64)
    rounds = torch.empty(B, device=dev, dtype=torch.int32)
    status = torch.empty(1, device=dev, dtype=torch.int32)
    flags = MAP_NORMALIZE if cosine else 0
    with torch.cuda.device(dev):
        s = _stream()
        _lib.check(lib.shwd_sphere_map_fwd(_ptr(xd), _ptr(x4), None, B, N, flags, s), "shwd_sphere_map_fwd")
        _lib.check(lib.shwd_sphere_map_fwd(_ptr(yd), _ptr(y4), None, B, N, flags, s), "shwd_sphere_map_fwd")
        _lib.check(lib.shwd_exact_assignment(_ptr(x4), _ptr(y4), B, N, k, float(p), float(n_power), _ptr(sigma), _ptr(prices),
                                             _ptr(rounds), _ptr(status), s), "shwd_exact_assignment")
    sig = sigma.long()
    if return_info:
        return sig, prices, rounds, status
    return sig


def exact_assignment_dense(C, return_info=False):
    """Optimal assignment sigma (B,N) int64 for explicit square cost matrices C (B,N,N) [or (N,N)] float32 with uniform
    weights -- the LP ``ot.emd2(a, b, M)`` solves when called on a cost matrix (main_rotation.py:63-79)."""
    lib = _lib.lib()
    if not isinstance(C, torch.Tensor) or not C.is_cuda:
        raise RuntimeError("the cost matrix must live on a CUDA device: the B200 loss path has no CPU fallback")
    Cb = C.detach()
    if Cb.dim() == 2:
        Cb = Cb.unsqueeze(0)
    if Cb.dim() != 3 or Cb.shape[1] != Cb.shape[2]:
        raise ValueError("exact assignment needs square cost matrices (uniform weights, n == m), got %s" % (tuple(C.shape),))
    Cb = Cb.float().contiguous()
    B, N, _ = Cb.shape
    if N > lib.shwd_exact_assignment_max_points():
        raise ValueError("exact assignment supports at most %d points per cloud" % lib.shwd_exact_assignment_max_points())
    dev = Cb.device
    sigma = torch.empty(B, N, device=dev, dtype=torch.int32)
    prices = torch.empty(B, N, device=dev, dtype=torch.float64)
    rounds = torch.empty(B, device=dev, dtype=torch.int32)
    status = torch.empty(1, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(lib.shwd_exact_assignment_dense(_ptr(Cb), B, N, _ptr(sigma), _ptr(prices), _ptr(rounds), _ptr(status), _stream()),
                   "shwd_exact_assignment_dense")
    sig = sigma.long()
    if return_info:
        return sig, prices, rounds, status
    return sig


def exact_copies(n, m):
    """Uniform marginals 1/n and 1/m with n != m: the transport LP is the assignment problem between L/n copies of every
    source point and L/m copies of every target point, L = lcm(n, m) (each copy carries mass 1/L; any optimal assignment of
    the copies, merged back, is an optimal plan and vice versa).  Returns (L/n, L/m), or None when L exceeds what the
    assignment kernel takes."""
    import math
    L = n * m // math.gcd(n, m)
    if L > _lib.lib().shwd_exact_assignment_max_points():
        return None
    return L // n, L // m


def exact_emd2_dense(M):
    """``ot.emd2(a, b, M)`` for uniform ``a``, ``b`` and a cost matrix M (N,N') [or (B,N,N') -> (B,)]: the value
    (1/n) sum_i M[i, sigma(i)] accumulated in float64 like POT, returned in M's dtype; through autograd on the n matched
    entries the gradient w.r.t. M is the optimal plan -- what POT's torch backend attaches.  N != N': solved on the
    lcm(N, N') copies of ``exact_copies`` (rows / columns repeated; the gradient sums back over the copies)."""
    if M.shape[-2] != M.shape[-1]:
        rc = exact_copies(M.shape[-2], M.shape[-1])
        if rc is None:
            raise NotImplementedError("exact emd2 between %d and %d points needs lcm = %d copies; the assignment kernel takes %d"
                                      % (M.shape[-2], M.shape[-1], M.shape[-2] * M.shape[-1] // __import__("math").gcd(M.shape[-2], M.shape[-1]),
                                         _lib.lib().shwd_exact_assignment_max_points()))
        return exact_emd2_dense(M.repeat_interleave(rc[0], dim=-2).repeat_interleave(rc[1], dim=-1))
    sigma, _, _, status = exact_assignment_dense(M, return_info=True)
    Mb = M if M.dim() == 3 else M.unsqueeze(0)
    c = torch.gather(Mb, 2, sigma.unsqueeze(-1)).squeeze(-1)  # (B,N): M[b, i, sigma(i)]
    v = (c.double().sum(dim=1) / Mb.shape[1]).to(M.dtype)
    v = torch.where(status != 0, torch.full_like(v, float("nan")), v)  # failed solve (non-finite costs): never a number
    return v if M.dim() == 3 else v.reshape(())


def _pair_cost(x, ys, kind, p, n_power=1.0):
    """C(x_i, ys_i) for matched points, with the reference's own torch formulas (s2_wasserstein.py:52-63,112-123;
    max_spherical_w_cos_with_regulation.py:745) so the value and its autograd are the reference's on those n entries."""
    if kind == "geodesic":
        c = torch.acos(torch.nn.functional.cosine_similarity(x, ys, dim=-1)) ** p
    elif kind == "one_minus_cos":
        c = (1 - torch.nn.functional.cosine_similarity(x, ys, dim=-1)) ** p
    elif kind == "sqeuclid":
        c = torch.sum(torch.abs(x - ys) ** p, -1)
    elif kind == "euclid":
        c = torch.pow(torch.sum(torch.abs(x - ys) ** p, -1), 1.0 / p)
    else:
        raise ValueError(kind)
    return c if n_power == 1.0 else torch.pow(c, n_power)


def exact_emd2(x, y, kind="sqeuclid", p=2.0):
    """Per-pair exact ``emd2`` (B,) with uniform weights: the optimal permutation from the auction kernel, the value
    (1/n) sum_i C(x_i, y_sigma(i)) accumulated in float64 like POT does, and -- through autograd on the n matched
    entries -- exactly the gradient POT attaches (d emd2 / dC = optimal plan)."""
    xc, _ = _as_cloud(x, "x")
    yc, _ = _as_cloud(y, "y")
    if xc.shape[1] != yc.shape[1]:  # n != m: the assignment of lcm(n, m) copies (exact_copies); gradients sum over the copies
        rc = exact_copies(xc.shape[1], yc.shape[1])
        if rc is None:
            raise ValueError("exact emd2 between %d and %d points: lcm(n, m) exceeds the %d points the assignment kernel takes"
                             % (xc.shape[1], yc.shape[1], _lib.lib().shwd_exact_assignment_max_points()))
        xc = xc.repeat_interleave(rc[0], dim=1)
        yc = yc.repeat_interleave(rc[1], dim=1)
    sigma, _, _, status = exact_assignment(xc, yc, kind, p, return_info=True)
    ys = torch.gather(yc, 1, sigma.unsqueeze(-1).expand(-1, -1, 3))
    c = _pair_cost(xc, ys, kind, p)
    v = (c.double().sum(dim=1) / xc.shape[1]).to(c.dtype)
    return torch.where(status != 0, torch.full_like(v, float("nan")), v)  # failed solve (non-finite costs): never a number


class ResidualFlowStackFn(torch.autograd.Function):
    """y = phi(x) for a stack of Residual flows, one fused launch per direction (csrc/resflow.cu).
    ``apply(x, uv, n_layers, coeff, *params)``: ``params`` are the module's raw parameter tensors in the layout order of
    include/shwd.h (per flow: W0, b0, ..., W6, b6, beta0..beta6); they are concatenated once, and the flat raw-parameter
    gradient the kernel returns is handed back as views -- no eager parameter preparation on either side."""

    @staticmethod
    def forward(ctx, x, uv, n_layers, coeff, *params):
        lib = _lib.lib()
        if not x.is_cuda:
            raise RuntimeError("phi inputs must live on a CUDA device: no CPU fallback")
        xc = x.contiguous().float()
        flat = torch.cat([p.reshape(-1) for p in params]).float()
        if flat.numel() != n_layers * lib.shwd_resflow_params_per_layer():
            raise ValueError("parameter list does not match the fused Residual-flow layout")
        npts = xc.numel() // 3
        y = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(lib.shwd_resflow_fwd(_ptr(xc), npts, _ptr(flat), _ptr(uv), int(n_layers), float(coeff), _ptr(y), _stream()),
                       "shwd_resflow_fwd")
        ctx.save_for_backward(xc, flat, uv)
        ctx.n_layers, ctx.coeff = int(n_layers), float(coeff)
        ctx.shapes = [p.shape for p in params]
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, flat, uv = ctx.saved_tensors
        lib = _lib.lib()
        gy = gy.contiguous().float()
        npts = xc.numel() // 3
        gx = torch.empty_like(xc)
        gp = torch.empty_like(flat)
        wsb = lib.shwd_resflow_workspace_bytes(npts, ctx.n_layers)
        ws = torch.empty(max(wsb, 8), device=xc.device, dtype=torch.uint8)
        with torch.cuda.device(xc.device):
            _lib.check(lib.shwd_resflow_bwd(_ptr(xc), _ptr(gy), npts, _ptr(flat), _ptr(uv), ctx.n_layers, ctx.coeff, _ptr(gx),
                                            _ptr(gp), _ptr(ws), wsb, _stream()), "shwd_resflow_bwd")
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for d in shp:
                n *= d
            grads.append(gp[off:off + n].view(shp))
            off += n
        return (gx, None, None, None) + tuple(grads)


class PlanarFlowStackFn(torch.autograd.Function):
    """y = phi(x) for a stack of Planar flows, one fused launch per direction (csrc/planar.cu).
    ``apply(x, n_layers, *params)``: ``params`` are the modules' raw tensors, per flow u (1,3), w (1,3), b (1,).  ``x`` is
    (N,3) -- lin = <w, z_n> -- or (B,N,3) -- lin = w_c sum_n z_bnc, what the reference's sum over dim 1 yields for batched
    clouds (flows/planar.py:49-60 as called from s2_wasserstein.py:160-163)."""

    @staticmethod
    def forward(ctx, x, n_layers, *params):
        lib = _lib.lib()
        if not x.is_cuda:
            raise RuntimeError("phi inputs must live on a CUDA device: no CPU fallback")
        if x.dim() not in (2, 3) or x.shape[-1] != 3:
            raise ValueError("fused Planar stack: x must be (N,3) or (B,N,3)")
        xc = x.contiguous().float()
        flat = torch.cat([p.reshape(-1) for p in params]).float()
        if flat.numel() != n_layers * lib.shwd_planar_params_per_layer() or not 0 < n_layers <= lib.shwd_planar_max_layers():
            raise ValueError("parameter list does not match the fused Planar-flow layout")
        clouds, npts = (xc.shape[0], xc.shape[1]) if xc.dim() == 3 else (0, xc.shape[0])
        y = torch.empty_like(xc)
        colsum = torch.empty((max(clouds, 1), 3), device=xc.device, dtype=torch.float64)
        if xc.numel():
            with torch.cuda.device(xc.device):
                _lib.check(lib.shwd_planar_fwd(_ptr(xc), clouds, npts, _ptr(flat), int(n_layers), _ptr(y), _ptr(colsum), _stream()),
                           "shwd_planar_fwd")
        ctx.save_for_backward(xc, flat, colsum)
        ctx.n_layers, ctx.clouds, ctx.npts = int(n_layers), clouds, npts
        ctx.shapes = [p.shape for p in params]
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, flat, colsum = ctx.saved_tensors
        lib = _lib.lib()
        gy = gy.contiguous().float()
        gx = torch.empty_like(xc)
        gp = torch.zeros_like(flat)
        if xc.numel():
            wsb = lib.shwd_planar_workspace_bytes(ctx.clouds, ctx.npts, ctx.n_layers)
            ws = torch.empty(max(wsb, 8), device=xc.device, dtype=torch.uint8)
            with torch.cuda.device(xc.device):
                _lib.check(lib.shwd_planar_bwd(_ptr(xc), _ptr(gy), _ptr(colsum), ctx.clouds, ctx.npts, _ptr(flat), ctx.n_layers,
                                               _ptr(gx), _ptr(gp), _ptr(ws), wsb, _stream()), "shwd_planar_bwd")
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for d in shp:
                n *= d
            grads.append(gp[off:off + n].view(shp))
            off += n
        return (gx, None) + tuple(grads)


def _sort_i32(keys2d):
    """(S,len) float32 keys -> (sorted values, int32 stable-sort permutation); device-side only."""
    lib = _lib.lib()
    S, length = keys2d.shape
    out = torch.empty_like(keys2d)
    perm = torch.empty(keys2d.shape, device=keys2d.device, dtype=torch.int32)
    wsb = lib.shwd_segmented_sort_workspace_bytes(S, length)
    ws = torch.empty(max(wsb, 8), device=keys2d.device, dtype=torch.uint8)
    _lib.check(lib.shwd_segmented_sort_i32(_ptr(keys2d), S, length, _ptr(out), _ptr(perm), _ptr(ws), wsb, _stream()),
               "shwd_segmented_sort_i32")
    return out, perm


class SlicedLossFn(torch.autograd.Function):
    """The whole sliced path of one call in five launches forward and two backward:
        project x, project y -> sort, sort (int32 permutations) -> one 1-D reduction kernel that also writes
        d w_slice / d(unsorted keys) through the permutations;   backward: the two projection backward kernels.
    mode "circle_w1":  sliced_cost p == 1  (max_spherical_sliced_w.py:251-286 with emd1D_circle :210-247)
    mode "circle_wp":  sliced_cost p != 1  (binary_search_circle :117-207), mean of W_p^p
    mode "line":       Euclidean sliced W  (Flow_ellipsoid.ipynb:208-220), returns mean_P sum_n |.|^p (root taken outside)
    Returns the per-pair mean over the slices (B,).  Compared with chaining ProjectCircleFn / SegmentedSortFn /
    CircularW1Fn it never writes the sorted-order gradients, needs no unsort launches, no gradient scaling passes and
    no int64 permutations (cfg3: 2 x 8 B x P x N per pair)."""

    @staticmethod
    def forward(ctx, x, y, frames, mode, p, tm, tp, tol):
        lib = _lib.lib()
        x = x.contiguous()
        y = y.contiguous()
        frames = frames.contiguous()
        B, n, _ = x.shape
        m = y.shape[1]
        # frames (P,3,2) / (P,3): one set for every pair; (B,P,3,2): one set PER PAIR (max_spherical_sliced_w_fast.py:298-319)
        pp = frames.dim() == 4
        if pp and (mode == "line" or frames.shape[0] != B):
            raise ValueError("per-pair frames must be (B,P,3,2) and are for the circle modes")
        P = frames.shape[1] if pp else frames.shape[0]
        sort_projected = lib.shwd_sort_projected_pp if pp else lib.shwd_sort_projected
        project_circle = lib.shwd_project_circle_pp if pp else lib.shwd_project_circle
        S = B * P
        dev = x.device
        ku = torch.empty(S, n, device=dev, dtype=torch.float32)
        kv = torch.empty(S, m, device=dev, dtype=torch.float32)
        w = torch.empty(S, device=dev, dtype=torch.float32)
        fused_max = lib.shwd_sort_projected_max_points()
        with torch.cuda.device(dev):
            st = _stream()
            sorted_perm = []
            for c, cnt, kbuf in ((x, n, ku), (y, m, kv)):
                if cnt <= fused_max:  # the sort CTAs compute their own keys: no key array, no projection launch
                    so = torch.empty(S, cnt, device=dev, dtype=torch.float32)
                    pe = torch.empty(S, cnt, device=dev, dtype=torch.int32)
                    _lib.check(sort_projected(_ptr(c), _ptr(frames), B, cnt, P, 2 if mode == "line" else 1, _ptr(so), _ptr(pe), st),
                               "shwd_sort_projected")
                    sorted_perm.append((so, pe))
                else:
                    if mode == "line":
                        _lib.check(lib.shwd_project_line(_ptr(c), _ptr(frames), B, cnt, P, _ptr(kbuf), st), "shwd_project_line")
                    else:
                        _lib.check(project_circle(_ptr(c), _ptr(frames), B, cnt, P, _ptr(kbuf), st), "shwd_project_circle")
                    sorted_perm.append(_sort_i32(kbuf))
            (su, pu), (sv, pv) = sorted_perm
            gku, gkv = ku, kv  # (dead once sorted / never filled): these buffers receive d w / d keys
            if mode == "circle_w1":
                _lib.check(lib.shwd_circular_w1_scatter(_ptr(su), _ptr(sv), _ptr(pu), _ptr(pv), S, n, m, _ptr(w), _ptr(gku), _ptr(gkv),
                                                        st), "shwd_circular_w1_scatter")
            elif mode == "circle_wp":
                wsb = lib.shwd_circular_wp_workspace_bytes(S, n, m)
                ws = torch.empty(max(wsb, 8), device=dev, dtype=torch.uint8)
                _lib.check(lib.shwd_circular_wp_scatter(_ptr(su), _ptr(sv), _ptr(pu), _ptr(pv), S, n, m, float(p), float(tm), float(tp),
                                                        float(tol), _ptr(w), _ptr(gku), _ptr(gkv), None, _ptr(ws), wsb, st),
                           "shwd_circular_wp_scatter")
            elif mode == "line":
                _lib.check(lib.shwd_euclid_sw_scatter(_ptr(su), _ptr(sv), _ptr(pu), _ptr(pv), S, n, float(p), _ptr(w), _ptr(gku),
                                                      _ptr(gkv), st), "shwd_euclid_sw_scatter")
            else:
                raise ValueError("unknown sliced mode %r" % (mode,))
        ctx.save_for_backward(x, y, frames, gku, gkv)
        ctx.mode, ctx.pp = mode, pp
        return w.view(B, P).mean(dim=1)

    @staticmethod
    def backward(ctx, gw):
        x, y, frames, gku, gkv = ctx.saved_tensors
        lib = _lib.lib()
        B, n, _ = x.shape
        m = y.shape[1]
        P = frames.shape[1] if ctx.pp else frames.shape[0]
        circle_bwd = lib.shwd_project_circle_bwd_scaled_pp if ctx.pp else lib.shwd_project_circle_bwd_scaled
        gwc = gw.reshape(B).contiguous().float()  # the kernels scale by gw[b] / P themselves
        gx = gy = None
        with torch.cuda.device(x.device):
            st = _stream()
            for idx, (c, gk, cnt) in enumerate(((x, gku, n), (y, gkv, m))):
                if not ctx.needs_input_grad[idx]:
                    continue
                g = torch.empty_like(c)
                if ctx.mode == "line":
                    _lib.check(lib.shwd_project_line_bwd_scaled(_ptr(frames), B, cnt, P, _ptr(gk), _ptr(gwc), _ptr(g), st),
                               "shwd_project_line_bwd_scaled")
                else:
                    _lib.check(circle_bwd(_ptr(c), _ptr(frames), B, cnt, P, _ptr(gk), _ptr(gwc), _ptr(g), st),
                               "shwd_project_circle_bwd_scaled")
                if idx == 0:
                    gx = g
                else:
                    gy = g
        return gx, gy, None, None, None, None, None, None


CIRCULAR_W1_MAX = 32768  # n + m the circular_w1 kernel takes per slice (512 threads x up to 64 merged entries each)


def circular_w1_large(us, vs, uw=None, vw=None):
    """emd1D_circle (max_spherical_sliced_w.py:230-247) on sorted rows us (S,n), vs (S,m) of ANY length and with ANY
    weights: the reference's own four-sort formulation, with every sort done by the segmented radix sort kernel and the
    scans / gathers by torch device ops.  Used for slices beyond CIRCULAR_W1_MAX and whenever weights are given (uw (S,n),
    vw (S,m): the weights of the SORTED entries, :224-228; the fused level-median kernel evaluates uniform CDFs in closed
    form).  Differentiable w.r.t. us / vs through the merged sort's permutation, like autograd through torch.sort, and
    w.r.t. the weights through the cumulative sum."""
    S, n = us.shape
    m = vs.shape[1]
    merged, mperm = SegmentedSortFn.apply(torch.cat((us, vs), -1))  # stable: a u entry precedes an equal v entry
    if uw is None and vw is None:
        wts = torch.cat((torch.full((n,), 1 / n, dtype=torch.float32, device=us.device),
                         -torch.full((m,), 1 / m, dtype=torch.float32, device=us.device)))
        cdf_diff = torch.cumsum(wts[mperm], -1)
    else:
        uw = torch.full((S, n), 1 / n, dtype=torch.float32, device=us.device) if uw is None else uw.to(us.dtype).expand(S, n)
        vw = torch.full((S, m), 1 / m, dtype=torch.float32, device=us.device) if vw is None else vw.to(us.dtype).expand(S, m)
        cdf_diff = torch.cumsum(torch.gather(torch.cat((uw, -vw), -1), -1, mperm), -1)
    cdf_sorted, cperm = segmented_sort_raw(cdf_diff)
    delta = torch.cat((merged[:, 1:], torch.ones_like(merged[:, :1])), -1) - merged  # the arc [0, first) is omitted (:238-239)
    cw = torch.cumsum(torch.gather(delta.detach(), -1, cperm), -1) - 0.5
    k = torch.argmin(torch.where(cw < 0, torch.full_like(cw, float("inf")), cw), dim=-1, keepdim=True)
    lev_med = torch.gather(cdf_diff, -1, torch.gather(cperm.long(), -1, k))  # = cdf_sorted[k], differentiable (weights)
    return torch.sum(delta * torch.abs(cdf_diff - lev_med), dim=-1)


def spherical_sliced_w1(Xs, Xt, U):
    """mean_P circular-W1 of the great-circle projections (sliced_cost with p == 1, explicit frames U (P,3,2), or one set per
    pair (B,P,3,2))."""
    xs, _ = _as_cloud(Xs, "Xs")
    xt, _ = _as_cloud(Xt, "Xt")
    U = U.to(device=xs.device, dtype=torch.float32)
    if xs.shape[1] + xt.shape[1] <= CIRCULAR_W1_MAX:
        return SlicedLossFn.apply(xs, xt, U, "circle_w1", 1.0, 0.0, 0.0, 0.0)  # (B,)
    if U.dim() == 4:
        raise ValueError("per-pair frames (B,P,3,2) are supported up to n + m = %d points per pair" % CIRCULAR_W1_MAX)
    ks = ProjectCircleFn.apply(xs, U)  # (B,P,n)
    kt = ProjectCircleFn.apply(xt, U)
    B, P, n = ks.shape
    ss, _ = SegmentedSortFn.apply(ks.reshape(B * P, n))
    st, _ = SegmentedSortFn.apply(kt.reshape(B * P, kt.shape[2]))
    return circular_w1_large(ss, st).reshape(B, P).mean(dim=1)


def spherical_sliced_wp(Xs, Xt, U, p=2.0, tm=-1.0, tp=1.0, tol=1e-7):
    """mean_P circular-W_p^p (no root, max_spherical_sliced_w.py:284-286) of the great-circle projections, p != 1."""
    xs, _ = _as_cloud(Xs, "Xs")
    xt, _ = _as_cloud(Xt, "Xt")
    U = U.to(device=xs.device, dtype=torch.float32)
    return SlicedLossFn.apply(xs, xt, U, "circle_wp", float(p), tm, tp, tol)  # (B,)


def euclid_sliced_w(x, y, theta, p=2.0):
    """(mean_P sum_n |sort(x theta) - sort(y theta)|^p)^(1/p) per pair (Flow_ellipsoid.ipynb:208-220)."""
    xc, _ = _as_cloud(x, "x")
    yc, _ = _as_cloud(y, "y")
    if xc.shape[1] != yc.shape[1]:
        raise ValueError("Euclidean sliced W needs equally sized clouds")
    theta = theta.to(device=xc.device, dtype=torch.float32)
    return SlicedLossFn.apply(xc, yc, theta, "line", float(p), 0.0, 0.0, 0.0).pow(1.0 / p)

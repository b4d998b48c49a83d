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
        keep = bool(need_grad) or thresh > 0
        HL = iters + 1 if keep else 1
        alpha = torch.empty(B, HL, N, **f32)
        beta = torch.empty(B, HL, M, **f32)
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
        ctx.last_ws = ws
        return gx, gy, None, None, None, None, None, None, None, None


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
        lvl = self.iterations() if self._alpha.shape[1] > 1 else 0
        inv_k = eps / 1.4426950408889634
        return self._alpha[:, lvl] * inv_k, self._beta[:, lvl] * inv_k

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
        HL = self._alpha.shape[1]
        lvl = self.iterations() if HL > 1 else 0
        a = self._alpha[:, lvl]
        b = self._beta[:, lvl]
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

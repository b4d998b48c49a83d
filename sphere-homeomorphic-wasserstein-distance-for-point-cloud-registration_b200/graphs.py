"""CUDA-graph replay of a whole loss call (forward AND backward) for the launch-bound end of the path.

At the reference's training shapes several loss calls are a few tens of microseconds of kernels inside a few hundred
microseconds of host work (Chamfer at B=32, N=1024 -- train_CD.py:161; one sliced pair -- Flow_ellipsoid.ipynb cell 8):
every call walks Python wrappers, the autograd graph and 5-20 launches.  None of the C-ABI launchers synchronises,
allocates or touches the host, so the complete ``loss = fn(*clouds); loss.backward()`` sequence is capturable;
:func:`graphed_loss` captures it once per input signature and afterwards a call is: copy the inputs into the captured
buffers, ONE ``cudaGraphLaunch``, hand out the loss and the gradients the graph has already computed.

The loss must be a scalar and a deterministic function of its tensor arguments (draw random frames outside: pass ``U`` to
``sliced_cost`` instead of calling ``sliced_wasserstein_sphere``, whose ``torch.linalg.qr`` is not capturable).  The
gradient of a scalar loss w.r.t. its inputs is computed inside the same graph, so the autograd node returned to the
caller only scales it by the incoming gradient -- double backward is not supported.
"""
import torch

__all__ = ["graphed_loss", "GraphedLoss"]


def _needs_grad(t):
    return bool(isinstance(t, torch.Tensor) and t.requires_grad and torch.is_grad_enabled())


def _signature(tensors):
    return tuple((tuple(t.shape), t.dtype, t.device.index, _needs_grad(t)) if isinstance(t, torch.Tensor) else ("?",) for t in tensors)


class _Capture:
    """One captured fwd+bwd for one input signature."""

    def __init__(self, fn, tensors, warmup):
        for t in tensors:
            if not (isinstance(t, torch.Tensor) and t.is_cuda):
                raise RuntimeError("graphed_loss: every argument must be a CUDA tensor (no CPU fallback)")
        self.needs = [_needs_grad(t) for t in tensors]
        self.static_in = [t.detach().clone().contiguous() for t in tensors]

        @torch.enable_grad()
        def run():
            ins = [s.detach().requires_grad_(n) for s, n in zip(self.static_in, self.needs)]
            loss = fn(*ins)
            if loss.numel() != 1:
                raise ValueError("graphed_loss: the wrapped function must return a scalar loss, got shape %s" % (tuple(loss.shape),))
            wrt = [i for i, n in zip(ins, self.needs) if n]
            grads = torch.autograd.grad(loss, wrt, allow_unused=True) if wrt else ()
            return loss.detach(), grads

        dev = self.static_in[0].device
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up off the capture: function attributes, lazy module state, allocator
                for _ in range(max(1, warmup)):
                    run()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                loss, grads = run()
        self.loss = loss
        it = iter(grads)
        self.grads = [next(it) if n else None for n in self.needs]

    def replay(self, tensors):
        for s, t in zip(self.static_in, tensors):
            s.copy_(t.detach())
        self.graph.replay()


class _ReplayFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cap, *tensors):
        cap.replay(tensors)
        # the captured buffers are overwritten by the next replay: hand out copies
        ctx.grads = [None if g is None else g.clone() for g in cap.grads]
        return cap.loss.clone()

    @staticmethod
    def backward(ctx, g):
        return (None,) + tuple(None if gr is None else gr * g for gr in ctx.grads)


class GraphedLoss:
    """``GraphedLoss(fn)(*clouds)`` == ``fn(*clouds)`` for a scalar loss of CUDA tensors, replayed from a CUDA graph that
    holds the forward and the backward.  One capture per (shapes, dtypes, device, requires_grad) signature."""

    def __init__(self, fn, warmup=2, max_captures=16):
        self.fn, self.warmup, self.max_captures = fn, warmup, max_captures
        self._captures = {}

    def _capture_for(self, tensors):
        key = _signature(tensors)
        cap = self._captures.get(key)
        if cap is None:
            if len(self._captures) >= self.max_captures:
                self._captures.pop(next(iter(self._captures)))
            cap = self._captures[key] = _Capture(self.fn, tensors, self.warmup)
        return cap

    def __call__(self, *tensors):
        """Loss as an autograd-connected 0-dim tensor (gradients flow to the arguments that require them)."""
        cap = self._capture_for(tensors)
        if torch.is_grad_enabled() and any(cap.needs):
            return _ReplayFn.apply(cap, *tensors)
        cap.replay(tensors)
        return cap.loss.clone()

    def value_and_grad(self, *tensors):
        """(loss, [grad or None per argument]) as views of the captured buffers -- valid until the next call; the
        zero-overhead form for gradient-flow loops (Flow_ellipsoid.ipynb cell 8: ``x -= lr * grad``)."""
        cap = self._capture_for(tensors)
        cap.replay(tensors)
        return cap.loss, cap.grads


def graphed_loss(fn, warmup=2, max_captures=16):
    """Wrap a scalar loss ``fn(*cuda_tensors)`` so that each call replays one CUDA graph of its forward + backward."""
    return GraphedLoss(fn, warmup=warmup, max_captures=max_captures)

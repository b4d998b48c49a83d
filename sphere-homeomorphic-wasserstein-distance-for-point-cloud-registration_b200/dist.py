"""Multi-GPU plumbing: the loss path shards by independent units -- pairs of the batch, or slices of a
single-pair sliced run -- one process per GPU, no collective on the data path.  The only collectives are the scalar
loss all-reduce of a data-parallel step (``mean of equal-sized local means == the reference's emd / B``,
s2_wasserstein.py:44) and, for slice sharding, the all-reduce of the per-rank partial sums / gradients.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world_size):
    """Contiguous [begin, end) of ``n_units`` independent units owned by ``rank`` (sizes differ by at most one)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("invalid rank/world_size %r/%r" % (rank, world_size))
    base, rem = divmod(int(n_units), world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def global_mean(local_sum, local_count, group=None, grad_reduction="sum"):
    """Mean over all ranks of per-unit values given this rank's sum and count: one all-reduce of a 2-vector.
    Differentiable w.r.t. ``local_sum``: the gradient w.r.t. a local unit is 1 / global_count, so the gradient of the
    global mean w.r.t. shared parameters is the SUM of the ranks' gradients -- ``grad_reduction="sum"`` (default): the
    caller all-reduces parameter gradients with ReduceOp.SUM.  ``torch.nn.parallel.DistributedDataParallel`` AVERAGES
    them instead: pass ``grad_reduction="mean"`` there, which scales this rank's differentiable part by the world size so
    that DDP's average is the single-process gradient (the returned value is the global mean either way)."""
    if grad_reduction not in ("sum", "mean"):
        raise ValueError("grad_reduction must be 'sum' or 'mean'")
    count = torch.as_tensor(float(local_count), dtype=local_sum.dtype, device=local_sum.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        tot = torch.stack([local_sum.detach(), count])
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
        scale = float(dist.get_world_size(group)) if grad_reduction == "mean" else 1.0
        # value: global; gradient: flows through the local contribution only
        return ((local_sum - local_sum.detach()) * scale + tot[0]) / tot[1]
    return local_sum / count


def sharded_pair_loss(loss_fn, x, y, rank=None, world_size=None, group=None, grad_reduction="sum"):
    """Evaluate ``loss_fn(x_shard, y_shard) -> per-pair losses (b,)`` on this rank's slice of the batch and return the
    global batch mean (what the reference's single-process ``emd / B`` returns).  ``grad_reduction``: see global_mean
    ("mean" inside a DistributedDataParallel training loop, whose gradient all-reduce averages)."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    b0, b1 = shard_range(x.shape[0], rank, world_size)
    if b1 > b0:
        per_pair = loss_fn(x[b0:b1], y[b0:b1])
        s = per_pair.sum()
    else:
        s = x.new_zeros(())
    return global_mean(s, b1 - b0, group, grad_reduction)


def sharded_slice_loss(slice_loss_fn, Xs, Xt, frames, rank=None, world_size=None, group=None):
    """Slice sharding of one sliced-loss call (cfg3 / cfg4 style single-pair runs): every rank holds the same
    clouds and the same frames ``(P, ...)``, evaluates ``slice_loss_fn(Xs, Xt, frames[p0:p1]) -> mean over its slices``
    (e.g. ``lambda a, b, U: losses.sliced_cost(a, b, U, p=2)``) and the global mean over all P slices is returned (the
    reference's ``torch.mean`` over slices, max_spherical_sliced_w.py:284-286).  The value is global on every rank; the
    gradient w.r.t. the clouds flows through the local slices only, so a caller that differentiates w.r.t. the clouds sums
    the ranks' gradients (``dist.all_reduce(x.grad)`` -- what DDP does for parameters)."""
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    p0, p1 = shard_range(frames.shape[0], rank, world_size)
    if p1 > p0:
        s = slice_loss_fn(Xs, Xt, frames[p0:p1]) * float(p1 - p0)  # local mean -> local sum
    else:
        s = Xs.new_zeros(())
    return global_mean(s.reshape(()) if s.dim() == 0 else s, p1 - p0, group)

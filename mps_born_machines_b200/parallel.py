"""Sample sharding across ranks (one process per GPU).  Environments, psi_n and gradient contributions are
independent per sample; the only exchange of the hot path is ONE all-reduce (sum) per two-site step of the
buffer  [ S = sum_n psi'_n/psi_n  |  sum_n ln|psi_n|  |  #zero psi ]  (2 MiB + 16 B at D=256), followed by a
replicated update + split (inputs are bitwise identical on every rank after the all-reduce)."""
import torch
import torch.distributed as dist


def shard_bounds(n_total, world, rank):
    """contiguous blocks of ceil(n/world) samples; the last ranks may hold fewer (or zero) samples."""
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def allreduce_sum_(buf, group=None):
    """in-place sum over ranks of the gradient buffer (NCCL on GPU tensors, gloo on CPU tensors)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def replicated_checksum(t, group=None):
    """max |x_rank - x_0| of a scalar fingerprint: guards the assumption that replicated state stays identical."""
    fp = t.double().sum().reshape(1).clone()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0.0
    ref = fp.clone()
    dist.broadcast(ref, src=0, group=group)
    diff = (fp - ref).abs()
    dist.all_reduce(diff, op=dist.ReduceOp.MAX, group=group)
    return float(diff.item())

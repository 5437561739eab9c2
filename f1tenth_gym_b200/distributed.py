"""Multi-GPU launcher helpers.  Environments never interact (only the agents inside one env do), so
the N envs shard embarrassingly: rank g owns global env ids [g*N/G, (g+1)*N/G).  There
is NO collective on the step path.  `all_gather_obs` is the optional observation concatenation for a
single-process trainer (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(num_envs_global, rank, world_size):
    """Contiguous env-id range of `rank`; sizes differ by at most one."""
    lo = (num_envs_global * rank) // world_size
    hi = (num_envs_global * (rank + 1)) // world_size
    return lo, hi


def env_seeds(base_seed, lo, hi):
    """Per-env RNG stream keys derived from the GLOBAL env id, so results do not depend on the GPU count."""
    return [base_seed + e for e in range(lo, hi)]


def all_gather_obs(local, group=None, sizes=None, out=None):
    """Concatenate per-rank observation shards along dim 0 on every rank.  Shards may have different
    leading sizes (uneven env split): sizes are exchanged first, shards padded to the max.
    sizes: the leading size of every rank's shard if the caller already knows them (skips the size exchange and its
    host sync).  When all shards are equal the gather writes straight into one output tensor (`out`, optional,
    [sum(sizes), ...]) with a single NCCL all_gather over NVLink -- no padding, no concatenation copy."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    if sizes is None:
        n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        got = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(got, n_local, group=group)
        sizes = [int(s.item()) for s in got]
    sizes = [int(x) for x in sizes]
    n_max = max(sizes)
    if min(sizes) == n_max and hasattr(dist, 'all_gather_into_tensor'):
        if out is None:
            out = torch.empty((n_max * world,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = local
    if local.shape[0] < n_max:
        pad = torch.cat([local, local.new_zeros((n_max - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def reduce_max_scalar(value, device, group=None):
    """max over ranks of a python float (bench timing: the slowest rank defines the step time)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def reduce_sum_scalar(value, device, group=None):
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())

"""Multi-GPU plumbing: the hot path shards by the independent samples of a batch — every sample is a
disconnected component of the batched graph (models/mpnn_2d.py:245, models/magnet_gnn.py:247,293), InstanceNorm is
per graph, LayerNorm per row — so there is NO data-path collective.  Training adds ONE sum all-reduce per step over a
flat gradient buffer (what Lightning's DDP does in 25 MB buckets; MAgNetGNN's 2.57 M parameters fit one bucket).

One process per GPU; `torch.distributed` (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests)."""
from typing import Dict, Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of the batch dimension; the first `n_samples % world` ranks get one extra sample."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Rank-local slice of a reference batch dict ({'u','x','t'} of datamodule/dataset_2d.py:54-58 or
    {'t','lr_frames','hr_points','coords_hr','coords_lr'} of :101-137): every tensor is split along dim 0."""
    sizes = {v.shape[0] for v in batch.values() if torch.is_tensor(v)}
    if len(sizes) != 1:
        raise ValueError(f"batch tensors disagree on the batch dimension: {sorted(sizes)}")
    lo, hi = shard_range(sizes.pop(), rank, world)
    return {k: (v[lo:hi] if torch.is_tensor(v) else v) for k, v in batch.items()}


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: int = None, average: bool = True) -> int:
    """Sum (and average) the gradients of `params` across ranks through one flat buffer and one collective.
    Parameters without a gradient contribute zeros (every rank must walk the same parameter list).
    Returns the number of elements reduced."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    world = world if world is not None else dist.get_world_size()
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(world)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off


def gather_samples(local: torch.Tensor, sizes: List[int]) -> torch.Tensor:
    """All-gather rank-local per-sample results (dim 0 = samples; `sizes[r]` samples on rank r) — for evaluation
    bookkeeping only, never on the data path."""
    world = dist.get_world_size()
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)

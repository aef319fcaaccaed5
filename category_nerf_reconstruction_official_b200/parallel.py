"""Multi-GPU layout of the hot path; one process per GPU, torch.distributed for the plumbing.

1. Category sharding (default): category models are independent (separate MLP / PE / code tables; the step loss is
   a plain sum over categories, loss.py:70), so each rank owns a subset of categories and the training step needs
   NO collective.  ``shard_categories`` balances ranks by cost (rays x samples).
2. Ray split of one large category: every rank holds the same model and takes R/G of the rays.  Globals that must
   match the single-GPU result are all-reduced BEFORE the fused step -- the three loss-mask counts (sum) and the
   per-category maximum depth (max) -- and the flat gradient buffers (MLP + PE + code tables, ~80 k floats) are
   summed with ONE all-reduce after it.  The code regulariser is added on rank 0 only, so the sum is exact.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_categories(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of categories to ranks. costs[k] ~ rays_k * samples_k * flop/pt."""
    order = sorted(range(len(costs)), key=lambda k: -costs[k])
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for k in order:
        r = min(range(world_size), key=lambda i: (load[i], i))
        out[r].append(k)
        load[r] += costs[k]
    for lst in out:
        lst.sort()
    return out


def split_rays(n_rays: int, world_size: int, rank: int) -> slice:
    """Contiguous, near-even slice of a category's ray batch for this rank."""
    base, rem = divmod(n_rays, world_size)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def flatten_tensors(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in tensors])


def unflatten_into(flat: torch.Tensor, tensors: Sequence[torch.Tensor]) -> None:
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def allreduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """Loss-mask counts int32 [K,4]: global denominators N_d, N_o, N_s (render_rays.py:86) for a ray-split batch."""
    c = counts.clone()
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return c


def allreduce_max_depth(depth: torch.Tensor, group=None) -> torch.Tensor:
    """Per-category max depth [K] over all ranks' rays (upper bound of the invalid-depth bins, scene_cateogries.py:486)."""
    m = depth.max(dim=-1).values.contiguous()
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return m


def allreduce_gradients(grad_tensors: Sequence[torch.Tensor], group=None, flat: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum the gradient tensors over ranks with ONE collective on a flat fp32 buffer, writing the result back in place."""
    buf = flatten_tensors(grad_tensors) if flat is None else flat
    if flat is not None:
        off = 0
        for t in grad_tensors:
            buf[off:off + t.numel()].copy_(t.reshape(-1))
            off += t.numel()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    unflatten_into(buf, grad_tensors)
    return buf


class RaySplitStep:
    """Ray-split training step of a FusedCategoryTrainer replica: this rank's slice of every category's ray batch."""

    def __init__(self, trainer, group=None):
        self.tr = trainer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._flat = None

    def step(self, rgbs, depth, origins, dirs, indices, rand: Optional[Dict[str, torch.Tensor]] = None):
        """Inputs are THIS rank's rays ([K, R_local, ...]). Returns the global loss (identical on all ranks)."""
        from . import ops, sampling
        tr = self.tr
        if rand is None:
            rand = sampling.draw_randoms(rgbs, depth, tr.N, tr.M, tr.eps, tr.min_bound, tr.this_obj)
        max_depth = allreduce_max_depth(depth, self.group)
        z, _, valid, labels, stats = ops.sample_batched(rgbs, depth, origins, dirs, tr.N, tr.M, tr.eps, tr.stop_eps, rand,
                                                       tr.min_bound, tr.this_obj, want_pcs=False, max_depth=max_depth)
        counts = allreduce_counts(stats[:, 4:8].contiguous(), self.group)
        tr.step_from_samples(origins, dirs, z, depth, rgbs, valid.view(torch.uint8), indices, counts, add_reg=(self.rank == 0))
        grads = [tr.g_params, tr.g_B, tr.g_shape, tr.g_tex, tr.out_loss]
        if self._flat is None:
            self._flat = torch.empty(sum(g.numel() for g in grads), device=tr.g_params.device, dtype=torch.float32)
        # loss terms are sums over local rays / global counts -> additive too; flags (out_loss[1:4]) are re-derived
        allreduce_gradients(grads, self.group, self._flat)
        return tr.out_loss[0]

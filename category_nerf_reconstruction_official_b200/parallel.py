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


def pack_counts_and_max(counts: torch.Tensor, max_depth: torch.Tensor) -> torch.Tensor:
    """[K,4] int32 mask counts + [K] fp32 max depth -> ONE int32 [K,5] message (the float travels as its bit pattern)."""
    return torch.cat([counts.to(torch.int32), max_depth.to(torch.float32).reshape(-1, 1).view(torch.int32)], dim=1).contiguous()


def reduce_counts_and_max(gathered: torch.Tensor):
    """[world,K,5] int32 (all-gathered ``pack_counts_and_max`` messages) -> (global counts int32 [K,4] (sum), max depth fp32 [K] (max))."""
    counts = gathered[:, :, :4].sum(dim=0, dtype=torch.int32).contiguous()
    max_depth = gathered[:, :, 4].contiguous().view(torch.float32).max(dim=0).values.contiguous()
    return counts, max_depth


class RaySplitStep:
    """Ray-split training step of a FusedCategoryTrainer replica (BASELINE config 4): this rank's slice of every category's rays.

    Per step TWO collectives, both on the compute stream:
      1. before the sampler, ONE all-gather of the rank's sampler stats rows ([K,8] int32, cnf_sample_stats): they hold the three
         loss-mask counts (denominators N_d, N_o, N_s, src/render_rays.py:86 -- functions of the inputs only) and the bit pattern
         of the local maximum depth
         (upper bound of the invalid-depth bins, src/scene_cateogries.py:486); every rank reduces the gathered messages itself
         (sum / max), so the sampler and the losses see the same globals as a single-GPU step over the whole batch;
      2. after the fused step, ONE all-reduce (sum) of the trainer's flat buffer {g_params | g_B | g_shape | g_tex | out_loss}
         (the gradient tensors are views of it: no packing).  The regulariser and the non-additive out_loss slots are produced
         by rank 0 only (cnf_train_step flags bit0 / bit2), so the sum is exactly the single-GPU result up to fp32 summation order.
    The whole step (collectives included) can be captured into a CUDA graph (``capture``)."""

    def __init__(self, trainer, group=None):
        self.tr = trainer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._gathered = None
        self._graph = None
        self._side = None

    # ---- collectives
    def global_counts_and_max(self, local_counts: torch.Tensor, local_max_depth: torch.Tensor):
        msg = pack_counts_and_max(local_counts, local_max_depth)
        K = msg.shape[0]
        if self._gathered is None or self._gathered.shape[0] != self.world * K:
            self._gathered = torch.empty(self.world * K, 5, dtype=torch.int32, device=msg.device)    # ranks concatenated along dim 0
        dist.all_gather_into_tensor(self._gathered, msg, group=self.group)
        return reduce_counts_and_max(self._gathered.view(self.world, K, 5))

    def allreduce_outputs(self):
        dist.all_reduce(self.tr.grad_flat, op=dist.ReduceOp.SUM, group=self.group)

    # ---- steps
    def step_from_samples(self, origins, dirs, z, depth, rgbs, valid_u8, indices, counts_global, prepared=False):
        """Fused step on this rank's rays with the GLOBAL counts, then the gradient all-reduce."""
        tr = self.tr
        tr.step_from_samples(origins, dirs, z, depth, rgbs, valid_u8, indices, counts_global, add_reg=(self.rank == 0),
                             prepared=prepared, secondary=(self.rank != 0))
        self.allreduce_outputs()
        return tr.out_loss[0]

    def global_stats(self, local_stats: torch.Tensor):
        """All-gather the ranks' sampler stats rows (int32 [K,8]) and reduce them on every rank: (counts [K,4], max depth [K])."""
        from . import ops
        K = local_stats.shape[0]
        if self._gathered is None or self._gathered.shape != (self.world * K, 8):
            self._gathered = torch.empty(self.world * K, 8, dtype=torch.int32, device=local_stats.device)
        dist.all_gather_into_tensor(self._gathered, local_stats, group=self.group)
        return ops.reduce_stats(self._gathered.view(self.world, K, 8))

    def step(self, rgbs, depth, origins, dirs, indices, rand: Optional[Dict[str, torch.Tensor]] = None):
        """Inputs are THIS rank's rays ([K, R_local, ...]). Returns the global loss (identical on all ranks).
        Critical path: sampler count pass (cnf_sample_stats: local mask counts + max depth in ONE launch) -> all-gather ->
        cnf_reduce_stats -> sampler -> fused step -> all-reduce; the instance sort / accumulator clear (prepare) and the random
        draws run on side streams next to the first collective."""
        from . import ops, sampling
        tr = self.tr
        dev = tr.device
        if indices.dtype != torch.int32:
            indices = indices.to(torch.int32)
        indices = indices.contiguous()
        main = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        tr._workspace(rgbs.shape[1])                                         # allocate on the main stream
        self._side[0].wait_stream(main)
        with torch.cuda.stream(self._side[0]):
            tr.prepare(indices, rgbs.shape[1])
        if rand is None:
            self._side[1].wait_stream(main)
            with torch.cuda.stream(self._side[1]):
                # (allocated on the side stream; the next step's side-stream work starts with wait_stream(main), i.e. after
                # this step's sampler has read them)
                rand = sampling.draw_randoms(rgbs, depth, tr.N, tr.M, tr.eps, tr.min_bound, tr.this_obj)
            drew = True
        else:
            drew = False
        stats = ops.sample_stats(rgbs, depth, tr.min_bound, tr.this_obj)     # scene_cateogries.py:484-486, loss.py:33-36 (inputs only)
        counts, max_depth = self.global_stats(stats)
        if drew:
            main.wait_stream(self._side[1])
        z, _, valid_s, _, _ = ops.sample_batched(rgbs, depth, origins, dirs, tr.N, tr.M, tr.eps, tr.stop_eps, rand,
                                                 tr.min_bound, tr.this_obj, want_pcs=False, max_depth=max_depth)
        main.wait_stream(self._side[0])
        return self.step_from_samples(origins, dirs, z, depth, rgbs, valid_s.view(torch.uint8), indices, counts, prepared=True)

    def capture(self, rgbs, depth, origins, dirs, indices):
        """Capture ``step`` (sampler, both collectives, fused step) over THESE static tensors into a CUDA graph; returns ``replay``."""
        dev = self.tr.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.step(rgbs, depth, origins, dirs, indices)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step(rgbs, depth, origins, dirs, indices)
        self._graph = graph
        return graph.replay

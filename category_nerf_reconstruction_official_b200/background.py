"""The background model's training step (reference train.py:172-180) without autograd, on tensor cores.

``FusedBackgroundTrainer`` holds the background ``OccupancyMap`` (hidden size 128, src/model.py:86-154) and its positional
embedding as flat fp32 buffers and runs one step per call through the C ABI only:

    sampler (cnf_sample_fwd) -> embedding + OccupancyMap forward (cnf_occmap_tc_fwd_points: tcgen05, bf16 operands, the embedding
    is computed inside the kernel) -> compositing + losses forward / backward (cnf_composite_loss_fwd / _bwd)
    -> OccupancyMap + embedding backward (cnf_occmap_tc_bwd_points)

Every buffer is preallocated, nothing is read back, so the step can be captured into a CUDA graph (``capture``).  Parameter
order and shapes are the module's (``model.OccupancyMap(87, 42, 128)``), so state interchanges with the drop-in module path."""
from __future__ import annotations

import contextlib
import os
from typing import Dict, Optional, Sequence

import torch

from . import ops, sampling
from ._lib import check, lib_for, ptr, stream
from .embedding import _DIRS


class FusedBackgroundTrainer:
    def __init__(self, scale: float = 5.0, device="cuda:0", n_bins_cam2surface: int = 5, n_bins: int = 9,
                 surface_eps: float = 0.1, stop_eps: float = 0.05, min_bound: float = 0.0, this_obj: int = 1,
                 color_scaling: float = 5.0, opacity_scaling: float = 10.0, generator: Optional[torch.Generator] = None):
        self.device = torch.device(device)
        lib_for(torch.empty(1, device=self.device))          # no CPU path: raises unless this is a CUDA (sm_100) device
        self.N, self.M = n_bins_cam2surface, n_bins
        self.S = self.N + self.M
        self.eps, self.stop_eps, self.min_bound, self.this_obj = surface_eps, stop_eps, min_bound, this_obj
        self.cs, self.os = color_scaling, opacity_scaling
        from .model import OccupancyMap, flat_params, init_weights
        g = generator
        with (torch.random.fork_rng(devices=[]) if g is not None else contextlib.nullcontext()):
            if g is not None:
                torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g)))
            m = OccupancyMap(87, 42, hidden_size=128)
            m.apply(init_weights)
        self._shapes = [tuple(p.shape) for p in m.parameters()]
        dev = self.device
        self.params = flat_params(m).detach().to(dev).contiguous()
        self.pe_B = torch.tensor(_DIRS, dtype=torch.float32, device=dev).contiguous()
        self.pe_scale = torch.tensor([float(scale)], dtype=torch.float32, device=dev)
        self.g_params = torch.zeros_like(self.params)
        self.g_B = torch.zeros_like(self.pe_B)
        self.out_loss = torch.zeros(7, device=dev)          # {loss, e_depth, e_obj, e_sem, depth, color, opacity}
        self._one = torch.ones(1, device=dev)
        self._buf_key = None
        self._side = None
        self._side2 = None
        self._side3 = None
        self._rand_key = None
        self._opt_state = None
        self.launches_per_step = 0

    # ------------------------------------------------------------------------------------------ views / state import
    def param_views(self):
        out, off = [], 0
        for s in self._shapes:
            n = 1
            for d in s:
                n *= d
            out.append(self.params[off:off + n].view(*s))
            off += n
        return out

    def grad_views(self):
        out, off = [], 0
        for s in self._shapes:
            n = 1
            for d in s:
                n *= d
            out.append(self.g_params[off:off + n].view(*s))
            off += n
        return out

    def load(self, fc_params: Sequence[torch.Tensor], B: torch.Tensor):
        """Copy reference-shaped tensors (14 parameters in named_parameters() order, B [21,3])."""
        self.params.copy_(torch.cat([p.detach().reshape(-1).float() for p in fc_params]).to(self.device))
        self.pe_B.copy_(B.detach().to(self.device))

    # ------------------------------------------------------------------------------------------ one step
    def _buffers(self, R: int):
        if self._buf_key != R:
            lib = lib_for(self.params)
            dev, S = self.device, self.S
            P = R * S
            f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
            self._alpha, self._rgb = f(P), f(P, 3)
            self._g_alpha, self._g_rgb = f(P), f(P, 3)
            self._rays = f(R, 6)
            self._counts = torch.empty(1, 4, device=dev, dtype=torch.int32)
            self._gt_color = f(R, 3)
            self._ws = torch.empty(lib.cnf_occmap_tc_workspace_bytes(P), device=dev, dtype=torch.uint8)
            self._cws = torch.empty(lib.cnf_composite_workspace_bytes(1, R), device=dev, dtype=torch.uint8)
            self._buf_key = R
        return self._buf_key

    def prepare(self, R: int):
        """Parameter-only half of a step (cnf_occmap_tc_prepare): bf16 operand images, constants, zeroed gradient buffer."""
        self._buffers(R)
        lib = lib_for(self.params)
        check(lib.cnf_occmap_tc_prepare(R * self.S, ptr(self.params), ptr(self.pe_B), ptr(self.pe_scale), ptr(self._ws),
                                        self._ws.numel(), stream()))

    def step_from_samples(self, pcs, z, gt_depth, gt_rgb_u8, labels, mask_depth, prepared: bool = False):
        """Forward + backward given sampled points: pcs [R,S,3], z [R,S], gt_depth [R], gt_rgb_u8 u8 [R,3+] (first three channels),
        labels u8 [R], mask_depth u8/bool [R].  Fills self.g_params / self.g_B / self.out_loss; returns the loss (0-d, async).
        The small input-only pieces (gt colours, mask counts, gradient clear) run on a side stream next to the embedding / MLP."""
        lib = lib_for(self.params)
        R, S = z.shape
        assert S == self.S
        P = R * S
        self._buffers(R)
        pcs = pcs.contiguous(); z = z.contiguous(); gt_depth = gt_depth.contiguous()
        labels = labels.contiguous()
        mask_depth = mask_depth.view(torch.uint8) if mask_depth.dtype == torch.bool else mask_depth
        mask_depth = mask_depth.contiguous()
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            torch.div(gt_rgb_u8[:, :3], 255.0, out=self._gt_color)                               # train.py:128 (gt_rgb / 255)
            check(lib.cnf_mask_counts(1, R, ptr(labels), ptr(mask_depth), ptr(self._counts), stream()))
        st = stream()
        flags = 1 if prepared else 0
        check(lib.cnf_occmap_tc_fwd_points(P, ptr(self.params), ptr(self.pe_B), ptr(self.pe_scale), ptr(pcs), ptr(self._alpha),
                                           ptr(self._rgb), ptr(self._ws), self._ws.numel(), flags, st))
        main.wait_stream(self._side)
        args = (1, R, S, ptr(self._alpha), ptr(self._rgb), ptr(z), ptr(gt_depth), ptr(self._gt_color), ptr(labels), ptr(mask_depth),
                ptr(self._counts), float(self.cs), float(self.os))
        check(lib.cnf_composite_loss_fwd_bwd(*args, ptr(self._rays), ptr(self.out_loss), ptr(self._g_alpha), ptr(self._g_rgb),
                                             ptr(self._cws), self._cws.numel(), st))
        check(lib.cnf_occmap_tc_bwd_points(P, ptr(self.params), ptr(pcs), ptr(self._rgb), ptr(self._g_alpha), ptr(self._g_rgb),
                                           ptr(self._ws), self._ws.numel(), ptr(self.g_params), ptr(self.g_B), flags, st))
        return self.out_loss[0]

    def step(self, rgbs, depth, origins, dirs, rand: Optional[Dict[str, torch.Tensor]] = None):
        """Sample + step: rgbs u8 [R,4] (rgb + label), depth [R], origins / dirs [R,3] (scene_cateogries.py:453-546 sampler)."""
        prefetch = rand is None
        main = torch.cuda.current_stream(self.device)
        split = os.environ.get("CNF_BG_NO_PREPARE_SPLIT") is None
        self._buffers(rgbs.shape[0])                         # allocate on the main stream
        if split:
            if self._side3 is None:
                self._side3 = torch.cuda.Stream(device=self.device)
            self._side3.wait_stream(main)
            with torch.cuda.stream(self._side3):             # parameter-only half next to the sampler
                self.prepare(rgbs.shape[0])
        if prefetch:
            rand = self._prefetched_randoms(rgbs.shape[0])
        elif rand["u_inv"].dim() == 2:
            rand = {k: v[None] for k, v in rand.items()}
        z, pcs, valid, labels, _ = ops.sample_batched(rgbs[None], depth[None], origins[None], dirs[None], self.N, self.M, self.eps,
                                                       self.stop_eps, rand, self.min_bound, self.this_obj, want_pcs=True)
        if prefetch:          # the next step's offsets: drawn into the same buffers once this step's sampler has read them
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(device=self.device)
            self._side2.wait_stream(main)
            with torch.cuda.stream(self._side2):
                self._draw_randoms_into_buffers()
        # kernels of libcnf_sm100.so: sampler 1 | mask counts, prep, OccupancyMap fwd | composite fwd+bwd (tile + finalize) |
        # OccupancyMap bwd, gradient scatter
        self.launches_per_step = 1 + 3 + 2 + 2
        if split:
            main.wait_stream(self._side3)
        loss = self.step_from_samples(pcs[0], z[0], depth, rgbs, labels[0], valid[0], prepared=split)
        if prefetch:
            main.wait_stream(self._side2)
        return loss

    def _draw_randoms_into_buffers(self):
        self._rand_u.uniform_()
        self._rand_n.normal_(mean=0., std=self.eps / 3.)

    def _prefetched_randoms(self, R: int):
        N, M = self.N, self.M
        n1, n2, n3 = R * (N + M), R * N, R * M
        if self._rand_key != R:
            self._rand_u = torch.empty(n1 + n2 + n3, device=self.device)
            self._rand_n = torch.empty(1, R, M, device=self.device)
            self._rand_key = R
            self._draw_randoms_into_buffers()
        u = self._rand_u
        return {"u_inv": u[:n1].view(1, R, N + M), "u_val": u[n1:n1 + n2].view(1, R, N), "n_obj": self._rand_n,
                "u_oth": u[n1 + n2:].view(1, R, M)}

    def _state_tensors(self):
        return [self.params, self.pe_B]

    def snapshot_state(self):
        st = self._opt_state
        return {"state": [t.clone() for t in self._state_tensors()],
                "opt": None if st is None else {"t": st["t"].clone(), "m": [t.clone() for t in st["m"]], "v": [t.clone() for t in st["v"]]}}

    def restore_state(self, snap):
        for t, c in zip(self._state_tensors(), snap["state"]):
            t.copy_(c)
        if self._opt_state is not None:
            st = self._opt_state
            if snap["opt"] is None:
                st["t"].zero_()
                for t in st["m"] + st["v"]:
                    t.zero_()
            else:
                st["t"].copy_(snap["opt"]["t"])
                for t, c in zip(st["m"] + st["v"], snap["opt"]["m"] + snap["opt"]["v"]):
                    t.copy_(c)

    def capture(self, rgbs, depth, origins, dirs):
        """Capture sampler + step over THESE (static) input tensors into a CUDA graph; returns ``replay``."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.step(rgbs, depth, origins, dirs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step(rgbs, depth, origins, dirs)
        self._graph = graph
        return graph.replay

    def loss_terms(self) -> Dict[str, torch.Tensor]:
        o = self.out_loss
        return {"loss": o[0], "depth": o[4:5], "color": o[5:6], "opacity": o[6:7]}

    # ------------------------------------------------------------------------------------------ optimizer
    def adamw_step(self, lr=1e-3, weight_decay=0.013, betas=(0.9, 0.999), eps=1e-8):
        """Fused AdamW over the flat buffers (train.py:40,54-64,183-184): one launch, device-side step count (graph capturable)."""
        if self._opt_state is None:
            z = torch.zeros_like
            self._opt_state = {"t": torch.zeros(2, dtype=torch.int32, device=self.device),
                               "m": [z(self.params), z(self.pe_B)], "v": [z(self.params), z(self.pe_B)]}
        st = self._opt_state
        ops.adamw_groups_([(self.params, self.g_params, st["m"][0], st["v"][0], lr, weight_decay),
                           (self.pe_B, self.g_B, st["m"][1], st["v"][1], lr, weight_decay)], betas[0], betas[1], eps, st["t"])


def capture_iteration(obj_trainer, obj_inputs, bg_trainer, bg_inputs, with_optimizer: bool = False, opt_kwargs=None):
    """One training iteration of the reference (train.py:123-184: every object category, then the background model) as ONE CUDA
    graph with the two steps on parallel branches: ``obj_inputs`` = (rgbs, depth, origins, dirs, indices) of
    ``FusedCategoryTrainer.step``, ``bg_inputs`` = (rgbs, depth, origins, dirs) of ``FusedBackgroundTrainer.step``; static tensors,
    updated in place between replays.  Returns ``replay``.  With ``with_optimizer`` the fused AdamW updates are part of the graph;
    their step count is a device counter that every replay advances (cnf_adamw_groups), and the warm-up iterations are rolled
    back, so N replays == N steps of torch.optim.AdamW from the state at capture time."""
    opt_kwargs = opt_kwargs or {}
    dev = obj_trainer.device
    branch = torch.cuda.Stream(device=dev)

    def iteration():
        main = torch.cuda.current_stream(dev)
        branch.wait_stream(main)
        with torch.cuda.stream(branch):
            bg_trainer.step(*bg_inputs)
            if with_optimizer:
                bg_trainer.adamw_step(**{k: v for k, v in opt_kwargs.items() if k in ("lr", "weight_decay", "betas", "eps")})
        obj_trainer.step(*obj_inputs)
        if with_optimizer:
            obj_trainer.adamw_step(**opt_kwargs)
        main.wait_stream(branch)

    warm = torch.cuda.Stream(device=dev)
    warm.wait_stream(torch.cuda.current_stream(dev))
    snaps = (obj_trainer.snapshot_state(), bg_trainer.snapshot_state())      # the warm-up iterations must not train
    with torch.cuda.stream(warm):
        for _ in range(2):
            iteration()
        obj_trainer.restore_state(snaps[0]); bg_trainer.restore_state(snaps[1])
    torch.cuda.current_stream(dev).wait_stream(warm)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        iteration()
    obj_trainer._iteration_graph = graph
    return graph.replay

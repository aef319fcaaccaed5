"""The fused tcgen05 training step (cnf_train_step): sample -> embed -> CodeNeRF -> composite -> losses, fwd+bwd.

``FusedCategoryTrainer`` holds K categories' state as flat fp32 buffers (MLP parameters in the reference's
named_parameters() order, PE projection, code tables) plus preallocated gradient / workspace buffers, and
runs one training step of train.py:123-184 per call: the sampler (3 small launches) and the fused step
(prep, fused fwd+bwd, PE-gradient reduce, finalize, loss) with optional fused AdamW.  Everything is
asynchronous; nothing is read back unless the caller asks (``.item()`` on a result)."""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import ops, sampling
from ._lib import check, lib_for, ptr, stream
from .embedding import _DIRS

MAXOBJ = 4096        # cnf_fused_a.inc MAX_INSTANCES


class FusedCategoryTrainer:
    def __init__(self, n_cat: int, n_obj, latent_dim: int = 256, scale=2.0, device="cuda:0",
                 n_bins_cam2surface: int = 1, n_bins: int = 9, surface_eps: float = 0.1, stop_eps: float = 0.05,
                 min_bound: float = 0.0, this_obj: int = 1, color_scaling: float = 5.0, opacity_scaling: float = 10.0,
                 reg_scaling: float = 0.0005, generator: Optional[torch.Generator] = None):
        self.K = n_cat
        self.L = latent_dim
        self.device = torch.device(device)
        n_obj = [n_obj] * n_cat if isinstance(n_obj, int) else list(n_obj)
        assert len(n_obj) == n_cat
        self.n_obj_list = n_obj
        self.n_obj_max = max(n_obj)
        if self.n_obj_max > MAXOBJ:
            raise NotImplementedError("fused step supports <= %d instances per category (got %d); use the "
                                      "module-level CodeNeRF path for larger categories" % (MAXOBJ, self.n_obj_max))
        self.N, self.M = n_bins_cam2surface, n_bins
        self.S = self.N + self.M
        self.eps, self.stop_eps, self.min_bound, self.this_obj = surface_eps, stop_eps, min_bound, this_obj
        self.cs, self.os, self.reg = color_scaling, opacity_scaling, reg_scaling
        lib = lib_for(torch.empty(1, device=self.device))
        self.n_param = lib.cnf_codenerf_param_count(32, latent_dim)
        dev = self.device
        # --- state (random init as the reference: xavier-normal weights, default Linear biases, randn/sqrt(L/2) codes)
        from .model import CodeNeRF, init_weights, flat_params
        g = generator
        ps = []
        import contextlib
        with (torch.random.fork_rng(devices=[]) if g is not None else contextlib.nullcontext()):
            if g is not None:               # nn.Linear / xavier_normal_ draw from the global CPU RNG: seed it from the generator
                torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g)))
            for _ in range(n_cat):
                m = CodeNeRF(87, 42, W=32, latent_dim=latent_dim)
                m.apply(init_weights)
                ps.append(flat_params(m).detach())
        self.params = torch.stack(ps).to(dev).contiguous()
        self.pe_B = torch.tensor(_DIRS, dtype=torch.float32).repeat(n_cat, 1, 1).to(dev).contiguous()
        sc = torch.as_tensor(scale, dtype=torch.float32).reshape(-1)
        self.pe_scale = (sc.repeat(n_cat) if sc.numel() == 1 else sc).to(dev).contiguous()
        self.shape_codes = (torch.randn(n_cat, self.n_obj_max, latent_dim, generator=g) / math.sqrt(latent_dim / 2)).to(dev)
        self.tex_codes = (torch.randn(n_cat, self.n_obj_max, latent_dim, generator=g) / math.sqrt(latent_dim / 2)).to(dev)
        self.n_obj = torch.tensor(n_obj, dtype=torch.int32, device=dev)
        # --- gradients / outputs: views of ONE flat fp32 buffer {g_params | g_B | g_shape | g_tex | out_loss} (16-byte aligned
        # segments), so that a ray-split step sums everything over ranks with a single all-reduce and no packing copies
        segs = [self.params.numel(), self.pe_B.numel(), self.shape_codes.numel(), self.tex_codes.numel(), 4 + 5 * n_cat]
        offs, tot = [], 0
        for n in segs:
            offs.append(tot)
            tot += (n + 3) // 4 * 4
        self.grad_flat = torch.zeros(tot, device=dev)
        view = lambda i, like: self.grad_flat[offs[i]:offs[i] + segs[i]].view(like)
        self.g_params = view(0, self.params.shape)
        self.g_B = view(1, self.pe_B.shape)
        self.g_shape = view(2, self.shape_codes.shape)
        self.g_tex = view(3, self.tex_codes.shape)
        self.out_loss = view(4, (4 + 5 * n_cat,))
        self._ws = None
        self._ws_key = None
        self._opt_state = None
        self._side = None
        self._side2 = None
        self._rand_key = None
        self.launches_per_step = 0
        # False: every step rebuilds the fp16 operand images from the fp32 parameters (safe if parameters are edited by hand).
        # True: they are rebuilt when the parameters change through this class (adamw_step, load_category, refresh_operands).
        self.operands_follow_optimizer = False
        self._operands_R = None
        # Sampler inside the fused kernel for small batches (cnf_train_step_sampled): no sampler launch, no z round trip through HBM,
        # depths bit-exact.  OFF by default: measured on the C2 batch it is SLOWER (0.098 vs 0.088 ms per step) -- the sampling
        # arithmetic in both roles' tile prologues costs 5 us of kernel time, the prepare launch (sort + ranks) is serial with the
        # kernel instead of parallel to the sampler, and the next step's random draws can no longer overlap the fused kernel
        # (it reads them until it ends).  Set CNF_FUSED_SAMPLER=1 or this attribute to use it.
        self.sample_in_kernel = os.environ.get("CNF_FUSED_SAMPLER") is not None
        self.stats = torch.zeros(n_cat, 8, dtype=torch.int32, device=dev)     # sampler stats of the last step (fused-sampler path)
        self.keep_z = False                     # verification: keep the depths the fused kernel used in self.z_used [K,R,S]
        self.z_used = None

    # ------------------------------------------------------------------------------------------ state import
    def load_category(self, k: int, fc_params: Sequence[torch.Tensor], B: torch.Tensor, shape_codes: torch.Tensor,
                      tex_codes: torch.Tensor):
        """Copy one category's reference-shaped tensors (28 params in named_parameters() order, B [21,3], code tables)."""
        flat = torch.cat([p.reshape(-1) for p in fc_params]).to(self.device)
        assert flat.numel() == self.n_param
        self.params[k].copy_(flat)
        self.pe_B[k].copy_(B.to(self.device))
        n = shape_codes.shape[0]
        self.shape_codes[k].zero_(); self.tex_codes[k].zero_()
        self.shape_codes[k, :n].copy_(shape_codes.to(self.device))
        self.tex_codes[k, :n].copy_(tex_codes.to(self.device))
        self._operands_R = None                     # operand images are stale

    def param_views(self, k: int) -> List[torch.Tensor]:
        """The 28 nn.Linear-shaped views of category k's flat parameter buffer (and same for grads via grad_views)."""
        return self._views(self.params[k])

    def grad_views(self, k: int) -> List[torch.Tensor]:
        return self._views(self.g_params[k])

    def _views(self, flat):
        W, L, E1, E2 = 32, self.L, 87, 42
        shapes = [(W, E1), (W,), (W, L), (W,), (W, W), (W,), (W, L), (W,), (W, W), (W,), (W, W + E1), (W,), (W, L), (W,),
                  (W, W), (W,), (1, W), (1,), (W, W + E2), (W,), (W, W), (W,), (W, L), (W,), (W // 2, W), (W // 2,),
                  (3, W // 2), (3,)]
        out, off = [], 0
        for s in shapes:
            n = 1
            for d in s:
                n *= d
            out.append(flat[off:off + n].view(*s))
            off += n
        return out

    # ------------------------------------------------------------------------------------------ one step
    def _workspace(self, R: int):
        key = (R,)
        if self._ws_key != key:
            lib = lib_for(self.params)
            nbytes = lib.cnf_train_step_workspace_bytes(self.K, R, self.S, self.L, self.n_obj_max)
            self._ws = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
            self._ws_key = key
        return self._ws

    def prepare(self, indices, R: int, what: Optional[int] = None):
        """First half of a step (cnf_train_step_prepare), independent of the sampled depths, so ``step`` runs it on a side stream
        next to the sampler.  Two parts: the OPERAND half (fp16 weight images, per-instance latent biases: functions of the
        parameters only) and the INPUT half (instance sort of this step's rays, accumulator clear).  With
        ``operands_follow_optimizer`` the operand half is refreshed by ``adamw_step`` / ``refresh_operands`` (once per parameter
        change, like the low-precision weight copy of mixed-precision training) and a step only runs the input half."""
        lib = lib_for(self.params)
        ws = self._workspace(R)
        if what is None:
            what = 2 if (self.operands_follow_optimizer and self._operands_R == R) else 3
        check(lib.cnf_train_step_prepare(self.K, R, self.S, self.L, self.n_obj_max, ptr(self.params), ptr(self.pe_B),
                                         ptr(self.pe_scale), ptr(self.shape_codes), ptr(self.tex_codes), ptr(self.n_obj),
                                         ptr(indices) if indices is not None else None, int(what), ptr(ws), ws.numel(), stream()))
        if what & 1:
            self._operands_R = R

    def refresh_operands(self, R: Optional[int] = None):
        """Rebuild the operand half of the step's workspace from the current parameters / codes (call after changing them by hand
        when ``operands_follow_optimizer`` is set; ``adamw_step`` and ``load_category`` do it themselves)."""
        R = R if R is not None else (self._ws_key[0] if self._ws_key else None)
        if R is not None:
            self.prepare(None, R, what=1)

    def step_from_samples(self, origins, dirs, z, gt_depth, rgbs, mask_depth, indices, counts, add_reg: bool = True,
                          prepared: bool = False, secondary: bool = False):
        """Fused fwd+bwd given sampled depths z [K,R,S]. rgbs u8 [K,R,4] (rgb + label); mask_depth u8/bool [K,R];
        indices int32 [K,R]; counts int32 [K,4] (or a [K,>=4]-strided view, e.g. the sampler's stats[:, 4:8]).
        Fills self.g_* and self.out_loss; returns the loss (0-d, async).  ``secondary``: this call is a non-first part of a ray-split
        step whose outputs are SUMMED over parts: the non-additive slots of out_loss (empty-mask flags, regulariser values) are
        written as zeros so that the sum carries them once."""
        lib = lib_for(self.params)
        K, R, S = z.shape
        assert K == self.K and S == self.S
        ws = self._workspace(R)
        if indices.dtype != torch.int32:
            indices = indices.to(torch.int32)
        if mask_depth.dtype != torch.uint8:
            mask_depth = mask_depth.to(torch.uint8)
        assert counts.dtype == torch.int32 and counts.stride(1) == 1 and 4 <= counts.stride(0) < 256
        flags = (1 if add_reg else 0) | (2 if prepared else 0) | (4 if secondary else 0) | (counts.stride(0) << 8)
        check(lib.cnf_train_step(K, R, S, self.L, self.n_obj_max, ptr(self.params), ptr(self.pe_B), ptr(self.pe_scale),
                                 ptr(self.shape_codes), ptr(self.tex_codes), ptr(self.n_obj), ptr(origins.contiguous()),
                                 ptr(dirs.contiguous()), ptr(z.contiguous()), ptr(gt_depth.contiguous()),
                                 ptr(rgbs.contiguous()), ptr(mask_depth.contiguous()), ptr(indices.contiguous()),
                                 counts.data_ptr(), float(self.cs), float(self.os), float(self.reg),     # row-strided view allowed
                                 flags, ptr(self.g_params), ptr(self.g_B), ptr(self.g_shape), ptr(self.g_tex),
                                 ptr(self.out_loss), ptr(ws), ws.numel(), stream()))
        return self.out_loss[0]

    def step(self, rgbs, depth, origins, dirs, indices, rand: Optional[Dict[str, torch.Tensor]] = None, add_reg=True):
        """Sample + fused step for K categories: rgbs u8 [K,R,4], depth [K,R], origins/dirs [K,R,3], indices [K,R].
        The step's sampler-independent half (prep + instance sort) runs on a side stream next to the random draws and the
        sampler; both join before the fused kernel (the fork/join is captured as graph dependencies by ``capture``)."""
        if indices.dtype != torch.int32:
            indices = indices.to(torch.int32)
        indices = indices.contiguous()
        if self.sample_in_kernel and rgbs.shape[1] <= self._sampled_max_rays():
            return self._step_sampled(rgbs, depth, origins, dirs, indices, rand, add_reg)
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._workspace(rgbs.shape[1])                       # allocate on the main stream
        overlap = os.environ.get("CNF_NO_SIDE_STREAM") is None      # debugging aid: run the prepare half inline
        if overlap:
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.prepare(indices, rgbs.shape[1])
        else:
            self.prepare(indices, rgbs.shape[1])
        prefetch = rand is None and overlap and os.environ.get("CNF_NO_RNG_PREFETCH") is None
        if prefetch:
            rand = self._prefetched_randoms(rgbs.shape[0], rgbs.shape[1])
        elif rand is None:
            rand = sampling.draw_randoms(rgbs, depth, self.N, self.M, self.eps, self.min_bound, self.this_obj)
        z, _, valid, labels, stats = ops.sample_batched(rgbs, depth, origins, dirs, self.N, self.M, self.eps, self.stop_eps,
                                                       rand, self.min_bound, self.this_obj, want_pcs=False)
        if prefetch:
            # the NEXT step's offsets are drawn into the same buffers as soon as this step's sampler has read them, on a third
            # stream next to the fused kernel (same generator, same order of draws: only their position in time changes)
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(device=self.device)
            self._side2.wait_stream(main)
            with torch.cuda.stream(self._side2):
                self._draw_randoms_into_buffers()
        if overlap:
            main.wait_stream(self._side)
        # kernels of libcnf_sm100.so per step: sampler (1 launch up to 4096 rays per category, else count + scan + emit), prepare half
        # (1 launch up to 32 ray blocks of 256, else prep + count + scan + scatter), fused fwd+bwd, finalize
        R_ = rgbs.shape[1]
        skip_ops = self.operands_follow_optimizer and self._operands_R == R_
        self.launches_per_step = (1 if R_ <= 4096 else 3) + (1 if (R_ + 255) // 256 <= 32 else (3 if skip_ops else 4)) + 2
        loss = self.step_from_samples(origins, dirs, z, depth, rgbs, valid.view(torch.uint8), indices, stats[:, 4:8], add_reg,
                                      prepared=True)
        if prefetch:
            main.wait_stream(self._side2)                    # join (a CUDA-graph capture needs every forked stream back)
        return loss

    def _sampled_max_rays(self) -> int:
        return lib_for(self.params).cnf_train_step_sampled_max_rays()

    def _step_sampled(self, rgbs, depth, origins, dirs, indices, rand, add_reg=True, max_depth=None, secondary=False):
        """Small batches: ONE call -- prepare launch (instance sort + sampling ranks + stats [+ operand images]), the fused kernel
        with the sampler in its tile prologues, finalize.  The next step's random offsets are drawn on a side stream next to it."""
        lib = lib_for(self.params)
        K, R = depth.shape
        N, M, S = self.N, self.M, self.S
        dev = self.device
        main = torch.cuda.current_stream(dev)
        ws = self._workspace(R)
        prefetch = rand is None and os.environ.get("CNF_NO_RNG_PREFETCH") is None
        if prefetch:
            rand = self._prefetched_randoms(K, R)
        elif rand is None:
            rand = sampling.draw_randoms(rgbs, depth, N, M, self.eps, self.min_bound, self.this_obj)
        f32c = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
        rgbs = rgbs.contiguous(); depth = f32c(depth); origins = f32c(origins); dirs = f32c(dirs)
        u_inv, u_val, n_draws, u_oth = (f32c(rand[k]) for k in ("u_inv", "u_val", "n_obj", "u_oth"))
        assert rgbs.dtype == torch.uint8 and rgbs.shape == (K, R, 4)
        assert u_inv.shape == (K, R, S) and n_draws.shape == (K, R, M) and u_oth.shape == (K, R, M) and (N == 0 or u_val.shape == (K, R, N))
        lin_S, lin_M = ops.linspace_table(S, dev), ops.linspace_table(M, dev)
        lin_N = ops.linspace_table(N, dev) if N > 0 else lin_S
        if N == 0:
            u_val = u_inv
        if self.keep_z:
            self.z_used = torch.zeros(K, R, S, device=dev)
        skip_ops = self.operands_follow_optimizer and self._operands_R == R
        flags = (1 if add_reg else 0) | (4 if secondary else 0) | (8 if skip_ops else 0)
        check(lib.cnf_train_step_sampled(K, R, N, M, self.L, self.n_obj_max, ptr(self.params), ptr(self.pe_B), ptr(self.pe_scale),
                                         ptr(self.shape_codes), ptr(self.tex_codes), ptr(self.n_obj), ptr(rgbs), ptr(depth),
                                         ptr(origins), ptr(dirs), ptr(indices), ptr(u_inv), ptr(u_val), ptr(n_draws), ptr(u_oth),
                                         ptr(lin_S), ptr(lin_N), ptr(lin_M), float(self.min_bound), float(self.eps),
                                         float(self.stop_eps), int(self.this_obj), ptr(max_depth), float(self.cs), float(self.os),
                                         float(self.reg), flags, ptr(self.g_params), ptr(self.g_B), ptr(self.g_shape), ptr(self.g_tex),
                                         ptr(self.out_loss), ptr(self.stats), ptr(self.z_used) if self.keep_z else None, ptr(ws),
                                         ws.numel(), stream()))
        if not skip_ops:
            self._operands_R = R
        self.launches_per_step = 3                           # prepare (sort | ranks | clear [| operands]), fused fwd+bwd, finalize
        if prefetch:
            # the NEXT step's offsets are drawn into the same buffers as soon as this step's kernels have read them: on a side
            # stream that waits for the step (the draws then overlap whatever follows the step: optimizer, copies, the next H2D)
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(device=dev)
            self._side2.wait_stream(main)
            with torch.cuda.stream(self._side2):
                self._draw_randoms_into_buffers()
            main.wait_stream(self._side2)
        return self.out_loss[0]

    def _draw_randoms_into_buffers(self):
        self._rand_u.uniform_()
        self._rand_n.normal_(mean=0., std=self.eps / 3.)

    def _prefetched_randoms(self, K: int, R: int):
        """Random offsets of this step: drawn by the previous step (or right now, the first time) into fixed buffers."""
        N, M = self.N, self.M
        n1, n2, n3 = K * R * (N + M), K * R * N, K * R * M
        if self._rand_key != (K, R):
            self._rand_u = torch.empty(n1 + n2 + n3, device=self.device)
            self._rand_n = torch.empty(K, R, M, device=self.device)
            self._rand_key = (K, R)
            self._draw_randoms_into_buffers()
        u = self._rand_u
        return {"u_inv": u[:n1].view(K, R, N + M), "u_val": u[n1:n1 + n2].view(K, R, N), "n_obj": self._rand_n,
                "u_oth": u[n1 + n2:].view(K, R, M)}

    # ------------------------------------------------------------------------------------------ state snapshot (graph warm-up)
    def _state_tensors(self):
        return [self.params, self.pe_B, self.shape_codes, self.tex_codes]

    def snapshot_state(self):
        """Copies of the trainable state and the optimizer state (moments + device step counter; None if no step was taken yet)."""
        st = self._opt_state
        return {"state": [t.clone() for t in self._state_tensors()],
                "opt": None if st is None else {"t": st["t"].clone(), "m": [t.clone() for t in st["m"]], "v": [t.clone() for t in st["v"]]}}

    def restore_state(self, snap):
        for t, c in zip(self._state_tensors(), snap["state"]):
            t.copy_(c)
        if self._opt_state is not None:
            st = self._opt_state
            if snap["opt"] is None:                  # the optimizer state was created after the snapshot: back to "no step taken"
                st["t"].zero_()
                for t in st["m"] + st["v"]:
                    t.zero_()
            else:
                st["t"].copy_(snap["opt"]["t"])
                for t, c in zip(st["m"] + st["v"], snap["opt"]["m"] + snap["opt"]["v"]):
                    t.copy_(c)
        if self.operands_follow_optimizer:
            self.refresh_operands()

    def capture(self, rgbs, depth, origins, dirs, indices, add_reg=True, pre=None, post=None):
        """Capture sampler + fused step over THESE (static) input tensors into a CUDA graph and return ``replay``.
        The step is ~15 launches of microsecond-scale work at Replica-sized batches, i.e. launch-bound from Python;
        replaying the captured graph removes the host from the loop.  Update the input tensors in place between replays.
        ``pre`` / ``post``: optional callables captured before / after the step on the same stream, e.g. the upload of the step's
        inputs from pinned host memory into the static tensors and the download of ``out_loss`` into a pinned host buffer
        (``dst.copy_(src, non_blocking=True)``), so that a whole host-to-host step is ONE graph launch."""
        def body():
            if pre is not None:
                pre()
            self.step(rgbs, depth, origins, dirs, indices, add_reg=add_reg)
            if post is not None:
                post()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        snap = self.snapshot_state()                # ``post`` may be an optimizer step: the warm-up must not train
        with torch.cuda.stream(side):
            for _ in range(2):                      # warm-up: allocations, function attributes, RNG state registration
                body()
            self.restore_state(snap)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        self._graph = graph
        return graph.replay

    def loss_terms(self) -> Dict[str, torch.Tensor]:
        K = self.K
        o = self.out_loss
        return {"loss": o[0], "depth": o[4:4 + K], "color": o[4 + K:4 + 2 * K], "opacity": o[4 + 2 * K:4 + 3 * K],
                "reg_shape": o[4 + 3 * K:4 + 4 * K], "reg_texture": o[4 + 4 * K:4 + 5 * K]}

    # ------------------------------------------------------------------------------------------ optimizer
    def adamw_step(self, lr=1e-3, weight_decay=0.013, code_lr=1e-3, code_weight_decay=0.013, betas=(0.9, 0.999), eps=1e-8):
        """Fused AdamW over the flat buffers (train.py:40,54-64,183): MLP + PE with (lr, wd), codes with the code group -- ONE launch
        for the four groups.  The step count is a device counter (``self.opt_steps``), so the call can be part of a captured CUDA
        graph: every replay advances it (cnf_adamw_groups)."""
        if self._opt_state is None:
            z = torch.zeros_like
            self._opt_state = {"t": torch.zeros(2, dtype=torch.int32, device=self.device),
                               "m": [z(self.params), z(self.pe_B), z(self.shape_codes), z(self.tex_codes)],
                               "v": [z(self.params), z(self.pe_B), z(self.shape_codes), z(self.tex_codes)]}
        st = self._opt_state
        groups = [(self.params, self.g_params, st["m"][0], st["v"][0], lr, weight_decay),
                  (self.pe_B, self.g_B, st["m"][1], st["v"][1], lr, weight_decay),
                  (self.shape_codes, self.g_shape, st["m"][2], st["v"][2], code_lr, code_weight_decay),
                  (self.tex_codes, self.g_tex, st["m"][3], st["v"][3], code_lr, code_weight_decay)]
        ops.adamw_groups_(groups, betas[0], betas[1], eps, st["t"])
        if self.operands_follow_optimizer:
            self.refresh_operands()

    @property
    def opt_steps(self) -> int:
        """Optimizer steps taken so far (reads the device counter: synchronises)."""
        return 0 if self._opt_state is None else int(self._opt_state["t"][0])


class HostStepPipeline:
    """Feeds a FusedCategoryTrainer from HOST batches with the copies off the critical path (the reference uploads every step's
    pool slices synchronously, scene_cateogries.py:421-431, and reads the loss back for logging, train.py:185-192).

    Two slots.  Each slot owns one pinned host block + one device block holding a whole step's inputs (tensor views at 256-byte
    aligned offsets) and a CUDA graph of the step over that device block.  ``submit(slot)`` enqueues the slot's host->device copy on
    a copy stream (it overlaps the previous step's kernels), then the graph replay and the device->host copy of the loss terms on
    the compute stream; ``result()`` waits for the OLDEST step in flight and returns its loss terms (pinned host tensor).  Keeping one
    step in flight (submit i+1, then read i) hides the copies and the host-side launch work behind the GPU step.

        pipe = HostStepPipeline(trainer, example_batch)          # dict of CPU tensors: rgbs, depth, origins, dirs, indices
        for i, batch in enumerate(loader):
            s = i & 1
            for k, v in batch.items(): pipe.host[s][k].copy_(v)   # or let the loader write into pipe.host[s] directly
            pipe.submit(s)
            if i: loss_terms = pipe.result()
        loss_terms = pipe.result()
    """
    KEYS = ("rgbs", "depth", "origins", "dirs", "indices")

    def __init__(self, trainer: "FusedCategoryTrainer", example_batch: Dict[str, torch.Tensor], post=None):
        self.tr = trainer
        dev = trainer.device
        offs, tot = {}, 0
        for k in self.KEYS:
            v = example_batch[k]
            offs[k] = tot
            tot += (v.numel() * v.element_size() + 255) // 256 * 256
        self.bytes_per_step = sum(example_batch[k].numel() * example_batch[k].element_size() for k in self.KEYS)

        def views(block):
            return {k: block[offs[k]:offs[k] + example_batch[k].numel() * example_batch[k].element_size()]
                    .view(example_batch[k].dtype).view(example_batch[k].shape) for k in self.KEYS}

        self._host_block = [torch.empty(tot, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._dev_block = [torch.empty(tot, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.host = [views(b) for b in self._host_block]
        self.dev = [views(b) for b in self._dev_block]
        for s in range(2):
            for k in self.KEYS:
                self.host[s][k].copy_(example_batch[k])
            self._dev_block[s].copy_(self._host_block[s])
        self.out_host = [torch.empty_like(trainer.out_loss, device="cpu").pin_memory() for _ in range(2)]
        self._replay = [trainer.capture(d["rgbs"], d["depth"], d["origins"], d["dirs"], d["indices"], post=post) for d in self.dev]
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._ready = [torch.cuda.Event() for _ in range(2)]      # slot's inputs are on the device
        self._free = [torch.cuda.Event() for _ in range(2)]       # slot's device block has been consumed
        self._done = [torch.cuda.Event() for _ in range(2)]       # slot's loss terms are on the host
        main = torch.cuda.current_stream(dev)
        for e in self._free:
            e.record(main)
        self._inflight = []

    def submit(self, slot: int):
        main = torch.cuda.current_stream(self.tr.device)
        cs = self._copy_stream
        cs.wait_event(self._free[slot])
        with torch.cuda.stream(cs):
            self._dev_block[slot].copy_(self._host_block[slot], non_blocking=True)
            self._ready[slot].record(cs)
        main.wait_event(self._ready[slot])
        self._replay[slot]()
        self._free[slot].record(main)
        self.out_host[slot].copy_(self.tr.out_loss, non_blocking=True)
        self._done[slot].record(main)
        self._inflight.append(slot)

    def result(self) -> torch.Tensor:
        """Loss terms (layout of ``FusedCategoryTrainer.out_loss``) of the oldest submitted step; blocks until they are on the host."""
        slot = self._inflight.pop(0)
        self._done[slot].synchronize()
        return self.out_host[slot]

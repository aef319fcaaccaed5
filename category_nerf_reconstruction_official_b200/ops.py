"""Raw op wrappers + autograd Functions over the C ABI (include/cnf.h).  Tensors in, tensors out; every
call goes to libcnf_sm100.so on the current CUDA stream.  No CPU path exists."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import check, lib_for, ptr, stream

EMB_DIM, EMB1, EMB2 = 129, 87, 42


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
_lin_cache: Dict[Tuple[int, str], torch.Tensor] = {}


def linspace_table(n_bins: int, device) -> torch.Tensor:
    """torch.linspace(0,1,n+1) exactly as the reference builds it (scene_cateogries.py:54)."""
    key = (n_bins, str(device))
    if key not in _lin_cache:
        _lin_cache[key] = torch.linspace(0, 1, n_bins + 1, dtype=torch.float32, device=device)
    return _lin_cache[key]


def sample_batched(rgbs, depth, origins, dirs, n_cam2surf: int, n_surf: int, eps: float, stop_eps: float,
                   rand: Dict[str, torch.Tensor], min_bound: float = 0.0, this_obj: int = 1, want_pcs: bool = True,
                   max_depth: Optional[torch.Tensor] = None):
    """K-batched sampler. rgbs u8 [K,R,4]; depth [K,R]; origins/dirs [K,R,3];
    rand: 'u_inv' [K,R,S], 'u_val' [K,R,N], 'n_obj' [K,R,M], 'u_oth' [K,R,M] (row-compact per category).
    Returns z [K,R,S], pcs [K,R,S,3] | None, valid bool [K,R], labels u8 [K,R], stats int32 [K,8]."""
    lib = lib_for(depth)
    K, R = depth.shape
    N, M = n_cam2surf, n_surf
    S = N + M
    dev = depth.device
    rgbs = rgbs.contiguous(); depth = _f32c(depth); origins = _f32c(origins); dirs = _f32c(dirs)
    assert rgbs.dtype == torch.uint8 and rgbs.shape == (K, R, 4)
    u_inv, u_val, n_obj, u_oth = (_f32c(rand[k]) for k in ("u_inv", "u_val", "n_obj", "u_oth"))
    assert u_inv.shape == (K, R, S) and n_obj.shape == (K, R, M) and u_oth.shape == (K, R, M)
    assert u_val.shape == (K, R, N) or N == 0
    z = torch.empty(K, R, S, device=dev, dtype=torch.float32)
    pcs = torch.empty(K, R, S, 3, device=dev, dtype=torch.float32) if want_pcs else None
    valid = torch.empty(K, R, device=dev, dtype=torch.uint8)
    labels = torch.empty(K, R, device=dev, dtype=torch.uint8)
    stats = torch.empty(K, 8, device=dev, dtype=torch.int32)
    ws_bytes = lib.cnf_sample_workspace_bytes(K, R)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    lin_S, lin_M = linspace_table(S, dev), linspace_table(M, dev)
    lin_N = linspace_table(N, dev) if N > 0 else lin_S
    if N == 0:
        u_val = u_inv
    check(lib.cnf_sample_fwd(K, R, N, M, ptr(rgbs), ptr(depth), ptr(origins), ptr(dirs), ptr(u_inv), ptr(u_val),
                             ptr(n_obj), ptr(u_oth), ptr(lin_S), ptr(lin_N), ptr(lin_M), float(min_bound), float(eps),
                             float(stop_eps), int(this_obj), ptr(max_depth), ptr(z), ptr(pcs), ptr(valid), ptr(labels), ptr(stats),
                             ptr(ws), ws_bytes, stream()))
    return z, pcs, valid.view(torch.bool), labels, stats


def sample_stats(rgbs, depth, min_bound: float = 0.0, this_obj: int = 1) -> torch.Tensor:
    """The sampler's count pass alone: stats int32 [K,8] = {nI, nV, nO, nT, N_depth&obj, N_obj, N_sem, float_bits(max depth)}."""
    lib = lib_for(depth)
    K, R = depth.shape
    rgbs = rgbs.contiguous(); depth = _f32c(depth)
    stats = torch.empty(K, 8, device=depth.device, dtype=torch.int32)
    ws_bytes = lib.cnf_sample_workspace_bytes(K, R)
    ws = torch.empty(ws_bytes, device=depth.device, dtype=torch.uint8)
    check(lib.cnf_sample_stats(K, R, ptr(rgbs), ptr(depth), float(min_bound), int(this_obj), ptr(stats), ptr(ws), ws_bytes, stream()))
    return stats


def reduce_stats(gathered: torch.Tensor):
    """All-gathered stats rows int32 [world,K,8] -> (global mask counts int32 [K,4], global max depth fp32 [K])."""
    lib = lib_for(gathered)
    world, K, _ = gathered.shape
    counts = torch.empty(K, 4, device=gathered.device, dtype=torch.int32)
    max_depth = torch.empty(K, device=gathered.device, dtype=torch.float32)
    check(lib.cnf_reduce_stats(world, K, ptr(gathered.contiguous()), ptr(counts), ptr(max_depth), stream()))
    return counts, max_depth


def stratified_bins_op(min_d, max_d, n_bins: int, u: torch.Tensor) -> torch.Tensor:
    lib = lib_for(u)
    n_rays = u.shape[0]
    dev = u.device
    u = _f32c(u)

    def prep(v):
        if not torch.is_tensor(v):
            return torch.full((1,), float(v), device=dev, dtype=torch.float32), 1
        v = _f32c(v.to(dev))
        if v.numel() == 1:
            return v.reshape(1), 1
        assert v.numel() == n_rays
        return v.reshape(-1), 0

    mn, mn_s = prep(min_d)
    mx, mx_s = prep(max_d)
    z = torch.empty(n_rays, n_bins, device=dev, dtype=torch.float32)
    check(lib.cnf_stratified_bins(n_rays, n_bins, ptr(mn), mn_s, ptr(mx), mx_s, ptr(u), ptr(linspace_table(n_bins, dev)),
                                  ptr(z), stream()))
    return z


# ------------------------------------------------------------------------------------------------
# embedding
# ------------------------------------------------------------------------------------------------
class EmbedFn(torch.autograd.Function):
    """emb = UniDirsEmbed(x) for K stacked categories: x [K,P,3], B [K,21,3], scale [K]."""

    @staticmethod
    def forward(ctx, x, B, scale):
        lib = lib_for(x)
        x = _f32c(x); B = _f32c(B); scale = _f32c(scale)
        K, P, _ = x.shape
        emb = torch.empty(K, P, EMB_DIM, device=x.device, dtype=torch.float32)
        check(lib.cnf_embed_fwd(K, P, ptr(x), ptr(B), ptr(scale), ptr(emb), stream()))
        ctx.save_for_backward(x, B, scale)
        return emb

    @staticmethod
    def backward(ctx, g_emb):
        x, B, scale = ctx.saved_tensors
        lib = lib_for(x)
        K, P, _ = x.shape
        g_emb = _f32c(g_emb)
        g_B = torch.zeros_like(B)
        g_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        check(lib.cnf_embed_bwd(K, P, ptr(x), ptr(B), ptr(scale), ptr(g_emb), ptr(g_B), ptr(g_x), stream()))
        return g_x, (g_B if ctx.needs_input_grad[1] else None), None


def embed(x, B, scale):
    return EmbedFn.apply(x, B, scale)


# ------------------------------------------------------------------------------------------------
# MLPs (fp32 CUDA-core path)
# ------------------------------------------------------------------------------------------------
class CodeNeRFFn(torch.autograd.Function):
    """sigma, rgb = CodeNeRF(emb, shape_lat, tex_lat) for K stacked categories.
    emb [K,P,129]; shape_lat/tex_lat [K,NR,L] (point p uses row p // pts_per_row); params [K,n_param] flat."""

    @staticmethod
    def forward(ctx, emb, shape_lat, tex_lat, params, W, pts_per_row):
        lib = lib_for(emb)
        emb = _f32c(emb); shape_lat = _f32c(shape_lat); tex_lat = _f32c(tex_lat); params = _f32c(params)
        K, P, _ = emb.shape
        NR, L = shape_lat.shape[1], shape_lat.shape[2]
        assert params.shape == (K, lib.cnf_codenerf_param_count(W, L)), "flat parameter buffer has the wrong size"
        wsf = lib.cnf_codenerf_workspace_floats(P, NR, W) * K
        ws = torch.empty(wsf, device=emb.device, dtype=torch.float32)
        sigma = torch.empty(K, P, device=emb.device, dtype=torch.float32)
        rgb = torch.empty(K, P, 3, device=emb.device, dtype=torch.float32)
        check(lib.cnf_codenerf_fwd(K, P, NR, pts_per_row, W, L, ptr(params), ptr(emb), ptr(shape_lat), ptr(tex_lat),
                                   ptr(sigma), ptr(rgb), ptr(ws), wsf, stream()))
        ctx.save_for_backward(emb, shape_lat, tex_lat, params, rgb, ws)
        ctx.meta = (K, P, NR, pts_per_row, W, L, wsf)
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        emb, shape_lat, tex_lat, params, rgb, ws = ctx.saved_tensors
        K, P, NR, ppr, W, L, wsf = ctx.meta
        lib = lib_for(emb)
        g_sigma = _f32c(g_sigma) if g_sigma is not None else torch.zeros(K, P, device=emb.device)
        g_rgb = _f32c(g_rgb) if g_rgb is not None else torch.zeros(K, P, 3, device=emb.device)
        g_params = torch.zeros_like(params)
        g_emb = torch.empty_like(emb) if ctx.needs_input_grad[0] else None
        g_sh = torch.empty_like(shape_lat) if ctx.needs_input_grad[1] else None
        g_tx = torch.empty_like(tex_lat) if ctx.needs_input_grad[2] else None
        check(lib.cnf_codenerf_bwd(K, P, NR, ppr, W, L, ptr(params), ptr(emb), ptr(shape_lat), ptr(tex_lat), ptr(rgb),
                                   ptr(g_sigma), ptr(g_rgb), ptr(ws), wsf, ptr(g_params), ptr(g_emb), ptr(g_sh),
                                   ptr(g_tx), stream()))
        return g_emb, g_sh, g_tx, g_params, None, None


class OccMapFn(torch.autograd.Function):
    """alpha, rgb = OccupancyMap(emb) for K stacked models. emb [K,P,129]; params [K,n_param] flat."""

    @staticmethod
    def forward(ctx, emb, params, H):
        lib = lib_for(emb)
        emb = _f32c(emb); params = _f32c(params)
        K, P, _ = emb.shape
        assert params.shape == (K, lib.cnf_occmap_param_count(H))
        wsf = lib.cnf_occmap_workspace_floats(P, H) * K
        ws = torch.empty(wsf, device=emb.device, dtype=torch.float32)
        alpha = torch.empty(K, P, device=emb.device, dtype=torch.float32)
        rgb = torch.empty(K, P, 3, device=emb.device, dtype=torch.float32)
        check(lib.cnf_occmap_fwd(K, P, H, ptr(params), ptr(emb), ptr(alpha), ptr(rgb), ptr(ws), wsf, stream()))
        ctx.save_for_backward(emb, params, rgb, ws)
        ctx.meta = (K, P, H, wsf)
        return alpha, rgb

    @staticmethod
    def backward(ctx, g_alpha, g_rgb):
        emb, params, rgb, ws = ctx.saved_tensors
        K, P, H, wsf = ctx.meta
        lib = lib_for(emb)
        g_alpha = _f32c(g_alpha) if g_alpha is not None else torch.zeros(K, P, device=emb.device)
        g_rgb = _f32c(g_rgb) if g_rgb is not None else torch.zeros(K, P, 3, device=emb.device)
        g_params = torch.zeros_like(params)
        g_emb = torch.empty_like(emb) if ctx.needs_input_grad[0] else None
        check(lib.cnf_occmap_bwd(K, P, H, ptr(params), ptr(emb), ptr(rgb), ptr(g_alpha), ptr(g_rgb), ptr(ws), wsf,
                                 ptr(g_params), ptr(g_emb), stream()))
        return g_emb, g_params, None


class OccMapTCFn(torch.autograd.Function):
    """alpha, rgb = OccupancyMap(emb), hidden size 128, on tensor cores (cnf_occmap_tc_fwd / _bwd; bf16 operands).
    emb [P,129]; params [n_param] flat (one model)."""

    @staticmethod
    def forward(ctx, emb, params):
        lib = lib_for(emb)
        emb = _f32c(emb); params = _f32c(params)
        P = emb.shape[0]
        assert emb.shape[1] == EMB_DIM and params.numel() == lib.cnf_occmap_param_count(128)
        ws = torch.empty(lib.cnf_occmap_tc_workspace_bytes(P), device=emb.device, dtype=torch.uint8)
        alpha = torch.empty(P, device=emb.device, dtype=torch.float32)
        rgb = torch.empty(P, 3, device=emb.device, dtype=torch.float32)
        check(lib.cnf_occmap_tc_fwd(P, ptr(params), ptr(emb), ptr(alpha), ptr(rgb), ptr(ws), ws.numel(), stream()))
        ctx.save_for_backward(params, rgb, ws)
        ctx.P = P
        return alpha, rgb

    @staticmethod
    def backward(ctx, g_alpha, g_rgb):
        params, rgb, ws = ctx.saved_tensors
        P = ctx.P
        lib = lib_for(params)
        g_alpha = _f32c(g_alpha) if g_alpha is not None else torch.zeros(P, device=params.device)
        g_rgb = _f32c(g_rgb) if g_rgb is not None else torch.zeros(P, 3, device=params.device)
        g_params = torch.empty_like(params)
        g_emb = torch.empty(P, EMB_DIM, device=params.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        check(lib.cnf_occmap_tc_bwd(P, ptr(params), ptr(rgb), ptr(g_alpha), ptr(g_rgb), ptr(ws), ws.numel(), ptr(g_params),
                                    ptr(g_emb), stream()))
        return g_emb, g_params


def occmap_eval_points_fused(points, flat_params, pe_B, pe_scale):
    """Forward-only tensor-core evaluation of the hidden-128 OccupancyMap (cnf_occmap_tc_eval_points): points [P,3] ->
    (occupancy [P] = sigmoid(alpha), rgb [P,3])."""
    lib = lib_for(points)
    P = points.shape[0]
    dev = points.device
    occ = torch.empty(P, device=dev, dtype=torch.float32)
    rgb = torch.empty(P, 3, device=dev, dtype=torch.float32)
    ws = torch.empty(lib.cnf_occmap_tc_eval_workspace_bytes(), device=dev, dtype=torch.uint8)
    check(lib.cnf_occmap_tc_eval_points(P, ptr(_f32c(flat_params.detach())), ptr(_f32c(pe_B.detach())), ptr(_f32c(pe_scale.detach().reshape(1))),
                                        ptr(_f32c(points)), ptr(occ), ptr(rgb), ptr(ws), ws.numel(), stream()))
    return occ, rgb


def occmap_render_rays_fused(origins, dirs, z, flat_params, pe_B, pe_scale, want_weights: bool = False):
    """Forward-only rendering through the hidden-128 OccupancyMap (cnf_occmap_tc_render_rays): origins/dirs [R,3], z [R,S] ->
    dict(depth [R], var [R], rgb [R,3], opacity [R][, weights [R,S]])."""
    lib = lib_for(origins)
    R, S = z.shape
    dev = origins.device
    out = torch.empty(R, 6, device=dev, dtype=torch.float32)
    w = torch.empty(R, S, device=dev, dtype=torch.float32) if want_weights else None
    ws = torch.empty(lib.cnf_occmap_tc_render_workspace_bytes(R, S), device=dev, dtype=torch.uint8)
    check(lib.cnf_occmap_tc_render_rays(R, S, ptr(_f32c(flat_params.detach())), ptr(_f32c(pe_B.detach())), ptr(_f32c(pe_scale.detach().reshape(1))),
                                        ptr(_f32c(origins)), ptr(_f32c(dirs)), ptr(_f32c(z)), ptr(out), ptr(w) if want_weights else None,
                                        ptr(ws), ws.numel(), stream()))
    res = {"depth": out[:, 0], "var": out[:, 1], "rgb": out[:, 2:5], "opacity": out[:, 5]}
    if want_weights:
        res["weights"] = w
    return res


def occmap_tc_points_fwd_bwd(points, params, pe_B, pe_scale, g_alpha_fn):
    """Tensor-core OccupancyMap with the embedding computed in-kernel (cnf_occmap_tc_fwd_points / _bwd_points), forward then
    backward: points [P,3], flat params (hidden 128), pe_B [21,3], pe_scale (0-d / [1]); ``g_alpha_fn(alpha, rgb) -> (g_alpha [P],
    g_rgb [P,3])`` supplies the upstream gradients.  Returns alpha, rgb, g_params, g_B."""
    lib = lib_for(points)
    points = _f32c(points); params = _f32c(params); pe_B = _f32c(pe_B); pe_scale = _f32c(pe_scale.reshape(1))
    P = points.shape[0]
    dev = points.device
    ws = torch.empty(lib.cnf_occmap_tc_workspace_bytes(P), device=dev, dtype=torch.uint8)
    alpha = torch.empty(P, device=dev, dtype=torch.float32)
    rgb = torch.empty(P, 3, device=dev, dtype=torch.float32)
    check(lib.cnf_occmap_tc_fwd_points(P, ptr(params), ptr(pe_B), ptr(pe_scale), ptr(points), ptr(alpha), ptr(rgb), ptr(ws), ws.numel(),
                                       0, stream()))
    g_alpha, g_rgb = g_alpha_fn(alpha, rgb)
    g_params = torch.empty_like(params)
    g_B = torch.empty_like(pe_B)
    check(lib.cnf_occmap_tc_bwd_points(P, ptr(params), ptr(points), ptr(rgb), ptr(_f32c(g_alpha)), ptr(_f32c(g_rgb)), ptr(ws), ws.numel(),
                                       ptr(g_params), ptr(g_B), 0, stream()))
    return alpha, rgb, g_params, g_B


# ------------------------------------------------------------------------------------------------
# compositing + losses
# ------------------------------------------------------------------------------------------------
def mask_counts(labels: torch.Tensor, mask_depth: torch.Tensor) -> torch.Tensor:
    lib = lib_for(labels)
    K, R = labels.shape
    counts = torch.empty(K, 4, device=labels.device, dtype=torch.int32)
    check(lib.cnf_mask_counts(K, R, ptr(labels), ptr(mask_depth), ptr(counts), stream()))
    return counts


class CompositeLossFn(torch.autograd.Function):
    """loss, out_loss[4+3K], rays[K,R,6] = step_batch_loss core (loss.py:18-74)."""

    @staticmethod
    def forward(ctx, alpha, color, z, gt_depth, gt_color, labels, mask_depth, cs, os_):
        lib = lib_for(alpha)
        alpha = _f32c(alpha); color = _f32c(color); z = _f32c(z)
        gt_depth = _f32c(gt_depth); gt_color = _f32c(gt_color)
        labels = labels.to(torch.uint8).contiguous()
        mask_depth = mask_depth.to(torch.uint8).contiguous()
        K, R, S = alpha.shape
        counts = mask_counts(labels, mask_depth)
        wsb = lib.cnf_composite_workspace_bytes(K, R)
        ws = torch.empty(wsb, device=alpha.device, dtype=torch.uint8)
        rays = torch.empty(K, R, 6, device=alpha.device, dtype=torch.float32)
        out = torch.empty(4 + 3 * K, device=alpha.device, dtype=torch.float32)
        check(lib.cnf_composite_loss_fwd(K, R, S, ptr(alpha), ptr(color), ptr(z), ptr(gt_depth), ptr(gt_color),
                                         ptr(labels), ptr(mask_depth), ptr(counts), float(cs), float(os_), ptr(rays),
                                         ptr(out), ptr(ws), wsb, stream()))
        ctx.save_for_backward(alpha, color, z, gt_depth, gt_color, labels, mask_depth, counts)
        ctx.meta = (K, R, S, float(cs), float(os_))
        ctx.mark_non_differentiable(out, rays)
        return out[0].clone(), out, rays

    @staticmethod
    def backward(ctx, g_loss, _g_out, _g_rays):
        alpha, color, z, gt_depth, gt_color, labels, mask_depth, counts = ctx.saved_tensors
        K, R, S, cs, os_ = ctx.meta
        lib = lib_for(alpha)
        g_loss = _f32c(g_loss).reshape(1)
        g_alpha = torch.empty_like(alpha)
        g_color = torch.empty_like(color)
        check(lib.cnf_composite_loss_bwd(K, R, S, ptr(alpha), ptr(color), ptr(z), ptr(gt_depth), ptr(gt_color),
                                         ptr(labels), ptr(mask_depth), ptr(counts), cs, os_, ptr(g_loss), ptr(g_alpha),
                                         ptr(g_color), stream()))
        return g_alpha, g_color, None, None, None, None, None, None, None


def termination(occ: torch.Tensor) -> torch.Tensor:
    """occupancy_to_termination forward (render_rays.py:25-33) on [..., S]."""
    lib = lib_for(occ)
    occ = _f32c(occ)
    S = occ.shape[-1]
    n = occ.numel() // S
    w = torch.empty_like(occ)
    check(lib.cnf_termination_fwd(n, S, ptr(occ), ptr(w), stream()))
    return w


def termination_bwd(occ: torch.Tensor, g_w: torch.Tensor) -> torch.Tensor:
    """Gradient of occupancy_to_termination w.r.t. the occupancies (cnf_termination_bwd)."""
    lib = lib_for(occ)
    occ = _f32c(occ); g_w = _f32c(g_w)
    S = occ.shape[-1]
    g_occ = torch.empty_like(occ)
    check(lib.cnf_termination_bwd(occ.numel() // S, S, ptr(occ), ptr(g_w), ptr(g_occ), stream()))
    return g_occ


class CodeRegFn(torch.autograd.Function):
    """sum_i ||codes_i||_2 (0 when n_obj == 1), loss.py:5-15."""

    @staticmethod
    def forward(ctx, codes):
        lib = lib_for(codes)
        codes = _f32c(codes)
        n_obj, L = codes.shape
        out = torch.empty(1, device=codes.device, dtype=torch.float32)
        check(lib.cnf_code_reg(n_obj, L, ptr(codes), 0.0, ptr(out), None, stream()))
        ctx.save_for_backward(codes)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (codes,) = ctx.saved_tensors
        n_obj = codes.shape[0]
        if n_obj <= 1:
            return torch.zeros_like(codes)
        nrm = codes.norm(dim=-1, keepdim=True)
        return g * codes / nrm.clamp_min(1e-30)


def adamw_step_(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step: int):
    lib = lib_for(param)
    check(lib.cnf_adamw_step(param.numel(), ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), float(lr), float(beta1),
                             float(beta2), float(eps), float(weight_decay), int(step), stream()))


def adamw_groups_(groups, beta1, beta2, eps, state):
    """cnf_adamw_groups: ``groups`` = [(param, grad, exp_avg, exp_avg_sq, lr, weight_decay), ...] (<= 8), all updated by ONE launch;
    ``state`` = int32[2] device tensor {steps done, 0}: the step count lives on the device, so the call is CUDA-graph capturable."""
    import ctypes
    from ._lib import AdamwGroup
    lib = lib_for(state)
    assert state.dtype == torch.int32 and state.numel() >= 2
    arr = (AdamwGroup * len(groups))()
    for i, (p, g, m, v, lr, wd) in enumerate(groups):
        assert p.numel() == g.numel() == m.numel() == v.numel()
        arr[i] = AdamwGroup(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(wd))
    check(lib.cnf_adamw_groups(len(groups), ctypes.cast(arr, ctypes.c_void_p), float(beta1), float(beta2), float(eps), ptr(state), stream()))


def eval_points_fused(points: torch.Tensor, flat_params: torch.Tensor, pe_B: torch.Tensor, pe_scale: torch.Tensor,
                      shape_code: torch.Tensor, tex_code: torch.Tensor):
    """Forward-only tensor-core evaluation (cnf_eval_points): points [P,3], one category's flat CodeNeRF(W=32) parameters,
    B [21,3], scale (0-d / [1]), one instance's codes [L] -> (occupancy [P] = sigmoid(10*sigma), rgb [P,3])."""
    lib = lib_for(points)
    P, L = points.shape[0], shape_code.numel()
    dev = points.device
    occ = torch.empty(P, device=dev, dtype=torch.float32)
    rgb = torch.empty(P, 3, device=dev, dtype=torch.float32)
    ws = torch.empty(lib.cnf_eval_points_workspace_bytes(), device=dev, dtype=torch.uint8)
    check(lib.cnf_eval_points(P, L, ptr(_f32c(flat_params.detach())), ptr(_f32c(pe_B.detach())), ptr(_f32c(pe_scale.detach().reshape(1))),
                              ptr(_f32c(shape_code.detach().reshape(-1))), ptr(_f32c(tex_code.detach().reshape(-1))), ptr(_f32c(points)),
                              ptr(occ), ptr(rgb), ptr(ws), ws.numel(), stream()))
    return occ, rgb


def render_rays_fused(origins: torch.Tensor, dirs: torch.Tensor, z: torch.Tensor, flat_params: torch.Tensor, pe_B: torch.Tensor,
                      pe_scale: torch.Tensor, shape_code: torch.Tensor, tex_code: torch.Tensor, want_weights: bool = False):
    """Forward-only rendering (cnf_render_rays): origins/dirs [R,3], z [R,S], one category's flat CodeNeRF(W=32) parameters and one
    instance's codes -> dict(depth [R], var [R], rgb [R,3], opacity [R][, weights [R,S] = termination probabilities])."""
    lib = lib_for(origins)
    R, S = z.shape
    L = shape_code.numel()
    dev = origins.device
    out = torch.empty(R, 6, device=dev, dtype=torch.float32)
    w = torch.empty(R, S, device=dev, dtype=torch.float32) if want_weights else None
    ws = torch.empty(lib.cnf_render_rays_workspace_bytes(R, S), device=dev, dtype=torch.uint8)
    check(lib.cnf_render_rays(R, S, L, ptr(_f32c(flat_params.detach())), ptr(_f32c(pe_B.detach())), ptr(_f32c(pe_scale.detach().reshape(1))),
                              ptr(_f32c(shape_code.detach().reshape(-1))), ptr(_f32c(tex_code.detach().reshape(-1))),
                              ptr(_f32c(origins)), ptr(_f32c(dirs)), ptr(_f32c(z)), ptr(out), ptr(w) if want_weights else None,
                              ptr(ws), ws.numel(), stream()))
    res = {"depth": out[:, 0], "var": out[:, 1], "rgb": out[:, 2:5], "opacity": out[:, 5]}
    if want_weights:
        res["weights"] = w
    return res

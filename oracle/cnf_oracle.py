"""CPU oracle for the neural-field training hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain PyTorch-on-CPU restatement of the reference
algorithm (Taekbum/category-nerf-reconstruction-official).  It may be imported by ``tests/``,
by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
leg, and by nothing else: the product package never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference's own modules, generated in the authoring
container by ``tests/golden/make_golden.py`` (which imports ``/root/reference/src``) and
committed under ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every
function below against those fixtures (bit-exact for the sampler and embedding, 1e-6 for
the float MLP / loss / gradients).

Every function cites the reference file:line it follows (paths relative to the reference
repository root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# constants fixed by the reference
# --------------------------------------------------------------------------------------
N_DIRS = 21            # icosahedron directions, src/embedding.py:51-73
EMB_SIZE1 = 21 * 4 + 3  # 87, src/trainer.py:20
EMB_SIZE2 = 21 * 6 + 3 - EMB_SIZE1  # 42, src/trainer.py:21

# src/embedding.py:51-73 (numeric table; data, not code)
ICOSA_DIRS = [
    0.8506508, 0, 0.5257311,
    0.809017, 0.5, 0.309017,
    0.5257311, 0.8506508, 0,
    1, 0, 0,
    0.809017, 0.5, -0.309017,
    0.8506508, 0, -0.5257311,
    0.309017, 0.809017, -0.5,
    0, 0.5257311, -0.8506508,
    0.5, 0.309017, -0.809017,
    0, 1, 0,
    -0.5257311, 0.8506508, 0,
    -0.309017, 0.809017, -0.5,
    0, 0.5257311, 0.8506508,
    -0.309017, 0.809017, 0.5,
    0.309017, 0.809017, 0.5,
    0.5, 0.309017, 0.809017,
    0.5, -0.309017, 0.809017,
    0, 0, 1,
    -0.5, 0.309017, 0.809017,
    -0.809017, 0.5, 0.309017,
    -0.809017, 0.5, -0.309017,
]


def icosa_matrix(dtype=torch.float32) -> torch.Tensor:
    """B_layer.weight initial value [21,3] (src/embedding.py:51-76)."""
    return torch.tensor(ICOSA_DIRS, dtype=dtype).reshape(-1, 3)


# parameter order of CodeNeRF.named_parameters() (src/model.py:36-54)
def codenerf_param_shapes(W: int = 32, L: int = 256) -> List[Tuple[str, Tuple[int, ...]]]:
    E1, E2 = EMB_SIZE1, EMB_SIZE2
    return [
        ("encoding_xyz.0.weight", (W, E1)), ("encoding_xyz.0.bias", (W,)),
        ("shape_latent_layer_1.0.weight", (W, L)), ("shape_latent_layer_1.0.bias", (W,)),
        ("shape_layer_1.0.weight", (W, W)), ("shape_layer_1.0.bias", (W,)),
        ("shape_latent_layer_2.0.weight", (W, L)), ("shape_latent_layer_2.0.bias", (W,)),
        ("shape_layer_2.0.weight", (W, W)), ("shape_layer_2.0.bias", (W,)),
        ("cat_layer.0.weight", (W, W + E1)), ("cat_layer.0.bias", (W,)),
        ("cat_latent_layer.0.weight", (W, L)), ("cat_latent_layer.0.bias", (W,)),
        ("encoding_shape.weight", (W, W)), ("encoding_shape.bias", (W,)),
        ("sigma.0.weight", (1, W)), ("sigma.0.bias", (1,)),
        ("encoding_viewdir.0.weight", (W, W + E2)), ("encoding_viewdir.0.bias", (W,)),
        ("texture_layer_1.0.weight", (W, W)), ("texture_layer_1.0.bias", (W,)),
        ("texture_latent_layer_1.0.weight", (W, L)), ("texture_latent_layer_1.0.bias", (W,)),
        ("rgb.0.weight", (W // 2, W)), ("rgb.0.bias", (W // 2,)),
        ("rgb.2.weight", (3, W // 2)), ("rgb.2.bias", (3,)),
    ]


# parameter order of OccupancyMap.named_parameters() (src/model.py:98-119)
def occmap_param_shapes(H: int = 128) -> List[Tuple[str, Tuple[int, ...]]]:
    E1, E2 = EMB_SIZE1, EMB_SIZE2
    return [
        ("in_layer.0.weight", (H, E1)), ("in_layer.0.bias", (H,)),
        ("mid1.0.0.weight", (H, H)), ("mid1.0.0.bias", (H,)),
        ("cat_layer.0.weight", (H, H + E1)), ("cat_layer.0.bias", (H,)),
        ("mid2.0.0.weight", (H, H)), ("mid2.0.0.bias", (H,)),
        ("out_alpha.weight", (1, H)), ("out_alpha.bias", (1,)),
        ("color_linear.0.weight", (H, E2 + H)), ("color_linear.0.bias", (H,)),
        ("out_color.weight", (3, H)), ("out_color.bias", (3,)),
    ]


def init_params(shapes, gen: torch.Generator, dtype=torch.float32) -> List[torch.Tensor]:
    """xavier-normal weights (src/model.py:4-6) and nn.Linear default uniform biases."""
    out = []
    fan_in = 1
    for name, shp in shapes:
        if len(shp) == 2:
            fan_out, fan_in = shp
            std = math.sqrt(2.0 / (fan_in + fan_out))
            out.append(torch.randn(shp, generator=gen, dtype=dtype) * std)
        else:
            bound = 1.0 / math.sqrt(fan_in)
            out.append((torch.rand(shp, generator=gen, dtype=dtype) * 2 - 1) * bound)
    return out


def init_codes(n_obj: int, L: int, gen: torch.Generator, dtype=torch.float32) -> torch.Tensor:
    """src/trainer.py:57-58: randn(n_obj, L) / sqrt(L/2)."""
    return torch.randn(n_obj, L, generator=gen, dtype=dtype) / math.sqrt(L / 2)


# --------------------------------------------------------------------------------------
# sampler  (src/scene_cateogries.py:51-96, 453-546)
# --------------------------------------------------------------------------------------
def stratified_bins(min_depth, max_depth, n_bins: int, u: torch.Tensor) -> torch.Tensor:
    """src/scene_cateogries.py:51-81 with the uniform draws ``u`` [n_rays, n_bins] injected.

    Every arithmetic step is a separately rounded fp32 op, as in eager PyTorch.
    """
    n_rays = u.shape[0]
    dt = u.dtype
    lin = torch.linspace(0, 1, n_bins + 1, dtype=dt)            # :54  (NOT k/n exactly)
    if not torch.is_tensor(min_depth):
        min_depth = torch.ones(n_rays, dtype=dt) * min_depth      # :56-57
    if not torch.is_tensor(max_depth):
        max_depth = torch.ones(n_rays, dtype=dt) * max_depth      # :59-60
    rng = max_depth - min_depth                                   # :62
    lower = (rng[..., None] * lin + min_depth[..., None])[:, :-1]  # :64-65
    bin_len = rng / n_bins                                        # :69
    return lower + u * bin_len[..., None]                         # :73-77


def normal_bins(depth: torch.Tensor, n: torch.Tensor, delta: float) -> torch.Tensor:
    """src/scene_cateogries.py:84-96 with the N(0, delta/3) draws ``n`` [n_rays, n_bins] injected."""
    bins = n.sort().values                                        # :90
    bins = torch.clip(bins, -delta, delta)                        # :91
    return depth[:, None] + bins                                  # :92


def sample_points(rgbs: torch.Tensor, depth: torch.Tensor, origins: torch.Tensor, dirs: torch.Tensor,
                  n_cam2surf: int, n_surf: int, eps: float, stop_eps: float,
                  rand: Dict[str, torch.Tensor], min_bound: float = 0.0, this_obj: int = 1, max_depth: Optional[float] = None):
    """sceneCategory.sample_3d_points (src/scene_cateogries.py:453-546).
    ``max_depth`` (not in the reference): upper bound of the invalid-depth rays when this batch is one part of a larger batch
    whose maximum depth (:486) has been computed over all parts (ray split across GPUs).

    ``rand`` holds the injected draws, consumed in the reference's order:
      'u_inv' [nI, N+M], 'u_val' [nV, N], 'n_obj' [nO, M] (already N(0, eps/3)), 'u_oth' [nT, M],
    where row k belongs to the k-th ray (in ray order) of the respective case.
    Returns (rgb u8 [R,3], depth, valid_mask bool, labels u8, pcs [R,S,3], z [R,S]).
    """
    R = rgbs.shape[0]
    N, M = n_cam2surf, n_surf
    z = torch.zeros(R, N + M, dtype=depth.dtype)                                   # :478-482
    invalid = (depth <= min_bound).view(-1)                                        # :484
    max_bound = torch.max(depth) if max_depth is None else torch.tensor(max_depth, dtype=depth.dtype)   # :486
    nI = int(invalid.count_nonzero())
    if nI:
        z[invalid, :] = stratified_bins(min_bound, max_bound, N + M, rand["u_inv"][:nI])   # :489-493
    valid = ~invalid
    nV = int(valid.count_nonzero())
    labels = rgbs[..., -1].view(-1)
    if nV:
        z[valid, :N] = stratified_bins(min_bound, depth.view(-1)[valid] - eps, N, rand["u_val"][:nV])  # :502-504
        obj = (labels == this_obj) & valid                                         # :507
        nO = int(obj.count_nonzero())
        if nO:
            z[obj, N:] = normal_bins(depth.view(-1)[obj], rand["n_obj"][:nO], eps)  # :519-524
        oth = (labels != this_obj) & valid                                          # :533
        nT = int(oth.count_nonzero())
        if nT:
            z[oth, N:] = stratified_bins(depth.view(-1)[oth] - eps, depth.view(-1)[oth] + stop_eps,
                                         M, rand["u_oth"][:nT])                      # :536-539
    pcs = origins[..., None, :] + (dirs[:, None, :] * z[..., None])                 # :543-544
    return rgbs[..., :3], depth, valid, labels, pcs, z


def draw_sampler_randoms(rgbs, depth, N, M, eps, gen: Optional[torch.Generator] = None,
                         min_bound=0.0, this_obj=1) -> Dict[str, torch.Tensor]:
    """Draw the four random tensors with the reference's shapes and order."""
    invalid = depth.view(-1) <= min_bound
    valid = ~invalid
    lab = rgbs[..., -1].view(-1)
    nI, nV = int(invalid.sum()), int(valid.sum())
    nO = int(((lab == this_obj) & valid).sum())
    nT = int(((lab != this_obj) & valid).sum())
    out = {}
    out["u_inv"] = torch.rand(nI, N + M, generator=gen) if nI else torch.zeros(0, N + M)
    out["u_val"] = torch.rand(nV, N, generator=gen) if nV else torch.zeros(0, N)
    if nV and nO:
        out["n_obj"] = torch.empty(nO, M).normal_(mean=0.0, std=eps / 3.0, generator=gen)
    else:
        out["n_obj"] = torch.zeros(0, M)
    out["u_oth"] = torch.rand(nT, M, generator=gen) if (nV and nT) else torch.zeros(0, M)
    return out


# --------------------------------------------------------------------------------------
# embedding  (src/embedding.py:82-92)
# --------------------------------------------------------------------------------------
def embed(x: torch.Tensor, B: torch.Tensor, scale, n_freq: int = 6) -> torch.Tensor:
    """UniDirsEmbed.forward: cat(x/scale, sin(pi * 2^f * (x/scale) B^T)), frequency-major."""
    scale_t = torch.as_tensor(scale, dtype=x.dtype)
    t = x / scale_t                                                        # :83
    if B.dim() == 3:      # stacked [K,21,3] with x [K,...,3]: what vmap(pe_model) computes
        proj = torch.matmul(t, B[(slice(None),) + (None,) * (x.dim() - 3)].transpose(-1, -2))
    else:
        proj = t @ B.t()                                                   # :84
    bands = 2.0 ** torch.linspace(0, n_freq - 1, n_freq, dtype=x.dtype)    # :78
    pb = proj[..., None, :] * bands[:, None]                               # :85
    xb = pb.reshape(list(proj.shape[:-1]) + [-1])                          # :86
    return torch.cat([t, torch.sin(xb * np.pi)], dim=-1)                  # :88-89


# --------------------------------------------------------------------------------------
# MLPs  (src/model.py:56-84, 124-154)
# --------------------------------------------------------------------------------------
def _lin(x, w, b):
    """y = x W^T + b.  With stacked parameters (w [K,out,in], b [K,out]) and x [K,...,in] this is the batched matmul
    functorch.vmap turns nn.Linear into (train.py:154-155)."""
    if w.dim() == 3:
        K = w.shape[0]
        y = torch.baddbmm(b[:, None, :], x.reshape(K, -1, x.shape[-1]), w.transpose(1, 2))    # one bmm per layer, as vmap emits
        return y.reshape(*x.shape[:-1], w.shape[1])
    return x @ w.transpose(-1, -2) + b


def codenerf_forward(P: Sequence[torch.Tensor], emb: torch.Tensor, shape_latent: torch.Tensor,
                     texture_latent: torch.Tensor):
    """CodeNeRF.forward with do_cat=True, noise_std=None (src/model.py:56-84).

    P: the 28 tensors in named_parameters() order.  emb [...,S,129]; latents [...,1,L].
    """
    (w_xyz, b_xyz, w_sl1, b_sl1, w_s1, b_s1, w_sl2, b_sl2, w_s2, b_s2, w_cat, b_cat, w_cl, b_cl,
     w_es, b_es, w_sig, b_sig, w_vd, b_vd, w_t1, b_t1, w_tl1, b_tl1, w_r1, b_r1, w_r2, b_r2) = P
    xg = emb[..., :EMB_SIZE1]
    xc = emb[..., EMB_SIZE1:]
    relu = torch.relu
    y = relu(_lin(xg, w_xyz, b_xyz))                                   # :59
    y = y + relu(_lin(shape_latent, w_sl1, b_sl1))                      # :66-67 (j=0)
    y = relu(_lin(y, w_s1, b_s1))                                       # :68
    y = y + relu(_lin(shape_latent, w_cl, b_cl))                        # :62-63
    y = relu(_lin(torch.cat((y, xg), dim=-1), w_cat, b_cat))           # :64-65
    y = y + relu(_lin(shape_latent, w_sl2, b_sl2))                      # :66-67 (j=1)
    y = relu(_lin(y, w_s2, b_s2))                                       # :68
    y = _lin(y, w_es, b_es)                                             # :70
    sigmas = _lin(y, w_sig, b_sig) * 10.0                               # :71,75
    h = relu(_lin(torch.cat((y, xc), dim=-1), w_vd, b_vd))             # :77-78
    h = h + relu(_lin(texture_latent, w_tl1, b_tl1))                    # :80-81
    h = relu(_lin(h, w_t1, b_t1))                                       # :82
    rgbs = torch.sigmoid(_lin(relu(_lin(h, w_r1, b_r1)), w_r2, b_r2))   # :54,83
    return sigmas, rgbs


def occmap_forward(P: Sequence[torch.Tensor], emb: torch.Tensor):
    """OccupancyMap.forward with defaults (src/model.py:124-154)."""
    (w_in, b_in, w_m1, b_m1, w_cat, b_cat, w_m2, b_m2, w_a, b_a, w_c, b_c, w_oc, b_oc) = P
    xg = emb[..., :EMB_SIZE1]
    xc = emb[..., EMB_SIZE1:]
    relu = torch.relu
    fc1 = relu(_lin(xg, w_in, b_in))                                    # :129
    fc2 = relu(_lin(fc1, w_m1, b_m1))                                   # :130
    fc3 = relu(_lin(torch.cat((fc2, xg), dim=-1), w_cat, b_cat))       # :133-134
    fc4 = relu(_lin(fc3, w_m2, b_m2))                                   # :137
    alpha = _lin(fc4, w_a, b_a) * 10.0                                  # :141,147
    h = relu(_lin(torch.cat((fc4, xc), dim=-1), w_c, b_c))             # :151
    color = torch.sigmoid(_lin(h, w_oc, b_oc))                          # :152-153
    return alpha, color


# --------------------------------------------------------------------------------------
# compositing and losses  (src/render_rays.py:3-95, src/loss.py:5-74)
# --------------------------------------------------------------------------------------
def termination(occ: torch.Tensor) -> torch.Tensor:
    """occupancy_to_termination (src/render_rays.py:25-33): w_i = occ_i * prod_{j<i}(1-occ_j+1e-10)."""
    free = (1.0 - occ + 1e-10)[..., :-1]
    first = torch.ones(list(occ.shape[:-1]) + [1], dtype=occ.dtype)
    T = torch.cumprod(torch.cat([first, free], dim=-1), dim=-1)
    return occ * T


def _reduce(loss_mat, mask, var=None):
    """reduce_batch_loss, avg=True, L1 (src/render_rays.py:66-95). Returns [K]."""
    mask_num = torch.sum(mask, dim=-1)
    if bool((mask_num == 0).any()):                                     # :68-72 (zero for ALL categories)
        return torch.zeros(loss_mat.shape[:-1], dtype=loss_mat.dtype)
    if var is not None:
        loss_mat = loss_mat * (1.0 / (torch.sqrt(var) + 1e-4))          # :73-80
    return torch.sum(loss_mat, dim=-1) / (mask_num + 1e-10)             # :86


def step_loss(alpha, color, gt_depth, gt_color, labels, mask_depth, z,
              color_scaling: float = 5.0, opacity_scaling: float = 10.0):
    """step_batch_loss (src/loss.py:18-74). alpha [K,R,S,1], color [K,R,S,3]; returns (loss, dict, loss_col)."""
    mask_obj = labels != 0                                              # :33
    mask_sem = labels != 2                                              # :35
    a = alpha.squeeze(-1)
    occ = torch.sigmoid(a)                                              # render_rays.py:3-7
    w = termination(occ)                                                # :42
    D = (w * z).sum(-1)                                                 # :44
    var = (w * (z - D[..., None]) ** 2).sum(-1).detach()                # :45-46
    C = (w[..., None] * color).sum(-2)                                  # :47
    O = w.sum(-1)                                                       # :48
    md = mask_depth & mask_obj
    l_d = _reduce(torch.abs(D - gt_depth) * md, md, var=var)            # :52-54
    l_c = _reduce(torch.abs(C - gt_color).sum(-1) * mask_obj, mask_obj)  # :58-60
    l_o = _reduce(torch.abs(O - mask_obj.to(O.dtype)) * mask_sem, mask_sem)  # :64-66
    l = l_d + l_c * color_scaling + l_o * opacity_scaling               # :69
    return l.sum(), {"depth": l_d, "color": l_c, "opacity": l_o}, l_c


def reg_loss(code_tables: Sequence[Tuple[torch.Tensor, torch.Tensor]]):
    """step_batch_loss_reg (src/loss.py:5-15): per category sum of per-instance L2 norms, 0 if n_obj==1."""
    rs, rt = [], []
    for sc, tc in code_tables:
        if sc.shape[0] > 1:
            rs.append(torch.norm(sc, dim=-1).sum())
            rt.append(torch.norm(tc, dim=-1).sum())
        else:
            rs.append(torch.zeros((), dtype=sc.dtype))
            rt.append(torch.zeros((), dtype=sc.dtype))
    return torch.stack(rs), torch.stack(rt)


def composite_closed_form(alpha, color, z, gt_depth, gt_color, labels, mask_depth,
                          color_scaling=5.0, opacity_scaling=10.0):
    """Closed-form forward/backward of compositing + loss for ONE category.

    alpha [R,S], color [R,S,3] (post-sigmoid), returns dict with D, var, C, O, loss terms,
    dL/dalpha [R,S], dL/dcolor [R,S,3]. Works in the dtype of the inputs (use float64 to check).
    """
    mo = labels != 0
    ms = labels != 2
    md = mask_depth & mo
    occ = torch.sigmoid(alpha)
    f = 1.0 - occ + 1e-10
    T = torch.cumprod(torch.cat([torch.ones_like(f[:, :1]), f[:, :-1]], -1), -1)
    w = occ * T
    D = (w * z).sum(-1)
    var = (w * (z - D[:, None]) ** 2).sum(-1)
    C = (w[..., None] * color).sum(-2)
    O = w.sum(-1)
    Nd, No, Ns = md.sum() + 1e-10, mo.sum() + 1e-10, ms.sum() + 1e-10
    zd = bool(md.sum() == 0); zo = bool(mo.sum() == 0); zs = bool(ms.sum() == 0)
    fz = alpha.new_zeros(())
    gD = fz * D if zd else torch.sign(D - gt_depth) * md / (torch.sqrt(var) + 1e-4) / Nd
    gC = fz * C if zo else color_scaling * torch.sign(C - gt_color) * mo[:, None] / No
    gO = fz * O if zs else opacity_scaling * torch.sign(O - mo.to(O.dtype)) * ms / Ns
    G = gD[:, None] * z + (gC[:, None, :] * color).sum(-1) + gO[:, None]
    Gw = G * w
    suffix = torch.flip(torch.cumsum(torch.flip(Gw, [-1]), -1), [-1]) - Gw   # sum_{i>j} G_i w_i
    docc = G * T - suffix / f
    dalpha = docc * occ * (1 - occ)
    dcolor = w[..., None] * gC[:, None, :]
    l_d = fz if zd else (torch.abs(D - gt_depth) * md / (torch.sqrt(var) + 1e-4)).sum() / Nd
    l_c = fz if zo else (torch.abs(C - gt_color).sum(-1) * mo).sum() / No
    l_o = fz if zs else (torch.abs(O - mo.to(O.dtype)) * ms).sum() / Ns
    return dict(D=D, var=var, C=C, O=O, w=w, depth=l_d, color=l_c, opacity=l_o,
                loss=l_d + color_scaling * l_c + opacity_scaling * l_o, dalpha=dalpha, dcolor=dcolor)


# --------------------------------------------------------------------------------------
# whole training step for K stacked categories (train.py:136-182)
# --------------------------------------------------------------------------------------
def train_step(fc_params: Sequence[torch.Tensor], pe_B: torch.Tensor, scale,
               shape_codes: Sequence[torch.Tensor], texture_codes: Sequence[torch.Tensor],
               pcs, z, gt_depth, gt_rgb_u8, labels, mask_depth, indices,
               reg_scaling: float = 0.0005, want_grads: bool = True):
    """One object-category training step, forward + backward, as train.py:136-182 runs it.

    fc_params: 28 tensors each stacked [K, ...]; pe_B [K,21,3]; shape/texture_codes: K tables
    [n_obj_k, L]; pcs [K,R,S,3]; z [K,R,S]; gt_depth [K,R]; gt_rgb_u8 [K,R,3]; labels u8 [K,R];
    mask_depth bool [K,R]; indices int64 [K,R].
    Returns dict(loss, depth, color, opacity, reg_shape, reg_texture, alpha, color_out, grads...).
    """
    K = pcs.shape[0]
    fc = [p.detach().clone().requires_grad_(want_grads) for p in fc_params]
    B = pe_B.detach().clone().requires_grad_(want_grads)
    sc = [c.detach().clone().requires_grad_(want_grads) for c in shape_codes]
    tc = [c.detach().clone().requires_grad_(want_grads) for c in texture_codes]
    sh = torch.stack([sc[k][indices[k]][:, None, :] for k in range(K)])    # train.py:136,148
    tx = torch.stack([tc[k][indices[k]][:, None, :] for k in range(K)])    # train.py:137,149
    emb = embed(pcs, B, scale)                                             # train.py:154 (vmap over K)
    alpha, color = codenerf_forward(fc, emb, sh, tx)                       # train.py:155 (vmap over K)
    gt_rgb = gt_rgb_u8.to(pcs.dtype) / 255.0                               # train.py:144
    loss, ld, lcol = step_loss(alpha, color, gt_depth, gt_rgb, labels, mask_depth, z)  # train.py:158-162
    rs, rt = reg_loss(list(zip(sc, tc)))                                   # train.py:166
    loss = loss + reg_scaling * (rs + rt).sum()                            # train.py:167
    out = dict(loss=loss.detach(), depth=ld["depth"].detach(), color=ld["color"].detach(),
               opacity=ld["opacity"].detach(), reg_shape=rs.detach(), reg_texture=rt.detach(),
               alpha=alpha.detach(), color_out=color.detach())
    if want_grads:
        loss.backward()                                                    # train.py:182
        z0 = lambda t: t.grad if t.grad is not None else torch.zeros_like(t)
        out["g_fc"] = [z0(p) for p in fc]
        out["g_B"] = z0(B)
        out["g_shape"] = [z0(c) for c in sc]
        out["g_texture"] = [z0(c) for c in tc]
    return out


def bg_train_step(fc_params: Sequence[torch.Tensor], pe_B: torch.Tensor, scale,
                  pcs, z, gt_depth, gt_rgb_u8, labels, mask_depth, want_grads=True):
    """Background (cls 0) step: un-vmapped OccupancyMap (train.py:172-180). Shapes [R,S,..]."""
    fc = [p.detach().clone().requires_grad_(want_grads) for p in fc_params]
    B = pe_B.detach().clone().requires_grad_(want_grads)
    emb = embed(pcs, B, scale)
    a, c = occmap_forward(fc, emb)
    gt_rgb = gt_rgb_u8.to(pcs.dtype) / 255.0
    loss, ld, lcol = step_loss(a[None], c[None], gt_depth[None], gt_rgb[None], labels[None],
                               mask_depth[None], z[None])
    out = dict(loss=loss.detach(), depth=ld["depth"].detach(), color=ld["color"].detach(),
               opacity=ld["opacity"].detach(), alpha=a.detach(), color_out=c.detach())
    if want_grads:
        loss.backward()
        out["g_fc"] = [p.grad for p in fc]
        out["g_B"] = B.grad
    return out


# --------------------------------------------------------------------------------------
# ray pools: per-step slicing, ray set-up and epoch reshuffle
# --------------------------------------------------------------------------------------
def origin_dirs(T: torch.Tensor, dirs_C: torch.Tensor, invert: bool):
    """src/scene_cateogries.py:24-35 (origin_dirs_O: invert the per-ray pose first) and :37-47 (origin_dirs_W)."""
    if invert:
        T = torch.linalg.inv(T)
    return T[:, :3, -1], (T[:, :3, :3] @ dirs_C[..., None]).squeeze(-1)


class RayPool:
    """One category's shuffled CPU ray pool as sceneCategory keeps it, with the batch slicing of get_training_samples.

    flat-pool branch (object categories), src/scene_cateogries.py:421-449: slice [i_batch, i_batch+n) of rgbs / depth / ray_dirs /
    per-ray poses / instance indices, origin_dirs_O (several instances) or origin_dirs_W (one instance); `i_batch += n`; when
    `i_batch >= P - n` every array is re-indexed by one `randperm(P)` and `i_batch = 0`.
    background branch, :353-418 with its single pool: the pose of a ray is `t_wc[frame]` (:373-379), same epoch rule (:404-412).
    ``rand_idx`` injects the permutation the reference draws with torch.randperm."""

    def __init__(self, rgbs, depth, ray_dirs, poses, indices=None, frames=None, invert=True):
        self.rgbs, self.depth, self.ray_dirs, self.poses = rgbs.clone(), depth.clone(), ray_dirs.clone(), poses.clone()
        self.indices = None if indices is None else indices.clone()
        self.frames = None if frames is None else frames.clone()
        self.invert = invert
        self.i_batch = 0

    def get_batch(self, n: int, rand_idx: Optional[torch.Tensor] = None):
        i = self.i_batch
        sl = slice(i, i + n)
        T = self.poses[self.frames[sl]] if self.frames is not None else self.poses[sl]
        origins, dirs = origin_dirs(T, self.ray_dirs[sl], self.invert)
        idx = self.indices[sl] if self.indices is not None else torch.zeros(n, dtype=torch.int64)
        out = (self.rgbs[sl], self.depth[sl], origins, dirs, idx)
        self.i_batch += n
        P = self.rgbs.shape[0]
        if self.i_batch >= P - n:
            if rand_idx is None:
                rand_idx = torch.randperm(P)
            self.rgbs, self.depth, self.ray_dirs = self.rgbs[rand_idx], self.depth[rand_idx], self.ray_dirs[rand_idx]
            if self.indices is not None:
                self.indices = self.indices[rand_idx]
            if self.frames is not None:
                self.frames = self.frames[rand_idx]
            else:
                self.poses = self.poses[rand_idx]
            self.i_batch = 0
        return out


def synth_pool(P: int, n_obj: int, n_frames: int, gen: torch.Generator):
    """A synthetic ray pool: rays of `synth_rays` plus sim3-shaped object->camera poses (rotation * scale | translation), one per
    (frame, instance), expanded per ray (t_co_batch_all) and also kept as a table with a per-ray frame index."""
    rgbs, depth, _, _, idx = synth_rays(P, n_obj, gen)
    u = torch.rand(P, generator=gen) * 1200; v = torch.rand(P, generator=gen) * 680
    ray_dirs = torch.stack([(u - 599.5) / 600.0, (v - 339.5) / 600.0, torch.ones(P)], -1)
    F = n_frames * n_obj
    q = torch.randn(F, 4, generator=gen); q = q / q.norm(dim=-1, keepdim=True)
    w_, x_, y_, z_ = q.unbind(-1)
    Rm = torch.stack([1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_),
                      2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_),
                      2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)], -1).reshape(F, 3, 3)
    s = torch.rand(F, 1, 1, generator=gen) * 0.7 + 0.3
    T = torch.zeros(F, 4, 4)
    T[:, :3, :3] = Rm * s
    T[:, :3, 3] = torch.randn(F, 3, generator=gen)
    T[:, 3, 3] = 1.0
    frame = torch.randint(0, n_frames, (P,), generator=gen)
    pose_index = frame * n_obj + idx
    return dict(rgbs=rgbs, depth=depth, ray_dirs=ray_dirs.contiguous(), indices=idx, pose_table=T, pose_index=pose_index,
                poses=T[pose_index].contiguous())


# --------------------------------------------------------------------------------------
# synthetic inputs (SURVEY 8(d) value distributions); shared by tests and bench
# --------------------------------------------------------------------------------------
def synth_rays(R: int, n_obj: int, gen: torch.Generator, invalid_frac: float = 0.15,
               label_mix=(0.60, 0.30, 0.10)):
    """One category's ray batch: rgbs u8 [R,4] (last col label 0/1/2 -> labels {1: this, 0: other, 2: unknown}),
    depth [R] (0 = invalid), origins [R,3], dirs [R,3], indices int64 [R]."""
    depth = torch.rand(R, generator=gen) * 3.0 + 0.5
    inval = torch.rand(R, generator=gen) < invalid_frac
    depth = torch.where(inval, torch.zeros_like(depth), depth)
    rgb = torch.randint(0, 256, (R, 3), generator=gen, dtype=torch.int64).to(torch.uint8)
    r = torch.rand(R, generator=gen)
    lab = torch.where(r < label_mix[0], 1, torch.where(r < label_mix[0] + label_mix[1], 0, 2)).to(torch.uint8)
    rgbs = torch.cat([rgb, lab[:, None]], dim=1)
    # camera rays ((u-cx)/fx, (v-cy)/fy, 1) rotated by a random rotation, scaled by 1/s (sim3)
    u = torch.rand(R, generator=gen) * 1200; v = torch.rand(R, generator=gen) * 680
    dc = torch.stack([(u - 599.5) / 600.0, (v - 339.5) / 600.0, torch.ones(R)], -1)
    q = torch.randn(4, generator=gen); q = q / q.norm()
    w_, x_, y_, z_ = q.tolist()
    Rm = torch.tensor([[1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_)],
                       [2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_)],
                       [2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)]])
    s = torch.rand(1, generator=gen).item() * 0.7 + 0.3
    dirs = (dc @ Rm.t()) / (s * 4.0)
    o = torch.randn(R, 3, generator=gen)
    origins = o / o.norm(dim=-1, keepdim=True) * (torch.rand(R, 1, generator=gen) * 0.5 + 0.5)
    idx = (torch.arange(R) % n_obj)[torch.randperm(R, generator=gen)].to(torch.int64)
    return rgbs, depth, origins.contiguous(), dirs.contiguous(), idx

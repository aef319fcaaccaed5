"""Depth-guided ray sampling -- drop-in for the sampling half of the reference's src/scene_cateogries.py
(stratified_bins :51-81, normal_bins_sampling :84-96, sceneCategory.sample_3d_points :453-546).

Extension for parity tests: the random draws can be injected (``u=``, ``n=``, ``rand=``); when
absent they are drawn on the device with torch.  With ``exact_rng=True`` the draws have exactly the
reference's shapes and order (this needs the four host syncs the reference also performs); the default draws
full-capacity [R, .] tensors without any host sync (same distribution, different RNG stream position)."""
from typing import Dict, Optional

import torch

from . import ops


def stratified_bins(min_depth, max_depth, n_bins, n_rays, type=torch.float32, device="cuda:0", z_fixed=False, u=None):
    """z [n_rays, n_bins]; min/max may be python floats, 0-d or [n_rays] tensors. z_fixed is ignored (as in the reference)."""
    if u is None:
        u = torch.rand(n_rays, n_bins, device=device, dtype=torch.float32)
    assert u.shape == (n_rays, n_bins)
    return ops.stratified_bins_op(min_depth, max_depth, n_bins, u)


def _origin_dirs(T, dirs_C, invert):
    assert T.shape[0] == dirs_C.shape[0] and T.shape[1:] == (4, 4) and dirs_C.shape[1] == 3
    from ._lib import check, lib_for, ptr, stream
    lib = lib_for(T)
    T = T.float().contiguous(); dirs_C = dirs_C.float().contiguous()
    n = T.shape[0]
    origins = torch.empty(n, 3, device=T.device, dtype=torch.float32)
    dirs = torch.empty(n, 3, device=T.device, dtype=torch.float32)
    check(lib.cnf_origin_dirs(n, int(invert), ptr(T), ptr(dirs_C), ptr(origins), ptr(dirs), stream()))
    return origins, dirs


def origin_dirs_O(T_CO, dirs_C):
    """Ray origins / directions in the object frame from per-ray object->camera poses (scene_cateogries.py:24-35):
    T_OC = inv(T_CO); dirs_O = T_OC[:, :3, :3] @ dirs_C; origins = T_OC[:, :3, 3]."""
    return _origin_dirs(T_CO, dirs_C, True)


def origin_dirs_W(T_WC, dirs_C):
    """Ray origins / directions in the world frame (scene_cateogries.py:37-47)."""
    return _origin_dirs(T_WC, dirs_C, False)


def normal_bins_sampling(depth, n_bins, n_rays, delta, device="cuda:0", n=None):
    """depth[:,None] + clip(sort(N(0, delta/3)), -delta, delta)  ->  [n_rays, n_bins]."""
    if n is None:
        n = torch.empty(n_rays, n_bins, dtype=torch.float32, device=device).normal_(mean=0., std=delta / 3.)
    dev = n.device
    depth = depth.to(dev).float().reshape(1, n_rays)
    rgbs = torch.ones(1, n_rays, 4, dtype=torch.uint8, device=dev)            # every ray: label 1 = this object
    zero3 = torch.zeros(1, n_rays, 3, device=dev)
    dummy = torch.zeros(1, n_rays, n_bins, device=dev)
    rand = {"u_inv": dummy, "u_val": dummy[..., :0], "n_obj": n.reshape(1, n_rays, n_bins), "u_oth": dummy}
    z, _, _, _, _ = ops.sample_batched(rgbs, depth, zero3, zero3, 0, n_bins, delta, delta, rand,
                                       min_bound=float("-inf"), this_obj=1, want_pcs=False)
    return z[0]


def draw_randoms(rgbs, depth, n_cam2surf, n_surf, eps, min_bound=0.0, this_obj=1, exact_rng=False) -> Dict[str, torch.Tensor]:
    """Draw the sampler's random tensors for K stacked categories ([K,R,...])."""
    K, R = depth.shape
    N, M = n_cam2surf, n_surf
    dev = depth.device
    if not exact_rng:
        # all uniforms in ONE launch (contiguous slices of a flat draw), the normals in a second one
        n1, n2, n3 = K * R * (N + M), K * R * N, K * R * M
        u = torch.rand(n1 + n2 + n3, device=dev)
        return {"u_inv": u[:n1].view(K, R, N + M), "u_val": u[n1:n1 + n2].view(K, R, N),
                "n_obj": torch.empty(K, R, M, device=dev).normal_(mean=0., std=eps / 3.),
                "u_oth": u[n1 + n2:].view(K, R, M)}
    out = {"u_inv": torch.zeros(K, R, N + M, device=dev), "u_val": torch.zeros(K, R, N, device=dev),
           "n_obj": torch.zeros(K, R, M, device=dev), "u_oth": torch.zeros(K, R, M, device=dev)}
    for k in range(K):          # reference order per category: rand[nI,S], rand[nV,N], normal_[nO,M], rand[nT,M]
        inval = depth[k] <= min_bound
        lab = rgbs[k, :, 3]
        nI = int(inval.sum()); nV = R - nI
        nO = int(((lab == this_obj) & ~inval).sum()); nT = nV - nO
        if nI:
            out["u_inv"][k, :nI] = torch.rand(nI, N + M, device=dev)
        if nV:
            out["u_val"][k, :nV] = torch.rand(nV, N, device=dev)
            if nO:
                out["n_obj"][k, :nO] = torch.empty(nO, M, device=dev).normal_(mean=0., std=eps / 3.)
            if nT:
                out["u_oth"][k, :nT] = torch.rand(nT, M, device=dev)
    return out


def _pad_rand(rand, R, N, M, dev):
    """Accept the reference-shaped compact draws ([nI,S], [nV,N], [nO,M], [nT,M]) and pad rows to capacity R."""
    shapes = {"u_inv": N + M, "u_val": N, "n_obj": M, "u_oth": M}
    out = {}
    for k, c in shapes.items():
        t = rand.get(k)
        buf = torch.zeros(1, R, c, device=dev, dtype=torch.float32)
        if t is not None and t.numel():
            t = t.to(dev).float().reshape(-1, c)
            buf[0, :t.shape[0]] = t
        out[k] = buf
    return out


def sample_3d_points(sampled_rgbs, sampled_depth, origins, dirs_o, n_bins_cam2surface, n_bins, surface_eps, stop_eps,
                     min_bound=0.0, this_obj=1, rand: Optional[Dict[str, torch.Tensor]] = None, exact_rng=False):
    """Functional form of sceneCategory.sample_3d_points for one category.
    Returns (rgb u8 [R,3], depth [R], valid_mask bool [R], labels u8 [R], pcs [R,S,3], z [R,S])."""
    R = sampled_rgbs.shape[0]
    dev = sampled_depth.device
    rg = sampled_rgbs.reshape(1, R, 4)
    dp = sampled_depth.reshape(1, R)
    if rand is None:
        r = draw_randoms(rg, dp, n_bins_cam2surface, n_bins, surface_eps, min_bound, this_obj, exact_rng)
    else:
        r = _pad_rand(rand, R, n_bins_cam2surface, n_bins, dev)
    z, pcs, valid, labels, _ = ops.sample_batched(rg, dp, origins.reshape(1, R, 3), dirs_o.reshape(1, R, 3),
                                                  n_bins_cam2surface, n_bins, surface_eps, stop_eps, r, min_bound,
                                                  this_obj)
    return sampled_rgbs[..., :3], sampled_depth, valid[0], labels[0], pcs[0], z[0]


class RaySampler:
    """The sampling-related attributes of the reference's sceneCategory plus its sample_3d_points method
    (scene_cateogries.py:453), so callers written against ``scene.sample_3d_points(rgbs, depth, origins, dirs)`` work."""

    def __init__(self, n_bins_cam2surface, n_bins, surface_eps, stop_eps, data_device="cuda:0", min_bound=0.0, this_obj=1):
        self.n_bins_cam2surface = n_bins_cam2surface
        self.n_bins = n_bins
        self.surface_eps = surface_eps
        self.stop_eps = stop_eps
        self.data_device = data_device
        self.min_bound = min_bound
        self.this_obj = this_obj

    def sample_3d_points(self, sampled_rgbs, sampled_depth, origins, dirs_o, rand=None, exact_rng=False):
        return sample_3d_points(sampled_rgbs, sampled_depth, origins, dirs_o, self.n_bins_cam2surface, self.n_bins,
                                self.surface_eps, self.stop_eps, self.min_bound, self.this_obj, rand, exact_rng)

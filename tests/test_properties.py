"""Property tests (hypothesis) of the hot path's size-independent invariants.

CPU (oracle): sampler bounds / sortedness / case bookkeeping on random case mixes, termination weights are a sub-probability,
ray pools serve every pool row exactly once per epoch and i_batch follows the reference's rule.
GPU: the CUDA sampler equals the oracle bit for bit on random shapes (rays, bin counts, invalid fraction, label mix, categories),
the compositing kernels reproduce the oracle's loss on random shapes, pool gather == oracle pool on random sizes."""
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import cnf_oracle as O

DEV = "cuda:0"


def _case(seed, R, N, M, inv, p_this, p_other):
    g = torch.Generator().manual_seed(seed)
    rgbs, depth, org, dirs, _ = O.synth_rays(R, 3, g, invalid_frac=inv, label_mix=(p_this, p_other, max(0.0, 1.0 - p_this - p_other)))
    rand = {"u_inv": torch.rand(R, N + M, generator=g), "u_val": torch.rand(R, N, generator=g),
            "n_obj": torch.randn(R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(R, M, generator=g)}
    return rgbs, depth, org, dirs, rand


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10 ** 6), R=st.integers(1, 300), N=st.integers(0, 6), M=st.integers(1, 12),
       inv=st.floats(0.0, 1.0), p_this=st.floats(0.0, 1.0), p_other=st.floats(0.0, 1.0))
def test_oracle_sampler_invariants(seed, R, N, M, inv, p_this, p_other):
    p_other = min(p_other, 1.0 - p_this)
    rgbs, depth, org, dirs, rand = _case(seed, R, N, M, inv, p_this, p_other)
    eps, stop = 0.1, 0.05
    _, _, valid, labels, pcs, z = O.sample_points(rgbs, depth, org, dirs, N, M, eps, stop, rand)
    assert torch.equal(valid, depth > 0)
    d = depth[:, None]
    this = valid & (labels == 1)
    oth = valid & (labels != 1)
    if this.any():                                   # sorted normals inside the clip band around the surface (:519-524)
        zs = z[this][:, N:]
        assert torch.all(zs[:, 1:] >= zs[:, :-1])
        assert torch.all((zs - d[this]).abs() <= eps * (1 + 1e-6) + 1e-6)
    if oth.any():                                    # stratified band [d - eps, d + stop_eps] (:536-539), increasing bins
        zo = z[oth][:, N:]
        assert torch.all(zo >= d[oth] - eps - 1e-5) and torch.all(zo <= d[oth] + stop + 1e-5)
        assert torch.all(zo[:, 1:] >= zo[:, :-1] - 1e-6)
    if valid.any() and N:                            # camera -> surface - eps (:502-504)
        zc = z[valid][:, :N]
        assert torch.all(zc <= (d[valid] - eps).clamp(min=0) + 1e-5) or torch.all(d[valid] - eps < 0)
    if (~valid).any():                               # invalid depth: bins up to the batch maximum (:484-493)
        zi = z[~valid]
        assert torch.all(zi >= -1e-6) and torch.all(zi <= float(depth.max()) + 1e-5)
    torch.testing.assert_close(pcs, org[:, None, :] + dirs[:, None, :] * z[..., None], rtol=0, atol=0)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 50), S=st.integers(1, 120))
def test_oracle_termination_is_sub_probability(seed, n, S):
    g = torch.Generator().manual_seed(seed)
    occ = torch.sigmoid(torch.randn(n, S, generator=g) * 4)
    w = O.termination(occ)
    assert torch.all(w >= 0) and torch.all(w.sum(-1) <= 1 + 1e-5)
    # re-composition: 1 - sum w = prod (1 - occ + 1e-10) over all but ... the last factor is never used (render_rays.py:28)
    T_last = torch.cumprod(torch.cat([torch.ones(n, 1), (1 - occ + 1e-10)[:, :-1]], -1), -1)[:, -1]
    torch.testing.assert_close(w[:, -1], occ[:, -1] * T_last, rtol=1e-5, atol=1e-7)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), P=st.integers(40, 400), n=st.integers(1, 39))
def test_oracle_pool_serves_each_row_once_per_epoch(seed, P, n):
    g = torch.Generator().manual_seed(seed)
    pool = O.synth_pool(P, 3, 4, g)
    tag = torch.arange(P, dtype=torch.float32)       # depth doubles as a row tag
    rp = O.RayPool(pool["rgbs"], tag, pool["ray_dirs"], pool["poses"], pool["indices"], invert=True)
    seen = []
    for _ in range(3 * (P // n) + 3):
        before = rp.i_batch
        _, d, _, _, _ = rp.get_batch(n, torch.randperm(P, generator=g))
        seen.append(d)
        assert rp.i_batch == (0 if before + n >= P - n else before + n)          # scene_cateogries.py:437-449
        if rp.i_batch == 0:                          # epoch ended: the rows served in it were distinct
            ep = torch.cat(seen); seen = []
            assert ep.unique().numel() == ep.numel()


# ------------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), K=st.integers(1, 3), R=st.integers(1, 9000), N=st.integers(0, 5), M=st.integers(1, 18),
       inv=st.floats(0.0, 1.0), p_this=st.floats(0.0, 1.0))
def test_gpu_sampler_equals_oracle_on_random_shapes(seed, K, R, N, M, inv, p_this):
    from category_nerf_reconstruction_official_b200 import ops
    cases = [_case(seed + k, R, N, M, inv, p_this, (1.0 - p_this) * 0.7) for k in range(K)]
    rgbs, depth, org, dirs = (torch.stack([c[i] for c in cases]) for i in range(4))
    rand = {k: torch.stack([c[4][k] for c in cases]) for k in cases[0][4]}
    z, pcs, valid, labels, stats = ops.sample_batched(rgbs.to(DEV), depth.to(DEV), org.to(DEV), dirs.to(DEV), N, M, 0.1, 0.05,
                                                      {k: v.to(DEV) for k, v in rand.items()})
    for k in range(K):
        _, _, v_o, l_o, pcs_o, z_o = O.sample_points(rgbs[k], depth[k], org[k], dirs[k], N, M, 0.1, 0.05, {kk: vv[k] for kk, vv in rand.items()})
        assert torch.equal(z[k].cpu(), z_o) and torch.equal(pcs[k].cpu(), pcs_o) and torch.equal(valid[k].cpu(), v_o)
        assert int(stats[k, 0]) == int((~v_o).sum()) and int(stats[k, 2]) == int((v_o & (l_o == 1)).sum())


@pytest.mark.gpu
@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10 ** 6), K=st.integers(1, 3), R=st.integers(1, 1500), S=st.integers(1, 40))
def test_gpu_composite_loss_equals_oracle_on_random_shapes(seed, K, R, S):
    from category_nerf_reconstruction_official_b200 import loss
    g = torch.Generator().manual_seed(seed)
    alpha = torch.randn(K, R, S, 1, generator=g) * 3
    color = torch.rand(K, R, S, 3, generator=g)
    z = torch.sort(torch.rand(K, R, S, generator=g) * 3 + 0.2, dim=-1).values
    gd = torch.rand(K, R, generator=g) * 3 + 0.3; gc = torch.rand(K, R, 3, generator=g)
    lab = torch.randint(0, 3, (K, R), generator=g).to(torch.uint8); md = torch.rand(K, R, generator=g) > 0.2
    lo, ldo, _ = O.step_loss(alpha, color, gd, gc, lab, md, z)
    l1, ld1, _ = loss.step_batch_loss(alpha.to(DEV), color.to(DEV), gd.to(DEV), gc.to(DEV), lab.to(DEV), md.to(DEV), z.to(DEV))
    torch.testing.assert_close(l1.cpu(), lo, rtol=3e-5, atol=1e-5)
    for kk in ("depth", "color", "opacity"):
        torch.testing.assert_close(ld1[kk].cpu(), ldo[kk], rtol=5e-5, atol=1e-5)


@pytest.mark.gpu
@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 10 ** 6), P=st.integers(50, 3000), n=st.integers(1, 49))
def test_gpu_pool_equals_oracle_pool_on_random_sizes(seed, P, n):
    from category_nerf_reconstruction_official_b200.pool import DeviceRayPools
    g = torch.Generator().manual_seed(seed)
    pool = O.synth_pool(P, 4, 3, g)
    dp = DeviceRayPools([dict(rgbs=pool["rgbs"], depth=pool["depth"], ray_dirs=pool["ray_dirs"], poses=pool["pose_table"],
                              pose_index=pool["pose_index"], indices=pool["indices"])], invert=True, device=DEV)
    rp = O.RayPool(pool["rgbs"], pool["depth"], pool["ray_dirs"], pool["pose_table"], pool["indices"], frames=pool["pose_index"], invert=True)
    for _ in range(2 * (P // n) + 2):
        shuf = rp.i_batch + n >= P - n
        perm = torch.randperm(P, generator=g) if shuf else None
        r, d, og, di, ix = rp.get_batch(n, perm)
        gr, gd_, go, gdi, gi = [v.cpu() for v in dp.next_batch_of(0, n, perm)]
        assert torch.equal(gr, r) and torch.equal(gd_, d) and torch.equal(gi, ix)
        torch.testing.assert_close(go, og, rtol=2e-5, atol=2e-6)
        torch.testing.assert_close(gdi, di, rtol=2e-5, atol=2e-6)

"""GPU parity of the HBM-bound standalone kernels at the sizes where their large-batch code paths run (round 2 rewrites):
  * sampler: count -> (no scan launch) -> thread-per-ray emit (cnf_sampler.cu: sample_emit_fast_kernel), several count CTAs per
    category, ragged last ray block, surface-bin counts 7 / 9 / 12 / 16, the all-reduced max-depth override, and the 17-bin generic
    path with the scan launch -- all bit-exact against the oracle (src/scene_cateogries.py:453-546);
  * compositing + losses: the dense double-buffered kernel (bulk copies, several tiles per CTA, ragged last tile, categories whose
    base offset is not 16-byte aligned) for S = 10 / 14 / 12 / 7 and the padded kernel (S = 8 / 16) against the oracle's
    autograd (src/render_rays.py:25-95, src/loss.py:18-74) -- loss 1e-5, gradients 2e-4."""
import pytest
import torch

from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("K,R,N,M,override", [(1, 70001, 1, 9, False), (2, 33000, 5, 9, False), (1, 20000, 1, 7, False), (1, 9000, 0, 12, False),
                                              (2, 8200, 3, 16, True), (1, 12000, 1, 17, False)])
def test_sampler_large_paths_bit_exact(K, R, N, M, override):
    from category_nerf_reconstruction_official_b200 import ops
    g = torch.Generator().manual_seed(R + M)
    cats = [O.synth_rays(R, 4, g, invalid_frac=0.15 if k == 0 else 0.4) for k in range(K)]
    rgbs, depth, org, dirs = (torch.stack([c[i] for c in cats]) for i in range(4))
    rand = {"u_inv": torch.rand(K, R, N + M, generator=g), "u_val": torch.rand(K, R, N, generator=g),
            "n_obj": torch.randn(K, R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(K, R, M, generator=g)}
    md = (depth.max(dim=1).values + torch.arange(1, K + 1)) if override else None
    z, pcs, valid, labels, stats = ops.sample_batched(rgbs.to(DEV), depth.to(DEV), org.to(DEV), dirs.to(DEV), N, M, 0.1, 0.05,
                                                      {k: v.to(DEV) for k, v in rand.items()},
                                                      max_depth=None if md is None else md.to(DEV))
    for k in range(K):
        _, _, v_o, l_o, pcs_o, z_o = O.sample_points(rgbs[k], depth[k], org[k], dirs[k], N, M, 0.1, 0.05, {kk: vv[k] for kk, vv in rand.items()},
                                                     **({} if md is None else {"max_depth": float(md[k])}))
        assert torch.equal(z[k].cpu(), z_o), "z differs (category %d)" % k
        assert torch.equal(pcs[k].cpu(), pcs_o)
        assert torch.equal(valid[k].cpu(), v_o) and torch.equal(labels[k].cpu(), l_o)
        lab = rgbs[k, :, 3]
        st = stats[k].cpu()
        assert [int(v) for v in st[:7]] == [int((~v_o).sum()), int(v_o.sum()), int((v_o & (lab == 1)).sum()), int((v_o & (lab != 1)).sum()),
                                             int((v_o & (lab != 0)).sum()), int((lab != 0).sum()), int((lab != 2).sum())]
        want_max = float(md[k]) if md is not None else float(depth[k].max())
        assert st[7:8].view(torch.float32).item() == want_max
    # without pcs (the fused trainer's call)
    z2, pcs2, *_ = ops.sample_batched(rgbs.to(DEV), depth.to(DEV), org.to(DEV), dirs.to(DEV), N, M, 0.1, 0.05,
                                      {k: v.to(DEV) for k, v in rand.items()}, want_pcs=False, max_depth=None if md is None else md.to(DEV))
    assert pcs2 is None and torch.equal(z2, z)


@pytest.mark.parametrize("K,R,S", [(1, 40000, 10), (3, 1337, 14), (2, 5001, 12), (2, 999, 7), (1, 3000, 8), (2, 700, 16), (1, 170000, 10)])
def test_composite_loss_large_vs_oracle(K, R, S):
    from category_nerf_reconstruction_official_b200 import loss
    g = torch.Generator().manual_seed(S * 7 + K)
    alpha = torch.randn(K, R, S, 1, generator=g) * 3
    color = torch.rand(K, R, S, 3, generator=g)
    z = torch.sort(torch.rand(K, R, S, generator=g) * 3 + 0.2, dim=-1).values
    gd = torch.rand(K, R, generator=g) * 3 + 0.3; gc = torch.rand(K, R, 3, generator=g)
    lab = torch.randint(0, 3, (K, R), generator=g).to(torch.uint8); md = torch.rand(K, R, generator=g) > 0.2
    a0 = alpha.clone().requires_grad_(True); c0 = color.clone().requires_grad_(True)
    lo, ldo, _ = O.step_loss(a0, c0, gd, gc, lab, md, z); lo.backward()
    a1 = alpha.to(DEV).requires_grad_(True); c1 = color.to(DEV).requires_grad_(True)
    l1, ld1, _ = loss.step_batch_loss(a1, c1, gd.to(DEV), gc.to(DEV), lab.to(DEV), md.to(DEV), z.to(DEV)); l1.backward()
    torch.testing.assert_close(l1.cpu(), lo.detach(), rtol=1e-5, atol=1e-6)
    for kk in ("depth", "color", "opacity"):
        torch.testing.assert_close(ld1[kk].cpu(), ldo[kk].detach(), rtol=2e-5, atol=1e-6)
    # gradients: every entry within 2e-4 of the tensor's scale (the depth term's 1/(sqrt(var)+1e-4) weighting makes single rays huge)
    for got, want in ((a1.grad.cpu(), a0.grad), (c1.grad.cpu(), c0.grad)):
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 2e-4 * scale + 1e-9, (float((got - want).abs().max()), scale)


@pytest.mark.parametrize("n,S", [(10000, 96), (77, 14), (1, 1)])
def test_ray_uncertainty_fields(n, S):
    """Entropy / termination mass / heuristic of category_registration.py:152-162 (numpy formula of the reference, float64) from
    termination weights produced by occupancy_to_termination."""
    import numpy as np
    from category_nerf_reconstruction_official_b200 import render_rays
    g = torch.Generator().manual_seed(n + S)
    occ = torch.sigmoid(torch.randn(n, S, generator=g) * 3 - 1)
    term = render_rays.occupancy_to_termination(occ.to(DEV))
    ent, mass, heur = (v.cpu().numpy() for v in render_rays.ray_uncertainty(term))
    tp = O.termination(occ).double().numpy()                                  # the reference's weights (src/render_rays.py:25-33)
    ent_ref = np.sum(-tp * np.log(tp + 1e-10), axis=-1)
    mass_ref = np.sum(tp, axis=-1)
    np.testing.assert_allclose(ent, ent_ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(mass, mass_ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(heur, mass_ref * np.exp(-0.5 * ent_ref), rtol=3e-5, atol=2e-6)

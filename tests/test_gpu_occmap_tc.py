"""Tensor-core OccupancyMap (hidden size 128, the reference's background model; src/model.py:86-154, train.py:172-180).

Two references:
  * a torch fp32 restatement with the kernel's operand roundings made explicit (operands of every GEMM fp16, accumulation fp32):
    the kernel must agree with it to accumulation-order noise;
  * the fp32 oracle (oracle/cnf_oracle.py): what the reference computes.  ReLU masks flip for pre-activations within the fp16
    forward error of zero, which bounds the agreement of the gradients.
"""
import pytest
import torch

from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf(x):
    """Round to the kernels' GEMM operand type (IEEE fp16 by default: cnf_umma.cuh CNF_OPERAND_FP16; the gradient operands carry an
    exact power-of-two scale per tile, which commutes with the rounding)."""
    return x.to(torch.float16).to(torch.float32)


def occmap_bf16_emulation(params, emb, g_alpha, g_rgb):
    Wi, bi, W1, b1, Wc, bc, W2, b2, wa, ba, Wl, bl, Wo, bo = [p.detach().float() for p in params]
    xg = bf(emb[:, :87]); xc = bf(emb[:, 87:])
    fc1 = bf(torch.relu(xg @ bf(Wi).T + bf(bi)))
    fc2 = bf(torch.relu(fc1 @ bf(W1).T + b1))
    fc3 = bf(torch.relu(fc2 @ bf(Wc[:, :128]).T + xg @ bf(Wc[:, 128:]).T + bf(bc)))
    pre4 = torch.relu(fc3 @ bf(W2).T + b2); fc4 = bf(pre4)
    alpha = 10.0 * (pre4 @ wa.T + ba).reshape(-1)
    preh = torch.relu(fc4 @ bf(Wl[:, :128]).T + xc @ bf(Wl[:, 128:]).T + bf(bl)); h = bf(preh)
    rgb = torch.sigmoid(preh @ Wo.T + bo)
    dp = g_rgb * rgb * (1 - rgb); draw = 10.0 * g_alpha
    hd = bf(torch.cat([dp, draw[:, None]], 1))
    dh = bf((dp @ Wo) * (h > 0))
    g_Wo = hd[:, :3].T @ h; g_bo = dp.sum(0)
    g_wa = hd[:, 3:4].T @ fc4; g_ba = draw.sum().reshape(1)
    g_Wl = torch.cat([dh.T @ fc4, dh.T @ xc], 1); g_bl = dh.sum(0)
    d4 = bf((dh @ bf(Wl[:, :128]) + draw[:, None] * wa) * (fc4 > 0)); g_xc = dh @ bf(Wl[:, 128:])
    g_W2 = d4.T @ fc3; g_b2 = d4.sum(0)
    d3 = bf((d4 @ bf(W2)) * (fc3 > 0))
    g_Wc = torch.cat([d3.T @ fc2, d3.T @ xg], 1); g_bc = d3.sum(0)
    d2 = bf((d3 @ bf(Wc[:, :128])) * (fc2 > 0)); g_xg = d3 @ bf(Wc[:, 128:])
    g_W1 = d2.T @ fc1; g_b1 = d2.sum(0)
    d1 = bf((d2 @ bf(W1)) * (fc1 > 0))
    g_Wi = d1.T @ xg; g_bi = d1.sum(0)
    g_xg = g_xg + d1 @ bf(Wi)
    gp = torch.cat([t.reshape(-1) for t in (g_Wi, g_bi, g_W1, g_b1, g_Wc, g_bc, g_W2, g_b2, g_wa, g_ba, g_Wl, g_bl, g_Wo, g_bo)])
    return alpha, rgb, gp, torch.cat([g_xg, g_xc], 1)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("P", [1, 127, 128, 129, 1200 * 14, 148 * 128 + 5, 3 * 148 * 128 + 77])
def test_occmap_tc_fwd_bwd(P):
    """Ragged point counts: one partial tile, exactly one tile, the reference's 1200 x 14 background batch (one wave), more tiles
    than SMs (per-CTA partials accumulate over several tiles)."""
    from category_nerf_reconstruction_official_b200 import ops
    g = torch.Generator().manual_seed(400 + P % 97)
    params = O.init_params(O.occmap_param_shapes(128), g)
    B = O.icosa_matrix() + 0.01 * torch.randn(21, 3, generator=g)
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 6.0
    emb = O.embed(pts[:, None, :], B, 5.0).reshape(P, 129)
    ga = torch.randn(P, generator=g) * 0.1; gc = torch.randn(P, 3, generator=g)
    flat = torch.cat([p.reshape(-1) for p in params]).to(DEV).requires_grad_(True)
    emb_d = emb.to(DEV).requires_grad_(True)
    alpha, rgb = ops.OccMapTCFn.apply(emb_d, flat)
    ((alpha * ga.to(DEV)).sum() + (rgb * gc.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    # (1) bf16 emulation
    a_e, c_e, gp_e, ge_e = occmap_bf16_emulation(params, emb, ga, gc)
    assert float((alpha.detach().cpu() - a_e).abs().max()) < 2e-2 and float((rgb.detach().cpu() - c_e).abs().max()) < 5e-4
    r_p, r_e = rel(flat.grad.cpu(), gp_e), rel(emb_d.grad.cpu(), ge_e)
    # (2) fp32 oracle
    fc = [p.clone().requires_grad_(True) for p in params]
    emb_o = emb.clone().requires_grad_(True)
    a_o, c_o = O.occmap_forward(fc, emb_o)
    ((a_o.reshape(-1) * ga).sum() + (c_o.reshape(-1, 3) * gc).sum()).backward()
    gp_o = torch.cat([p.grad.reshape(-1) for p in fc])
    err_a = float((alpha.detach().cpu() - a_o.detach().reshape(-1)).abs().max())
    err_c = float((rgb.detach().cpu() - c_o.detach().reshape(-1, 3)).abs().max())
    r_po, r_eo = rel(flat.grad.cpu(), gp_o), rel(emb_d.grad.cpu(), emb_o.grad)
    print("P=%d: vs fp16 emulation grads %.2e g_emb %.2e | vs fp32 oracle alpha %.2e rgb %.2e grads %.3f g_emb %.3f"
          % (P, r_p, r_e, err_a, err_c, r_po, r_eo))
    assert torch.isfinite(flat.grad).all() and torch.isfinite(emb_d.grad).all()
    # bounds = 2x measured (B200, fp16 operands): vs emulation 3.0e-3; vs fp32 oracle alpha 9.2e-3 (|alpha| up to ~15), rgb 2.3e-4,
    # gradients 2.2 %
    assert r_p < 6e-3 and r_e < 6e-3
    assert err_a < 2e-2 and err_c < 5e-4
    assert r_po < 0.045 and r_eo < 0.045


@pytest.mark.parametrize("P", [77, 128, 1200 * 14, 2 * 148 * 128 + 9])
def test_occmap_tc_points_variant(P):
    """Embedding computed inside the kernels (cnf_occmap_tc_fwd_points / _bwd_points) against the embedding-in kernels fed with the
    fp32 embedding of the same points: outputs 2e-2 / 2e-3 (fast sincos + double-angle recurrence vs libm, before the same bf16
    rounding), parameter gradient 2e-2 relative L2, and the B_layer gradient against cnf_embed_bwd of the other path's g_emb
    (5e-2; also against the fp32 oracle: 0.2)."""
    from category_nerf_reconstruction_official_b200 import ops
    g = torch.Generator().manual_seed(900 + P % 89)
    params = O.init_params(O.occmap_param_shapes(128), g)
    B = O.icosa_matrix() + 0.01 * torch.randn(21, 3, generator=g)
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 6.0
    ga = torch.randn(P, generator=g) * 0.1; gc = torch.randn(P, 3, generator=g)
    flat = torch.cat([p.reshape(-1) for p in params]).to(DEV)
    scale = torch.tensor(5.0, device=DEV)
    alpha, rgb, gp, gB = ops.occmap_tc_points_fwd_bwd(pts.to(DEV), flat, B.to(DEV), scale, lambda a, c: (ga.to(DEV), gc.to(DEV)))
    # the embedding-in path on the same points
    Bd = B.to(DEV).requires_grad_(True)
    emb = ops.EmbedFn.apply(pts.to(DEV)[None], Bd[None], scale.reshape(1))[0]
    flat_r = flat.clone().requires_grad_(True)
    a2, c2 = ops.OccMapTCFn.apply(emb, flat_r)
    ((a2 * ga.to(DEV)).sum() + (c2 * gc.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    # fp32 oracle for the B gradient
    Bo = B.clone().requires_grad_(True)
    a_o, c_o = O.occmap_forward([p.clone() for p in params], O.embed(pts[:, None, :], Bo, 5.0).reshape(P, 129))
    ((a_o.reshape(-1) * ga).sum() + (c_o.reshape(-1, 3) * gc).sum()).backward()
    e_a, e_c = float((alpha - a2).abs().max()), float((rgb - c2).abs().max())
    r_p, r_B, r_Bo = rel(gp, flat_r.grad), rel(gB, Bd.grad), rel(gB.cpu(), Bo.grad)
    print("P=%d points variant vs embedding-in: alpha %.2e rgb %.2e grads %.2e B %.2e | B vs fp32 oracle %.3f" % (P, e_a, e_c, r_p, r_B, r_Bo))
    assert torch.isfinite(gp).all() and torch.isfinite(gB).all()
    assert e_a < 1.4e-2 and e_c < 3e-4          # measured 6.8e-3 / 1.4e-4
    assert r_p < 1e-2 and r_B < 1.4e-2 and r_Bo < 0.05          # measured 5.0e-3 / 7.0e-3 / 0.023


def test_occmap_tc_eval_points_vs_oracle():
    """cnf_occmap_tc_eval_points (Trainer.eval_points of the background model) against the fp32 oracle: occupancy = sigmoid(alpha)
    and colour, 2e-2 absolute, ragged point count over several tiles per CTA."""
    from category_nerf_reconstruction_official_b200 import ops
    g = torch.Generator().manual_seed(77)
    P = 148 * 128 * 2 + 55
    params = O.init_params(O.occmap_param_shapes(128), g)
    B = O.icosa_matrix() + 0.01 * torch.randn(21, 3, generator=g)
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 8.0
    flat = torch.cat([p.reshape(-1) for p in params]).to(DEV)
    occ, rgb = ops.occmap_eval_points_fused(pts.to(DEV), flat, B.to(DEV), torch.tensor(5.0, device=DEV))
    torch.cuda.synchronize()
    a_o, c_o = O.occmap_forward(params, O.embed(pts[:, None, :], B, 5.0).reshape(P, 129))
    occ_o = torch.sigmoid(a_o.reshape(-1))
    e_o, e_c = float((occ.cpu() - occ_o).abs().max()), float((rgb.cpu() - c_o.reshape(-1, 3)).abs().max())
    print("occmap eval_points: occupancy max abs err %.2e (range %.3f..%.3f) colour %.2e" % (e_o, float(occ_o.min()), float(occ_o.max()), e_c))
    assert float(occ_o.max() - occ_o.min()) > 0.5
    assert e_o < 4e-3 and e_c < 5e-4          # measured 1.6e-3 / 2.0e-4


@pytest.mark.parametrize("R,S", [(700, 14), (130, 96)])
def test_occmap_tc_render_rays_vs_oracle(R, S):
    """cnf_occmap_tc_render_rays (dense views of the background model) against the oracle chain embed -> OccupancyMap -> sigmoid ->
    termination -> render: every output max-abs 5e-2 (bf16 operands, x10 occupancy logit), weights mean-abs 5e-3."""
    from category_nerf_reconstruction_official_b200 import ops
    g = torch.Generator().manual_seed(50 + S)
    params = O.init_params(O.occmap_param_shapes(128), g)
    B = O.icosa_matrix() + 0.01 * torch.randn(21, 3, generator=g)
    org = (torch.rand(R, 3, generator=g) - 0.5) * 4.0
    dirs = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    z = torch.sort(torch.rand(R, S, generator=g) * 5.0, dim=-1).values
    flat = torch.cat([p.reshape(-1) for p in params]).to(DEV)
    out = ops.occmap_render_rays_fused(org.to(DEV), dirs.to(DEV), z.to(DEV), flat, B.to(DEV), torch.tensor(5.0, device=DEV), want_weights=True)
    torch.cuda.synchronize()
    pcs = org[:, None, :] + dirs[:, None, :] * z[..., None]
    a_o, c_o = O.occmap_forward(params, O.embed(pcs, B, 5.0))
    occ = torch.sigmoid(a_o.reshape(R, S)); col = c_o.reshape(R, S, 3)
    w = O.termination(occ)
    D = (w * z).sum(-1); Cc = (w[..., None] * col).sum(-2); Op = w.sum(-1)
    errs = [float((out["weights"].cpu() - w).abs().max()), float((out["depth"].cpu() - D).abs().max()),
            float((out["rgb"].cpu() - Cc).abs().max()), float((out["opacity"].cpu() - Op).abs().max())]
    print("occmap render_rays R=%d S=%d: max abs err weights %.2e depth %.2e colour %.2e opacity %.2e" % (R, S, *errs))
    assert max(errs) < 3e-3 and float((out["weights"].cpu() - w).abs().mean()) < 1.2e-3          # measured max 1.24e-3
    torch.testing.assert_close((out["weights"] * z.to(DEV)).sum(-1), out["depth"], rtol=1e-4, atol=1e-5)


def test_occmap_module_tensor_core_background_step():
    """OccupancyMap module with use_tensor_cores: the reference's background step (train.py:172-180: embed -> OccupancyMap ->
    step_batch_loss -> backward) against the fp32 oracle step on a well-conditioned loss (all depths invalid -> depth term off)."""
    from category_nerf_reconstruction_official_b200 import model, embedding, loss
    g = torch.Generator().manual_seed(5)
    R, S = 1200, 14
    params = O.init_params(O.occmap_param_shapes(128), g)
    m = model.OccupancyMap(87, 42, hidden_size=128).to(DEV)
    with torch.no_grad():
        for prm, src in zip(m.parameters(), params):
            prm.copy_(src)
    m.use_tensor_cores = True
    pe = embedding.UniDirsEmbed(max_deg=5, scale=5.0).to(DEV)
    B = pe.B_layer.weight.detach().cpu().clone()
    org = (torch.rand(R, 3, generator=g) - 0.5) * 2; dirs = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    z = torch.sort(torch.rand(R, S, generator=g) * 4.0, dim=-1).values
    pcs = org[:, None, :] + dirs[:, None, :] * z[..., None]
    gt_d = torch.zeros(R); gt_rgb = torch.randint(0, 256, (R, 3), generator=g, dtype=torch.uint8)
    labels = torch.randint(0, 3, (R,), generator=g, dtype=torch.uint8); md = torch.zeros(R, dtype=torch.bool)
    emb = pe(pcs.to(DEV))
    a, c = m(emb)
    l, ld, _ = loss.step_batch_loss(a[None], c[None], gt_d.to(DEV)[None], (gt_rgb.to(DEV) / 255.)[None], labels.to(DEV)[None],
                                    md.to(DEV)[None], z.to(DEV)[None])
    l.backward()
    torch.cuda.synchronize()
    out = O.bg_train_step(params, B, 5.0, pcs, z, gt_d, gt_rgb, labels, md)
    torch.testing.assert_close(l.detach().cpu(), out["loss"], rtol=1e-4, atol=1e-5)
    gp = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu()
    gp_o = torch.cat([t.reshape(-1) for t in out["g_fc"]])
    r_p, r_B = rel(gp, gp_o), rel(pe.B_layer.weight.grad.cpu(), out["g_B"])
    print("background step: loss %.5f vs %.5f, whole MLP gradient rel L2 %.3f, B_layer gradient rel L2 %.3f" % (float(l), float(out["loss"]), r_p, r_B))
    assert r_p < 0.02 and r_B < 0.03          # measured 0.010 / 0.015


@pytest.mark.parametrize("well_conditioned", [True, False])
def test_fused_background_trainer_step_vs_oracle(well_conditioned):
    """FusedBackgroundTrainer.step (sampler -> embed -> tensor-core OccupancyMap -> composite/loss -> backward, C ABI only) against
    the oracle's background step on the reference's 1200 x (5+9) batch.  well_conditioned: all depths invalid, so the depth term
    (1/(sqrt(var)+1e-4) weights) is off by the reference's empty-mask rule: whole-gradient rel L2 <= 0.1; otherwise the loose
    mapping bounds of the CodeNeRF sweep (rel L2 <= 0.35, cosine >= 0.93).  Also: CUDA-graph replay reproduces the eager step."""
    from category_nerf_reconstruction_official_b200.background import FusedBackgroundTrainer
    g = torch.Generator().manual_seed(31)
    R, N, M = 1200, 5, 9
    tr = FusedBackgroundTrainer(scale=5.0, device=DEV, n_bins_cam2surface=N, n_bins=M, generator=g)
    rgbs, depth, org, dirs, _ = O.synth_rays(R, 1, g)
    if well_conditioned:
        depth = torch.zeros(R)
    rand = {"u_inv": torch.rand(R, N + M, generator=g), "u_val": torch.rand(R, N, generator=g),
            "n_obj": torch.randn(R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(R, M, generator=g)}
    dv = lambda t: t.to(DEV)
    loss = tr.step(dv(rgbs), dv(depth), dv(org), dv(dirs), rand={k: dv(v) for k, v in rand.items()})
    torch.cuda.synchronize()
    _, _, valid, labels, pcs, z = O.sample_points(rgbs, depth, org, dirs, N, M, 0.1, 0.05, rand)
    fc = [p.cpu() for p in tr.param_views()]
    out = O.bg_train_step(fc, tr.pe_B.cpu(), 5.0, pcs, z, depth, rgbs[:, :3], labels, valid)
    gp_o = torch.cat([t.reshape(-1) for t in out["g_fc"]])
    r_p, r_B = rel(tr.g_params.cpu(), gp_o), rel(tr.g_B.cpu(), out["g_B"])
    cos = float(torch.nn.functional.cosine_similarity(tr.g_params.cpu().double(), gp_o.double(), dim=0))
    print("background trainer (well_conditioned=%s): loss %.5f vs %.5f, gradient rel L2 %.3f cos %.4f, B rel L2 %.3f"
          % (well_conditioned, float(loss), float(out["loss"]), r_p, cos, r_B))
    torch.testing.assert_close(loss.cpu(), out["loss"], rtol=2e-4, atol=1e-5)
    if well_conditioned:
        assert r_p < 0.005 and r_B < 0.012          # measured 0.002 / 0.006
    else:
        assert r_p < 0.004 and cos > 0.9999 and r_B < 0.014          # measured 0.001 / 1.0000 / 0.007
    # graph replay == eager
    g_eager = tr.g_params.clone(); l_eager = tr.out_loss.clone()
    rgbs_d, depth_d, org_d, dirs_d = dv(rgbs), dv(depth), dv(org), dv(dirs)
    replay = tr.capture(rgbs_d, depth_d, org_d, dirs_d)
    # the captured step reads its random offsets from the trainer's prefetch buffers: put the eager step's offsets there
    tr._rand_u.copy_(torch.cat([rand["u_inv"].reshape(-1), rand["u_val"].reshape(-1), rand["u_oth"].reshape(-1)]).to(DEV))
    tr._rand_n.copy_(rand["n_obj"].reshape(1, R, M).to(DEV))
    tr.g_params.zero_(); replay(); torch.cuda.synchronize()
    torch.testing.assert_close(tr.out_loss, l_eager, rtol=1e-5, atol=1e-6)
    assert rel(tr.g_params, g_eager) < 1e-4


def test_capture_iteration_objects_plus_background():
    """One whole iteration (object categories + background model as parallel branches of one CUDA graph): replays run, both
    trainers' losses and gradients are finite and change from replay to replay (fresh random offsets every step), and the graph's
    first replay matches an eager step fed with the same random offsets."""
    from category_nerf_reconstruction_official_b200.background import FusedBackgroundTrainer, capture_iteration
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    g = torch.Generator().manual_seed(3)
    K, n_obj, R = 4, 3, 240
    obj = FusedCategoryTrainer(K, n_obj, latent_dim=32, scale=2.0, device=DEV, n_bins_cam2surface=1, n_bins=9, generator=g)
    bg = FusedBackgroundTrainer(scale=5.0, device=DEV, generator=g)
    cats = [O.synth_rays(R, n_obj, g) for _ in range(K)]
    dv = lambda t: t.to(DEV)
    oi = (dv(torch.stack([c[0] for c in cats])), dv(torch.stack([c[1] for c in cats])), dv(torch.stack([c[2] for c in cats])),
          dv(torch.stack([c[3] for c in cats])), dv(torch.stack([c[4] for c in cats]).to(torch.int32)))
    bi = tuple(dv(t) for t in O.synth_rays(600, 1, g)[:4])
    replay = capture_iteration(obj, oi, bg, bi)
    # eager reference for the NEXT replay: same offsets (the prefetch buffers hold what the next step will consume)
    ru, rn, bu, bn = obj._rand_u.clone(), obj._rand_n.clone(), bg._rand_u.clone(), bg._rand_n.clone()
    replay(); torch.cuda.synchronize()
    lo1, lb1 = obj.out_loss.clone(), bg.out_loss.clone()
    go1, gb1 = obj.g_params.clone(), bg.g_params.clone()
    assert torch.isfinite(lo1).all() and torch.isfinite(lb1).all() and torch.isfinite(go1).all() and torch.isfinite(gb1).all()
    assert float(go1.abs().sum()) > 0 and float(gb1.abs().sum()) > 0
    replay(); torch.cuda.synchronize()
    assert float((obj.out_loss - lo1).abs().max()) > 0 and float((bg.out_loss - lb1).abs().max()) > 0      # new offsets every step
    obj._rand_u.copy_(ru); obj._rand_n.copy_(rn); bg._rand_u.copy_(bu); bg._rand_n.copy_(bn)
    obj.step(*oi); bg.step(*bi); torch.cuda.synchronize()
    torch.testing.assert_close(obj.out_loss, lo1, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(bg.out_loss, lb1, rtol=1e-5, atol=1e-6)
    assert rel(obj.g_params, go1) < 1e-4 and rel(bg.g_params, gb1) < 1e-4

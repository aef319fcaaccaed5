"""GPU parity (through the C ABI) of the standalone kernels: sampler (bit-exact), embedding, fp32 MLPs,
compositing + losses -- against the committed golden vectors (reference outputs) and the CPU oracle."""
import numpy as np
import pytest
import torch

from helpers import SAMPLER_CASES, STEP_CASES, assert_grad_close, golden, gt, step_case
from oracle import cnf_oracle as O

from category_nerf_reconstruction_official_b200 import dev as cnf_dev      # hardware self-tests: development build (cnf_dev.h)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pkg():
    from category_nerf_reconstruction_official_b200 import embedding, ensemble, loss, model, ops, render_rays, sampling
    return embedding, ensemble, loss, model, ops, render_rays, sampling


@pytest.mark.parametrize("name", SAMPLER_CASES)
def test_sampler_bit_exact_vs_golden(name):
    *_, sampling = _pkg()
    p = "sampler/%s/" % name
    N, M = [int(v) for v in golden("sampler")[p + "NM"]]
    rand = {k: gt("sampler", p + k).to(DEV) for k in ("u_inv", "u_val", "n_obj", "u_oth")}
    rgb, d, valid, labels, pcs, z = sampling.sample_3d_points(
        gt("sampler", p + "rgbs").to(DEV), gt("sampler", p + "depth").to(DEV), gt("sampler", p + "origins").to(DEV),
        gt("sampler", p + "dirs").to(DEV), N, M, 0.1, 0.05, rand=rand)
    assert torch.equal(z.cpu(), gt("sampler", p + "z"))
    assert torch.equal(pcs.cpu(), gt("sampler", p + "pcs"))
    assert torch.equal(valid.cpu(), gt("sampler", p + "valid"))
    assert torch.equal(labels.cpu(), gt("sampler", p + "labels"))


def test_sampler_large_bit_exact_vs_oracle():
    """K=3 categories x 5000 rays (multi-block scan, ragged last block), both S=10 and S=14."""
    *_, ops, _, sampling = _pkg()
    for N, M in ((1, 9), (5, 9)):
        g = torch.Generator().manual_seed(5)
        K, R = 3, 5000
        cats = [O.synth_rays(R, 4, g) for _ in range(K)]
        rgbs = torch.stack([c[0] for c in cats]); depth = torch.stack([c[1] for c in cats])
        org = torch.stack([c[2] for c in cats]); dirs = torch.stack([c[3] for c in cats])
        rand = {"u_inv": torch.rand(K, R, N + M, generator=g), "u_val": torch.rand(K, R, N, generator=g),
                "n_obj": torch.randn(K, R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(K, R, M, generator=g)}
        z, pcs, valid, labels, stats = ops.sample_batched(rgbs.to(DEV), depth.to(DEV), org.to(DEV), dirs.to(DEV), N, M,
                                                          0.1, 0.05, {k: v.to(DEV) for k, v in rand.items()})
        for k in range(K):
            _, _, v_o, l_o, pcs_o, z_o = O.sample_points(rgbs[k], depth[k], org[k], dirs[k], N, M, 0.1, 0.05,
                                                         {kk: vv[k] for kk, vv in rand.items()})
            assert torch.equal(z[k].cpu(), z_o) and torch.equal(pcs[k].cpu(), pcs_o)
            assert torch.equal(valid[k].cpu(), v_o)
            lab = rgbs[k, :, 3]
            st = stats[k].cpu()
            assert int(st[0]) == int((~v_o).sum()) and int(st[1]) == int(v_o.sum())
            assert int(st[4]) == int((v_o & (lab != 0)).sum()) and int(st[5]) == int((lab != 0).sum())
            assert int(st[6]) == int((lab != 2).sum())
        # sortedness property of the surface-band samples of this-object rays
        lab = rgbs[..., 3].to(DEV)
        m = valid & (lab == 1)
        zs = z[m][:, N:]
        assert bool((zs[:, 1:] >= zs[:, :-1]).all())


def test_stratified_bins_and_normal_bins_module_level():
    *_, sampling = _pkg()
    u = gt("sampler", "sampler/strat96/u").to(DEV)
    z = sampling.stratified_bins(0.0, torch.tensor(2.5), 96, 16, device=DEV, u=u)
    assert torch.equal(z.cpu(), gt("sampler", "sampler/strat96/z"))
    g = torch.Generator().manual_seed(2)
    depth = torch.rand(50, generator=g) * 3 + 0.2
    n = torch.randn(50, 9, generator=g) * (0.1 / 3)
    zz = sampling.normal_bins_sampling(depth.to(DEV), 9, 50, 0.1, device=DEV, n=n.to(DEV))
    assert torch.equal(zz.cpu(), O.normal_bins(depth, n, 0.1))
    # default (device RNG) path: shapes + range
    z2 = sampling.stratified_bins(0.5, 1.5, 7, 33, device=DEV)
    assert z2.shape == (33, 7) and float(z2.min()) >= 0.5 and float(z2.max()) <= 1.5


@pytest.mark.parametrize("scale", [2.0, 3.0, 5.0])
def test_embed_vs_golden(scale):
    embedding, *_ = _pkg()
    p = "embed/s%g/" % scale
    pe = embedding.UniDirsEmbed(max_deg=5, scale=scale).to(DEV)
    pe.B_layer.weight.data = gt("embed", p + "B").to(DEV)
    x = gt("embed", p + "x").to(DEV).requires_grad_(True)
    emb = pe(x)
    # fp32 tolerance: arguments reach ~100 rad, one ulp there is 7.6e-6 (SURVEY 7 'Precision/parity')
    torch.testing.assert_close(emb.cpu(), gt("embed", p + "emb"), rtol=0, atol=3e-5)
    (emb * gt("embed", p + "gout").to(DEV)).sum().backward()
    torch.testing.assert_close(pe.B_layer.weight.grad.cpu(), gt("embed", p + "gB"), rtol=2e-4, atol=2e-3)
    torch.testing.assert_close(x.grad.cpu(), gt("embed", p + "gx"), rtol=2e-4, atol=2e-3)


@pytest.mark.parametrize("name", ["normal", "empty_depth_cat1"])
def test_step_batch_loss_vs_golden(name):
    _, _, loss, *_ = _pkg()
    p = "loss/%s/" % name
    a = gt("loss", p + "alpha").to(DEV).requires_grad_(True)
    c = gt("loss", p + "color").to(DEV).requires_grad_(True)
    l, ld, lcol = loss.step_batch_loss(a, c, gt("loss", p + "gt_depth").to(DEV), gt("loss", p + "gt_color").to(DEV),
                                       gt("loss", p + "labels").to(DEV), gt("loss", p + "mask_depth").to(DEV),
                                       gt("loss", p + "z").to(DEV))
    l.backward()
    torch.testing.assert_close(l.cpu(), gt("loss", p + "loss"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ld["depth"].cpu(), gt("loss", p + "depth"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ld["color"].cpu(), gt("loss", p + "colorl"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ld["opacity"].cpu(), gt("loss", p + "opacity"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a.grad.cpu(), gt("loss", p + "galpha"), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(c.grad.cpu(), gt("loss", p + "gcolor"), rtol=1e-4, atol=1e-7)


def test_termination_and_long_rays():
    _, _, loss, _, _, render_rays, _ = _pkg()
    t = render_rays.occupancy_to_termination(gt("loss", "loss/term2d/occ").to(DEV))
    torch.testing.assert_close(t.cpu(), gt("loss", "loss/term2d/term"), rtol=1e-5, atol=1e-7)
    # S=96 (multi-chunk scan) forward + backward against the oracle
    g = torch.Generator().manual_seed(9)
    K, R, S = 2, 37, 96
    alpha = torch.randn(K, R, S, 1, generator=g) * 2 - 2
    color = torch.rand(K, R, S, 3, generator=g)
    z = torch.sort(torch.rand(K, R, S, generator=g) * 3, dim=-1).values
    gd = torch.rand(K, R, generator=g) * 3; gc = torch.rand(K, R, 3, generator=g)
    lab = torch.randint(0, 3, (K, R), generator=g).to(torch.uint8); md = torch.rand(K, R, generator=g) > 0.2
    a0 = alpha.clone().requires_grad_(True); c0 = color.clone().requires_grad_(True)
    lo, ldo, _ = O.step_loss(a0, c0, gd, gc, lab, md, z); lo.backward()
    a1 = alpha.to(DEV).requires_grad_(True); c1 = color.to(DEV).requires_grad_(True)
    l1, ld1, _ = loss.step_batch_loss(a1, c1, gd.to(DEV), gc.to(DEV), lab.to(DEV), md.to(DEV), z.to(DEV)); l1.backward()
    torch.testing.assert_close(l1.cpu(), lo.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a1.grad.cpu(), a0.grad, rtol=2e-4, atol=1e-7)
    torch.testing.assert_close(c1.grad.cpu(), c0.grad, rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("name", STEP_CASES)
def test_module_api_train_step_vs_golden(name):
    """train.py:136-182 written against OUR modules (update_vmap / vmap / step_batch_loss / step_batch_loss_reg),
    fp32 CUDA-core path; compared with the reference's outputs and gradients."""
    embedding, ensemble, loss, model, ops, _, _ = _pkg()
    c = step_case(name)
    K, n_obj, L = c["K"], c["n_obj"], c["L"]
    fcs, pes, trs = [], [], []
    names = [n for n, _ in model.CodeNeRF(87, 42, W=32, latent_dim=L).named_parameters()]
    for k in range(K):
        m = model.CodeNeRF(87, 42, W=32, latent_dim=L).to(DEV)
        with torch.no_grad():
            for i, (n, p) in enumerate(m.named_parameters()):
                p.copy_(c["fc"][i][k])
        pe = embedding.UniDirsEmbed(max_deg=5, scale=c["scale"]).to(DEV)
        pe.B_layer.weight.data = c["B"][k].to(DEV)
        fcs.append(m); pes.append(pe)

        class T: pass
        class C: pass
        t = T(); t.n_obj = n_obj
        t.shape_codes = torch.nn.Embedding(n_obj, L).to(DEV); t.texture_codes = torch.nn.Embedding(n_obj, L).to(DEV)
        t.shape_codes.weight.data = c["sc"][k].to(DEV); t.texture_codes.weight.data = c["tc"][k].to(DEV)
        cc = C(); cc.trainer = t; cc.obj_ids = list(range(n_obj)); cc.training_device = DEV
        trs.append(cc)
    opt = torch.optim.AdamW([torch.zeros(1, requires_grad=True)], lr=1e-3)
    fc_model, fc_param, fc_buffer = ensemble.update_vmap(fcs, opt)
    pe_model, pe_param, pe_buffer = ensemble.update_vmap(pes, opt)
    idx = c["indices"].to(DEV)
    sh = torch.stack([trs[k].trainer.shape_codes(idx[k])[:, None, :] for k in range(K)])
    tx = torch.stack([trs[k].trainer.texture_codes(idx[k])[:, None, :] for k in range(K)])
    emb = ensemble.vmap(pe_model)(pe_param, pe_buffer, c["pcs"].to(DEV))
    alpha, color = ensemble.vmap(fc_model)(fc_param, fc_buffer, emb, sh, tx)
    l, ld, lcol = loss.step_batch_loss(alpha, color, c["gt_depth"].to(DEV), (c["gt_rgb"].to(DEV) / 255.),
                                       c["labels"].to(DEV), c["mask_depth"].to(DEV), c["z"].to(DEV))
    rs, rt = loss.step_batch_loss_reg({k: trs[k] for k in range(K)}, torch.arange(K))
    l = l + 0.0005 * (rs + rt).sum()
    l.backward()
    torch.testing.assert_close(alpha.detach().cpu(), c["alpha"], rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(color.detach().cpu(), c["color"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(l.detach().cpu(), c["loss"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rs.detach().cpu(), c["reg_shape"], rtol=1e-5, atol=1e-6)
    for i in range(28):
        for k in range(K):      # per category: gradient scales differ by orders of magnitude between categories
            assert_grad_close(fc_param[i].grad[k], c["gfc"][i][k], rel_l2=5e-3, max_bad_frac=0.1, name="%s[cat %d]" % (names[i], k))
    for k in range(K):
        assert_grad_close(pe_param[0].grad[k], c["gB"][k], elem_atol=5e-4, name="B_layer")
    for k in range(K):
        assert_grad_close(trs[k].trainer.shape_codes.weight.grad, c["gsc"][k], elem_atol=1e-5, name="shape_codes")
        assert_grad_close(trs[k].trainer.texture_codes.weight.grad, c["gtc"][k], elem_atol=1e-5, name="texture_codes")


def test_occmap_bg_step_vs_golden():
    embedding, _, loss, model, *_ = _pkg()
    p = "bg/H64/"
    m = model.OccupancyMap(87, 42, hidden_size=64).to(DEV)
    with torch.no_grad():
        for i, (n, prm) in enumerate(m.named_parameters()):
            prm.copy_(gt("bg", p + "fc%02d" % i))
    pe = embedding.UniDirsEmbed(max_deg=5, scale=5.0).to(DEV)
    pe.B_layer.weight.data = gt("bg", p + "B").to(DEV)
    emb = pe(gt("bg", p + "pcs").to(DEV))
    a, c = m(emb)
    l, ld, _ = loss.step_batch_loss(a[None], c[None], gt("bg", p + "gt_depth").to(DEV)[None],
                                    (gt("bg", p + "gt_rgb").to(DEV) / 255.)[None], gt("bg", p + "labels").to(DEV)[None],
                                    gt("bg", p + "mask_depth").to(DEV)[None], gt("bg", p + "z").to(DEV)[None])
    l.backward()
    torch.testing.assert_close(a.detach().cpu(), gt("bg", p + "alpha"), rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(c.detach().cpu(), gt("bg", p + "color"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(l.detach().cpu(), gt("bg", p + "loss"), rtol=1e-4, atol=1e-5)
    for i, (n, prm) in enumerate(m.named_parameters()):
        assert_grad_close(prm.grad, gt("bg", p + "gfc%02d" % i), name=n)
    assert_grad_close(pe.B_layer.weight.grad, gt("bg", p + "gB"), elem_atol=5e-4, name="B_layer")


def test_eval_points_single_code():
    """Trainer.eval_points semantics: one code broadcast over all points == oracle forward."""
    embedding, _, _, model, *_ = _pkg()
    c = step_case("K2_L32")
    m = model.CodeNeRF(87, 42, W=32, latent_dim=32).to(DEV)
    with torch.no_grad():
        for i, (n, p) in enumerate(m.named_parameters()):
            p.copy_(c["fc"][i][0])
    pe = embedding.UniDirsEmbed(max_deg=5, scale=c["scale"]).to(DEV)
    pts = c["pcs"][0].reshape(-1, 3)
    with torch.no_grad():
        a, col = m(pe(pts.to(DEV)), c["sc"][0][1].to(DEV), c["tc"][0][1].to(DEV))
    emb_o = O.embed(pts, O.icosa_matrix(), c["scale"])
    a_o, c_o = O.codenerf_forward([t[0] for t in c["fc"]], emb_o, c["sc"][0][1], c["tc"][0][1])
    torch.testing.assert_close(a.cpu(), a_o, rtol=1e-4, atol=3e-4)
    torch.testing.assert_close(col.cpu(), c_o, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("N,Kd", [(32, 96), (48, 80), (128, 32), (16, 32), (32, 128)])
def test_umma_descriptor_selftest(mode, N, Kd):
    """One tcgen05 tile per operand view; pins the smem/instruction descriptor encodings on hardware."""
    *_, ops, _, _ = _pkg()
    g = torch.Generator().manual_seed(mode * 100 + N + Kd)
    bf = lambda t: t.bfloat16().float()
    if mode == 0:
        A = bf(torch.randn(128, Kd, generator=g)); B = bf(torch.randn(N, Kd, generator=g)); ref = A @ B.t()
    elif mode == 1:
        A = bf(torch.randn(128, Kd, generator=g)); B = bf(torch.randn(Kd, N, generator=g)); ref = A @ B
    else:
        A = bf(torch.randn(Kd, 128, generator=g)); B = bf(torch.randn(Kd, N, generator=g)); ref = A.t() @ B
    D = cnf_dev.umma_selftest(mode, A.to(DEV), B.to(DEV), N, Kd)
    torch.testing.assert_close(D.cpu(), ref, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("N,Kd", [(32, 96), (48, 80), (32, 32), (96, 32), (16, 32), (32, 128)])
def test_umma_tmem_operand_selftest(N, Kd):
    """A operand staged in tensor memory (packed bf16 pairs per 32-bit column), as the fused forward / data-gradient GEMMs use it."""
    *_, ops, _, _ = _pkg()
    g = torch.Generator().manual_seed(300 + N + Kd)
    bf = lambda t: t.bfloat16().float()
    A = bf(torch.randn(128, Kd, generator=g)); B = bf(torch.randn(N, Kd, generator=g))
    D = cnf_dev.umma_selftest(3, A.to(DEV), B.to(DEV), N, Kd)
    torch.testing.assert_close(D.cpu(), A @ B.t(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("mode", [4, 5])
@pytest.mark.parametrize("N", [32, 48, 16])
def test_umma_m64_wgrad_lane_mapping(mode, N):
    """M = 64 weight-gradient tiles: accumulator row r lands in lane 32*(r/16) + r%16 (+16 with the lane offset); other lanes untouched."""
    *_, ops, _, _ = _pkg()
    g = torch.Generator().manual_seed(400 + mode + N)
    bf = lambda t: t.bfloat16().float()
    Kd = 128
    A = bf(torch.randn(Kd, 64, generator=g)); B = bf(torch.randn(Kd, N, generator=g))
    ref = A.t() @ B                                   # [64, N]
    D = cnf_dev.umma_selftest(mode, A.to(DEV), B.to(DEV), N, Kd).cpu()
    off = 16 if mode == 5 else 0
    lanes = torch.tensor([32 * (r // 16) + r % 16 + off for r in range(64)])
    torch.testing.assert_close(D[lanes], ref, rtol=1e-4, atol=2e-3)
    other = torch.ones(128, dtype=torch.bool); other[lanes] = False
    assert torch.all(D[other] == -777.0)


@pytest.mark.parametrize("n", [1, 255, 5000])
def test_origin_dirs_vs_linalg_inv(n):
    """origin_dirs_O / origin_dirs_W (scene_cateogries.py:24-47) against torch.linalg.inv in float64 on sim3-shaped per-ray poses
    (rotation x scale in [0.3, 3], translations up to 5): 2e-5 relative to the largest entry."""
    from category_nerf_reconstruction_official_b200 import sampling
    g = torch.Generator().manual_seed(n)
    q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=g, dtype=torch.float64))
    s = torch.rand(n, 1, 1, generator=g, dtype=torch.float64) * 2.7 + 0.3
    T = torch.zeros(n, 4, 4, dtype=torch.float64)
    T[:, :3, :3] = q * s; T[:, :3, 3] = (torch.rand(n, 3, generator=g, dtype=torch.float64) - 0.5) * 10; T[:, 3, 3] = 1
    d = torch.randn(n, 3, generator=g, dtype=torch.float64)
    Ti = torch.linalg.inv(T)
    o_ref, d_ref = Ti[:, :3, 3], (Ti[:, :3, :3] @ d[..., None]).squeeze(-1)
    o, dd = sampling.origin_dirs_O(T.float().to(DEV), d.float().to(DEV))
    assert float((o.cpu().double() - o_ref).abs().max()) < 2e-5 * float(o_ref.abs().max())
    assert float((dd.cpu().double() - d_ref).abs().max()) < 2e-5 * float(d_ref.abs().max())
    o, dd = sampling.origin_dirs_W(T.float().to(DEV), d.float().to(DEV))
    torch.testing.assert_close(o.cpu().double(), T[:, :3, 3], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(dd.cpu().double(), (T[:, :3, :3] @ d[..., None]).squeeze(-1), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("K,R,S", [(1, 1200, 14), (3, 333, 10), (2, 129, 32)])
def test_composite_fwd_bwd_one_pass_matches_two_calls(K, R, S):
    """cnf_composite_loss_fwd_bwd (one pass) against cnf_composite_loss_fwd + cnf_composite_loss_bwd (themselves checked against the
    golden reference outputs above): identical arithmetic, so loss terms, rays and gradients must agree to 1e-6."""
    from category_nerf_reconstruction_official_b200 import ops, _lib
    from category_nerf_reconstruction_official_b200._lib import check, ptr, stream
    g = torch.Generator().manual_seed(K * 100 + S)
    alpha = (torch.randn(K, R, S, generator=g) * 3).to(DEV); color = torch.rand(K, R, S, 3, generator=g).to(DEV)
    z = torch.sort(torch.rand(K, R, S, generator=g) * 4, dim=-1).values.to(DEV)
    gd = (torch.rand(K, R, generator=g) * 4).to(DEV); gc = torch.rand(K, R, 3, generator=g).to(DEV)
    lab = torch.randint(0, 3, (K, R), generator=g, dtype=torch.uint8).to(DEV)
    md = (torch.rand(K, R, generator=g) > 0.2).to(torch.uint8).to(DEV)
    lib = _lib.load()
    counts = ops.mask_counts(lab, md)
    ws = torch.empty(lib.cnf_composite_workspace_bytes(K, R), dtype=torch.uint8, device=DEV)
    one = torch.ones(1, device=DEV)
    f = lambda *s: torch.empty(*s, device=DEV)
    rays1, out1, ga1, gc1 = f(K, R, 6), f(4 + 3 * K), f(K, R, S), f(K, R, S, 3)
    rays2, out2, ga2, gc2 = f(K, R, 6), f(4 + 3 * K), f(K, R, S), f(K, R, S, 3)
    args = (K, R, S, ptr(alpha), ptr(color), ptr(z), ptr(gd), ptr(gc), ptr(lab), ptr(md), ptr(counts), 5.0, 10.0)
    check(lib.cnf_composite_loss_fwd(*args, ptr(rays1), ptr(out1), ptr(ws), ws.numel(), stream()))
    check(lib.cnf_composite_loss_bwd(*args, ptr(one), ptr(ga1), ptr(gc1), stream()))
    check(lib.cnf_composite_loss_fwd_bwd(*args, ptr(rays2), ptr(out2), ptr(ga2), ptr(gc2), ptr(ws), ws.numel(), stream()))
    torch.cuda.synchronize()
    assert float(out1[0]) > 0
    torch.testing.assert_close(out2, out1, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(rays2, rays1, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(ga2, ga1, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(gc2, gc1, rtol=1e-6, atol=1e-9)

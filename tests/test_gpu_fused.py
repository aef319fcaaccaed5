"""GPU parity of the fused tcgen05 training step (cnf_train_step) against the golden vectors (reference outputs)
and the CPU oracle.  fp16 operands (power-of-two scaled gradients) / fp32 accumulation.  Every bound below is <= 2x the value
measured on B200 (profiles/r02_parity.txt), so an arithmetic regression fails the suite."""
import pytest
import torch

from helpers import STEP_CASES, assert_grad_close, step_case
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# Tolerances of the tensor-core path (north_star: "within a stated relative tolerance, e.g. 1e-3 for the tf32/bf16 MLP").
# Loss terms: measured 8e-6 .. 9e-4 relative over all fixtures of this file.
LOSS_RTOL = 2e-3
# Gradients, per TENSOR, relative L2: fp16 rounds every GEMM operand to 2^-12 relative; the L1 losses and ReLUs turn that into a few
# flipped signs / masks, which on these tiny fixtures (24-2400 rays; the depth term weights rays by 1/(sqrt(var)+1e-4)) is up to a
# few percent of a small bias-like tensor: measured median 0.1-0.6 %, worst 4.4 % (a 32-element bias), cosine >= 0.99903.
# At the batch sizes the path is built for the same quantity is 0.08-0.9 % (tests/test_gpu_parity_r2.py).
GRAD_REL_L2 = 0.09
GRAD_COS = 0.998


_SEEN = []          # (name, rel L2, cosine) of every tensor compared since the last _summary()


def _grad_ok(actual, expected, name, rel_tol=GRAD_REL_L2, cos_tol=GRAD_COS):
    a = actual.detach().double().cpu().flatten(); e = expected.detach().double().cpu().flatten()
    if float(e.norm()) < 1e-9:
        assert float(a.norm()) < 1e-5, name
        return
    rel = float((a - e).norm() / e.norm())
    cos = float(torch.nn.functional.cosine_similarity(a, e, dim=0))
    _SEEN.append((name, rel, cos))
    assert rel <= rel_tol and cos >= cos_tol, "%s: rel L2 %.3f cos %.4f" % (name, rel, cos)


def _summary(tag, loss=None, ref_loss=None):
    """Print what was measured (pytest -s; copied into profiles/r02_parity.txt) and reset."""
    if _SEEN:
        w = max(_SEEN, key=lambda t: t[1]); c = min(_SEEN, key=lambda t: t[2])
        rels = sorted(t[1] for t in _SEEN)
        msg = "[fused %s] %d tensors: rel L2 median %.4f worst %.4f (%s), min cosine %.5f (%s)" % (tag, len(_SEEN), rels[len(rels) // 2], w[1], w[0], c[2], c[0])
        if loss is not None:
            msg += " | loss rel err %.2e" % (abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)))
        print("\n" + msg)
    del _SEEN[:]


def _trainer_from_case(c):
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    tr = FusedCategoryTrainer(c["K"], c["n_obj"], latent_dim=c["L"], scale=c["scale"], device=DEV,
                              n_bins_cam2surface=c["N"], n_bins=c["M"])
    for k in range(c["K"]):
        tr.load_category(k, [t[k] for t in c["fc"]], c["B"][k], c["sc"][k], c["tc"][k])
    return tr


def _run_case(c, tr):
    from category_nerf_reconstruction_official_b200 import ops
    rgbs = torch.cat([c["gt_rgb"], c["labels"][..., None]], dim=-1).to(DEV)
    labels = c["labels"].to(DEV); md = c["mask_depth"].to(torch.uint8).to(DEV)
    counts = ops.mask_counts(labels, md)
    loss = tr.step_from_samples(c["origins"].to(DEV), c["dirs"].to(DEV), c["z"].to(DEV), c["gt_depth"].to(DEV), rgbs, md,
                                c["indices"].to(DEV), counts)
    torch.cuda.synchronize()
    return loss


@pytest.mark.parametrize("name", STEP_CASES)
def test_fused_step_vs_golden(name):
    c = step_case(name)
    tr = _trainer_from_case(c)
    loss = _run_case(c, tr)
    lt = tr.loss_terms()
    K = c["K"]
    torch.testing.assert_close(lt["reg_shape"].cpu(), c["reg_shape"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lt["reg_texture"].cpu(), c["reg_texture"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lt["depth"].cpu(), c["depth"], rtol=3e-3, atol=1e-3)
    torch.testing.assert_close(lt["color"].cpu(), c["colorl"], rtol=LOSS_RTOL, atol=2e-4)
    torch.testing.assert_close(lt["opacity"].cpu(), c["opacity"], rtol=LOSS_RTOL, atol=2e-4)
    torch.testing.assert_close(loss.cpu(), c["loss"], rtol=LOSS_RTOL, atol=1e-3)
    names = [n for n, _ in O.codenerf_param_shapes(32, c["L"])]
    for k in range(K):
        gv = tr.grad_views(k)
        for i in range(28):
            _grad_ok(gv[i], c["gfc"][i][k], "%s[cat %d]" % (names[i], k))
        _grad_ok(tr.g_B[k], c["gB"][k], "B_layer")
        n = c["n_obj"]
        _grad_ok(tr.g_shape[k, :n], c["gsc"][k], "shape_codes")
        _grad_ok(tr.g_tex[k, :n], c["gtc"][k], "texture_codes")
    _summary("golden " + name, loss, c["loss"])


def test_fused_step_multi_tile_vs_oracle():
    """Several CTAs and tiles per category, ragged last tile, sampler in the loop (injected randoms)."""
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    g = torch.Generator().manual_seed(3)
    K, n_obj, R, N, M, L = 2, 5, 333, 1, 9, 32
    tr = FusedCategoryTrainer(K, n_obj, latent_dim=L, scale=2.0, device=DEV, n_bins_cam2surface=N, n_bins=M, generator=g)
    cats = [O.synth_rays(R, n_obj, g) for _ in range(K)]
    rgbs = torch.stack([c[0] for c in cats]); depth = torch.stack([c[1] for c in cats])
    org = torch.stack([c[2] for c in cats]); dirs = torch.stack([c[3] for c in cats]); idx = torch.stack([c[4] for c in cats])
    rand = {"u_inv": torch.rand(K, R, N + M, generator=g), "u_val": torch.rand(K, R, N, generator=g),
            "n_obj": torch.randn(K, R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(K, R, M, generator=g)}
    loss = tr.step(rgbs.to(DEV), depth.to(DEV), org.to(DEV), dirs.to(DEV), idx.to(DEV), rand={k: v.to(DEV) for k, v in rand.items()})
    torch.cuda.synchronize()
    # oracle on the same inputs
    pcs, zz, valid, labels = [], [], [], []
    for k in range(K):
        _, _, v, lab, p, z = O.sample_points(rgbs[k], depth[k], org[k], dirs[k], N, M, 0.1, 0.05, {kk: vv[k] for kk, vv in rand.items()})
        pcs.append(p); zz.append(z); valid.append(v); labels.append(lab)
    st = torch.stack
    fc = [torch.stack([tr.param_views(k)[i].cpu() for k in range(K)]) for i in range(28)]
    out = O.train_step(fc, tr.pe_B.cpu(), 2.0, [tr.shape_codes[k].cpu() for k in range(K)], [tr.tex_codes[k].cpu() for k in range(K)],
                       st(pcs), st(zz), depth, rgbs[..., :3], st(labels), st(valid), idx)
    torch.testing.assert_close(loss.cpu(), out["loss"], rtol=LOSS_RTOL, atol=1e-3)
    for k in range(K):
        gv = tr.grad_views(k)
        for i in range(28):
            _grad_ok(gv[i], out["g_fc"][i][k], "p%d" % i)
        _grad_ok(tr.g_B[k], out["g_B"][k], "B")
        _grad_ok(tr.g_shape[k], out["g_shape"][k], "sc")
        _grad_ok(tr.g_tex[k], out["g_texture"][k], "tc")
    _summary("multi_tile", loss, out["loss"])


def test_fused_gradients_well_conditioned_vs_oracle(capsys):
    """Tensor-core path (fp16 operands) against the fp32 oracle on a WELL-CONDITIONED loss: every ray has an invalid depth, so the
    depth term (whose 1/(sqrt(var)+1e-4) weighting amplifies rounding noise without bound) is switched off by the reference's
    empty-mask rule and the loss is 5*colour + 10*opacity.  Bounds (2x the measured values of this seed): loss 2e-4 relative
    (measured 6.9e-5); the WHOLE MLP gradient (all tensors concatenated) relative L2 <= 2e-3 (9e-4); every single gradient tensor
    relative L2 <= 0.05 (0.025, B_layer) and cosine >= 0.9993 (0.99969).  Over 12 weight / data seeds (scripts/seed_sweep.py,
    profiles/r02_parity_seed_sweep.txt): whole gradient 0.10-0.50 %, worst single tensor 1.6-5.8 % (bf16 operands, round 1:
    0.4-2.0 % / 6-14.5 %).  The residual is not plain rounding (the fp32 reference itself is 1e-6 from fp64,
    scripts/fp32_noise_floor.py): the losses are L1 and the units ReLUs, so a ray / unit within the forward error of its
    kink flips the sign / mask of its whole gradient path."""
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    g = torch.Generator().manual_seed(11)
    K, n_obj, R, N, M, L = 2, 4, 1536, 1, 9, 32
    tr = FusedCategoryTrainer(K, n_obj, latent_dim=L, scale=2.0, device=DEV, n_bins_cam2surface=N, n_bins=M, generator=g)
    cats = [O.synth_rays(R, n_obj, g) for _ in range(K)]
    rgbs = torch.stack([c[0] for c in cats]); depth = torch.zeros(K, R)
    org = torch.stack([c[2] for c in cats]); dirs = torch.stack([c[3] for c in cats]); idx = torch.stack([c[4] for c in cats])
    rand = {"u_inv": torch.rand(K, R, N + M, generator=g), "u_val": torch.rand(K, R, N, generator=g),
            "n_obj": torch.randn(K, R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(K, R, M, generator=g)}
    max_d = torch.full((K,), 3.0)
    from category_nerf_reconstruction_official_b200 import ops
    dv = lambda t: t.to(DEV)
    z, _, valid, labels, stats = ops.sample_batched(dv(rgbs), dv(depth), dv(org), dv(dirs), N, M, 0.1, 0.05, {k: dv(v) for k, v in rand.items()},
                                                     max_depth=dv(max_d), want_pcs=False)
    loss = tr.step_from_samples(dv(org), dv(dirs), z, dv(depth), dv(rgbs), valid.view(torch.uint8), dv(idx).to(torch.int32), stats[:, 4:8])
    torch.cuda.synchronize()
    zc = z.cpu()
    pcs = org[:, :, None, :] + dirs[:, :, None, :] * zc[..., None]
    fc = [torch.stack([tr.param_views(k)[i].cpu() for k in range(K)]) for i in range(28)]
    out = O.train_step(fc, tr.pe_B.cpu(), 2.0, [tr.shape_codes[k].cpu() for k in range(K)], [tr.tex_codes[k].cpu() for k in range(K)],
                       pcs, zc, depth, rgbs[..., :3], labels.cpu(), valid.cpu().bool(), idx)
    assert float(out["loss"]) > 1.0
    torch.testing.assert_close(loss.cpu(), out["loss"], rtol=2e-4, atol=1e-5)
    worst_rel, worst_cos = 0.0, 1.0
    errs = []
    def chk(a, e, name):
        nonlocal worst_rel, worst_cos
        a = a.detach().double().cpu().flatten(); e = e.detach().double().cpu().flatten()
        if float(e.norm()) < 1e-12:
            return
        rel = float((a - e).norm() / e.norm()); cos = float(torch.nn.functional.cosine_similarity(a, e, dim=0))
        worst_rel, worst_cos = max(worst_rel, rel), min(worst_cos, cos)
        errs.append("%s %.4f" % (name.split(".")[0], rel))
        assert rel <= 0.05 and cos >= 0.9993, "%s: rel L2 %.4f cos %.5f" % (name, rel, cos)
    names = [n for n, _ in O.codenerf_param_shapes(32, L)]
    for k in range(K):
        gv = tr.grad_views(k)
        for i in range(28):
            chk(gv[i], out["g_fc"][i][k], names[i])
        chk(tr.g_B[k], out["g_B"][k], "B_layer")
        chk(tr.g_shape[k], out["g_shape"][k], "shape_codes")
        chk(tr.g_tex[k], out["g_texture"][k], "texture_codes")
    A = tr.g_params.flatten().cpu().double()
    E = torch.cat([torch.cat([out["g_fc"][i][k].flatten() for i in range(28)]) for k in range(K)]).double()
    whole = float((A - E).norm() / E.norm())
    assert whole <= 2e-3, "whole MLP gradient rel L2 %.4f" % whole
    with capsys.disabled():
        print("\n[well-conditioned parity] whole MLP gradient rel L2 %.4f" % whole)
        print("[well-conditioned parity] loss rel err %.2e, worst gradient rel L2 %.4f, worst cosine %.6f"
              % (abs(float(loss) - float(out["loss"])) / float(out["loss"]), worst_rel, worst_cos))
        print("  per-tensor rel L2 (category 1):", " ".join(errs[len(errs) // 2:]))


def test_eval_points_tensor_core_vs_oracle():
    """cnf_eval_points (forward-only tcgen05 path behind Trainer.eval_points) against the fp32 oracle: ragged point count
    (3 chains x several tiles + a partial tile).  Bound: occupancy 4e-3, colour 1e-3 absolute (fp16 operands, a x10 occupancy logit;
    measured 1.7e-3 / 3.1e-4)."""
    from category_nerf_reconstruction_official_b200 import ops
    g = torch.Generator().manual_seed(21)
    L, P = 32, 128 * 7 + 37
    shapes = O.codenerf_param_shapes(32, L)
    params = O.init_params(shapes, g)
    B = O.icosa_matrix() + 0.01 * torch.randn(21, 3, generator=g)
    sc = O.init_codes(1, L, g)[0]; tc = O.init_codes(1, L, g)[0]
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 3.0
    flat = torch.cat([p.reshape(-1) for p in params])
    occ, rgb = ops.eval_points_fused(pts.to(DEV), flat.to(DEV), B.to(DEV), torch.tensor(2.0, device=DEV), sc.to(DEV), tc.to(DEV))
    torch.cuda.synchronize()
    emb = O.embed(pts[:, None, :], B, 2.0)
    a_o, c_o = O.codenerf_forward(params, emb, sc.reshape(1, 1, L).expand(P, 1, L), tc.reshape(1, 1, L).expand(P, 1, L))
    occ_o = torch.sigmoid(a_o.reshape(P)); c_o = c_o.reshape(P, 3)
    err_o = float((occ.cpu() - occ_o).abs().max()); err_c = float((rgb.cpu() - c_o).abs().max())
    print("eval_points: max abs err occupancy %.2e colour %.2e (occupancy range %.3f..%.3f)" % (err_o, err_c, float(occ_o.min()), float(occ_o.max())))
    assert float(occ_o.max() - occ_o.min()) > 0.05
    assert err_o < 4e-3 and err_c < 1e-3


@pytest.mark.parametrize("R,S,L", [(1000, 14, 64), (333, 96, 32), (257, 1, 32), (70000, 10, 256), (40, 200, 32)])
def test_render_rays_tensor_core_vs_oracle(R, S, L):
    """cnf_render_rays (BASELINE config 5: forward-only dense rendering; S = 96 as category_registration.py:146) against the oracle
    chain embed -> CodeNeRF -> sigmoid -> occupancy_to_termination -> render.  fp16 operands and a x10 occupancy logit: every output
    max-abs 7e-3 over up to 700k points (measured worst 3.5e-3: depth, on a 0..2.5 range), termination weights max-abs 5e-3 (2.4e-3)."""
    from category_nerf_reconstruction_official_b200 import ops
    g = torch.Generator().manual_seed(77 + S)
    params = O.init_params(O.codenerf_param_shapes(32, L), g)
    B = O.icosa_matrix() + 0.01 * torch.randn(21, 3, generator=g)
    sc = O.init_codes(1, L, g)[0]; tc = O.init_codes(1, L, g)[0]
    org = (torch.rand(R, 3, generator=g) - 0.5) * 2.0
    dirs = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    z = torch.sort(torch.rand(R, S, generator=g) * 2.5, dim=-1).values
    flat = torch.cat([p.reshape(-1) for p in params])
    out = ops.render_rays_fused(org.to(DEV), dirs.to(DEV), z.to(DEV), flat.to(DEV), B.to(DEV), torch.tensor(2.0, device=DEV),
                                sc.to(DEV), tc.to(DEV), want_weights=True)
    torch.cuda.synchronize()
    pcs = org[:, None, :] + dirs[:, None, :] * z[..., None]
    a_o, c_o = [], []
    for s0 in range(0, R, 8192):            # oracle in chunks (memory)
        e = O.embed(pcs[s0:s0 + 8192], B, 2.0)
        n = e.shape[0]
        a, c = O.codenerf_forward(params, e, sc.reshape(1, 1, L).expand(n, 1, L), tc.reshape(1, 1, L).expand(n, 1, L))
        a_o.append(a.reshape(n, S)); c_o.append(c.reshape(n, S, 3))
    occ = torch.sigmoid(torch.cat(a_o)); col = torch.cat(c_o)
    w = O.termination(occ)
    D = (w * z).sum(-1); Cc = (w[..., None] * col).sum(-2); Op = w.sum(-1)
    var = (w * (z - D[:, None]) ** 2).sum(-1)
    err_w = float((out["weights"].cpu() - w).abs().max())
    err_D = float((out["depth"].cpu() - D).abs().max()); err_C = float((out["rgb"].cpu() - Cc).abs().max())
    err_O = float((out["opacity"].cpu() - Op).abs().max()); err_v = float((out["var"].cpu() - var).abs().max())
    print("render_rays R=%d S=%d: max abs err weights %.2e depth %.2e colour %.2e opacity %.2e var %.2e (opacity range %.3f..%.3f)"
          % (R, S, err_w, err_D, err_C, err_O, err_v, float(Op.min()), float(Op.max())))
    assert float(Op.max() - Op.min()) > 0.02
    mean_w = float((out["weights"].cpu() - w).abs().mean())
    assert err_w < 5e-3 and mean_w < 5e-4 and err_D < 7e-3 and err_C < 3e-3 and err_O < 5e-3 and err_v < 5e-3
    # properties independent of the MLP: weights are a sub-probability distribution and reproduce the outputs exactly
    wk = out["weights"]
    assert float(wk.min()) >= 0.0 and float(wk.sum(-1).max()) <= 1.0 + 1e-5
    torch.testing.assert_close((wk * z.to(DEV)).sum(-1), out["depth"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(wk.sum(-1), out["opacity"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("K,n_obj,R,N,M,L,empty_inst", [
    (1, 4, 480, 5, 9, 256, None),      # BASELINE config 1: S = 14, L = 256
    (3, 6, 200, 1, 9, 32, 2),          # instance 2 has NO rays in the batch (empty segment), ragged tiles
    (2, 1, 97, 1, 9, 32, None),        # single-instance categories (regulariser skipped), fewer rays than one warp group
    (1, 64, 1024, 3, 5, 64, None),     # exactly one instance chunk of the finalize staging, S = 8
    (1, 150, 2400, 3, 5, 64, 70),      # three instance chunks (64 + 64 + 22); an empty instance in the second chunk
    (2, (130, 7), 1040, 1, 9, 32, None),   # ragged categories: 130 instances next to 7 (n_obj_max = 130)
    (5, 2, 64, 1, 15, 32, None),       # S = 16: two rays per warp
])
def test_fused_step_shapes_vs_oracle(K, n_obj, R, N, M, L, empty_inst):
    """Shape / edge-case sweep of the fused step against the oracle.  These 64-2400-ray fixtures include the ill-conditioned depth
    term (1/(sqrt(var)+1e-4) weights).  Bounds = 2x measured: per tensor rel L2 <= 0.09 (worst 0.044) and cosine >= 0.998
    (0.99903), whole MLP gradient of a category rel L2 <= 0.03.  (The suite seeds torch's global RNG per test, so the measured
    values repeat to atomic-order noise.)"""
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    n_objs = list(n_obj) if isinstance(n_obj, tuple) else [n_obj] * K
    g = torch.Generator().manual_seed(1000 + K * 7 + n_objs[0])
    tr = FusedCategoryTrainer(K, n_objs, latent_dim=L, scale=2.0, device=DEV, n_bins_cam2surface=N, n_bins=M, generator=g)
    cats = [O.synth_rays(R, n_objs[k], g) for k in range(K)]
    rgbs = torch.stack([c[0] for c in cats]); depth = torch.stack([c[1] for c in cats])
    org = torch.stack([c[2] for c in cats]); dirs = torch.stack([c[3] for c in cats]); idx = torch.stack([c[4] for c in cats])
    if empty_inst is not None:
        idx = torch.where(idx == empty_inst, torch.zeros_like(idx), idx)
    rand = {"u_inv": torch.rand(K, R, N + M, generator=g), "u_val": torch.rand(K, R, N, generator=g),
            "n_obj": torch.randn(K, R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(K, R, M, generator=g)}
    loss = tr.step(rgbs.to(DEV), depth.to(DEV), org.to(DEV), dirs.to(DEV), idx.to(DEV), rand={k: v.to(DEV) for k, v in rand.items()})
    torch.cuda.synchronize()
    pcs, zz, valid, labels = [], [], [], []
    for k in range(K):
        _, _, v, lab, p, z = O.sample_points(rgbs[k], depth[k], org[k], dirs[k], N, M, 0.1, 0.05, {kk: vv[k] for kk, vv in rand.items()})
        pcs.append(p); zz.append(z); valid.append(v); labels.append(lab)
    st = torch.stack
    fc = [torch.stack([tr.param_views(k)[i].cpu() for k in range(K)]) for i in range(28)]
    out = O.train_step(fc, tr.pe_B.cpu(), 2.0, [tr.shape_codes[k, :n_objs[k]].cpu() for k in range(K)],
                       [tr.tex_codes[k, :n_objs[k]].cpu() for k in range(K)],
                       st(pcs), st(zz), depth, rgbs[..., :3], st(labels), st(valid), idx)
    assert torch.isfinite(tr.g_params).all() and torch.isfinite(tr.g_shape).all() and torch.isfinite(tr.g_B).all()
    torch.testing.assert_close(loss.cpu(), out["loss"], rtol=LOSS_RTOL, atol=1e-3)
    for k in range(K):
        gv = tr.grad_views(k)
        for i in range(28):
            _grad_ok(gv[i], out["g_fc"][i][k], "p%d" % i)
        _grad_ok(tr.g_B[k], out["g_B"][k], "B")
        _grad_ok(tr.g_shape[k, :n_objs[k]], out["g_shape"][k], "sc")
        _grad_ok(tr.g_tex[k, :n_objs[k]], out["g_texture"][k], "tc")
        _grad_ok(tr.g_params[k], torch.cat([out["g_fc"][i][k].flatten() for i in range(28)]), "whole MLP gradient", 0.03, 0.9995)
        assert float(tr.g_shape[k, n_objs[k]:].abs().sum()) == 0.0 and float(tr.g_tex[k, n_objs[k]:].abs().sum()) == 0.0
    _summary("shapes K=%d n_obj=%s R=%d S=%d L=%d" % (K, n_obj, R, N + M, L), loss, out["loss"])
    if empty_inst is not None:      # an instance without rays only gets the regulariser gradient
        reg_only = 0.0005 * tr.shape_codes[:, empty_inst] / tr.shape_codes[:, empty_inst].norm(dim=-1, keepdim=True)
        torch.testing.assert_close(tr.g_shape[:, empty_inst], reg_only, rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("K,n_obj,R,N,M,inv", [(3, 3, 360, 1, 9, 0.15), (2, 4, 333, 5, 9, 0.4), (1, 2, 97, 0, 12, 0.0), (2, 5, 2000, 3, 16, 0.3),
                                               (1, 1, 8192, 1, 9, 0.15)])
def test_fused_sampler_depths_bit_exact_and_same_step(K, n_obj, R, N, M, inv):
    """cnf_train_step_sampled (sampler inside the fused kernel): the depths the kernel used (z_out) equal the oracle's sampler bit
    for bit (src/scene_cateogries.py:453-546), the stats rows equal the standalone sampler's, and the step's loss / gradients equal
    the two-call path (cnf_sample_fwd + cnf_train_step) on the same draws to fp32 accumulation-order noise."""
    from category_nerf_reconstruction_official_b200 import ops
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    g = torch.Generator().manual_seed(31 + R)
    tr = FusedCategoryTrainer(K, n_obj, latent_dim=32, scale=2.0, device=DEV, n_bins_cam2surface=N, n_bins=M, generator=g)
    cats = [O.synth_rays(R, n_obj, g, invalid_frac=inv) for _ in range(K)]
    rgbs, depth, org, dirs, idx = (torch.stack([c[i] for c in cats]) for i in range(5))
    rand = {"u_inv": torch.rand(K, R, N + M, generator=g), "u_val": torch.rand(K, R, N, generator=g),
            "n_obj": torch.randn(K, R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(K, R, M, generator=g)}
    dv = lambda t: t.to(DEV)
    rd = {k: dv(v) for k, v in rand.items()}
    tr.sample_in_kernel = True                      # (off by default: DESIGN 4)
    assert R <= tr._sampled_max_rays()
    tr.keep_z = True
    loss1 = float(tr.step(dv(rgbs), dv(depth), dv(org), dv(dirs), dv(idx), rand=rd))
    g1 = tr.grad_flat.clone(); z_used = tr.z_used.cpu(); stats1 = tr.stats.cpu()
    for k in range(K):
        _, _, v_o, l_o, _, z_o = O.sample_points(rgbs[k], depth[k], org[k], dirs[k], N, M, 0.1, 0.05, {kk: vv[k] for kk, vv in rand.items()})
        assert torch.equal(z_used[k], z_o), "category %d: in-kernel depths differ from the reference sampler" % k
    _, _, _, _, stats2 = ops.sample_batched(dv(rgbs), dv(depth), dv(org), dv(dirs), N, M, 0.1, 0.05, rd, want_pcs=False)
    assert torch.equal(stats1, stats2.cpu())
    tr.sample_in_kernel = False                     # the two-call path on the same draws
    loss2 = float(tr.step(dv(rgbs), dv(depth), dv(org), dv(dirs), dv(idx), rand=rd))
    assert abs(loss1 - loss2) <= 2e-6 * abs(loss2) + 1e-6
    torch.testing.assert_close(g1, tr.grad_flat, rtol=1e-4, atol=1e-6 * float(tr.grad_flat.abs().max()) + 1e-9)

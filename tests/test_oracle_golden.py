"""CPU: the oracle (oracle/cnf_oracle.py) against golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare
the CUDA path against the oracle and against the same fixtures."""
import numpy as np
import pytest
import torch

from oracle import cnf_oracle as O
from helpers import SAMPLER_CASES, STEP_CASES, golden, gt, step_case


@pytest.mark.parametrize("name", SAMPLER_CASES)
def test_sampler_bit_exact(name):
    p = "sampler/%s/" % name
    N, M = [int(v) for v in golden("sampler")[p + "NM"]]
    rand = {k: gt("sampler", p + k) for k in ("u_inv", "u_val", "n_obj", "u_oth")}
    rgb, d, valid, labels, pcs, z = O.sample_points(gt("sampler", p + "rgbs"), gt("sampler", p + "depth"),
                                                    gt("sampler", p + "origins"), gt("sampler", p + "dirs"),
                                                    N, M, 0.1, 0.05, rand)
    assert torch.equal(z, gt("sampler", p + "z"))
    assert torch.equal(pcs, gt("sampler", p + "pcs"))
    assert torch.equal(valid, gt("sampler", p + "valid"))
    assert torch.equal(labels, gt("sampler", p + "labels"))


def test_stratified_bins_module_level():
    z = O.stratified_bins(0.0, torch.tensor(2.5), 96, gt("sampler", "sampler/strat96/u"))
    assert torch.equal(z, gt("sampler", "sampler/strat96/z"))
    # linspace is not k/n in fp32
    lin = gt("sampler", "sampler/linspace15")
    assert torch.equal(lin, torch.linspace(0, 1, 15))
    assert float(lin[7]) != 0.5


@pytest.mark.parametrize("scale", [2.0, 3.0, 5.0])
def test_embed_bit_exact_and_grads(scale):
    p = "embed/s%g/" % scale
    x = gt("embed", p + "x").requires_grad_(True)
    B = gt("embed", p + "B").requires_grad_(True)
    emb = O.embed(x, B, scale)
    assert torch.equal(emb, gt("embed", p + "emb"))
    (emb * gt("embed", p + "gout")).sum().backward()
    torch.testing.assert_close(B.grad, gt("embed", p + "gB"), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x.grad, gt("embed", p + "gx"), rtol=1e-5, atol=1e-5)


def test_embed_init_table():
    assert torch.equal(O.icosa_matrix(), gt("embed", "embed/B_init"))
    assert list(golden("embed")["embed/state_keys"]) == ["B_layer.weight", "scale"]


def test_param_order_matches_reference():
    names = [str(s) for s in golden("step")["step/param_names"]]
    assert names == [n for n, _ in O.codenerf_param_shapes()]
    c = step_case("K1_L256")
    assert [tuple(t.shape[1:]) for t in c["fc"]] == [s for _, s in O.codenerf_param_shapes(32, 256)]
    names = [str(s) for s in golden("bg")["bg/H64/param_names"]]
    assert names == [n for n, _ in O.occmap_param_shapes()]


@pytest.mark.parametrize("name", STEP_CASES)
def test_train_step(name):
    c = step_case(name)
    out = O.train_step(c["fc"], c["B"], c["scale"], c["sc"], c["tc"], c["pcs"], c["z"], c["gt_depth"],
                       c["gt_rgb"], c["labels"], c["mask_depth"], c["indices"])
    tol = dict(rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(out["alpha"], c["alpha"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["color_out"], c["color"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["loss"], c["loss"], **tol)
    for k, gk in (("depth", "depth"), ("color", "colorl"), ("opacity", "opacity"), ("reg_shape", "reg_shape"),
                  ("reg_texture", "reg_texture")):
        torch.testing.assert_close(out[k], c[gk], **tol)
    for i in range(28):
        torch.testing.assert_close(out["g_fc"][i], c["gfc"][i], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(out["g_B"], c["gB"], rtol=1e-4, atol=2e-6)
    for k in range(c["K"]):
        torch.testing.assert_close(out["g_shape"][k], c["gsc"][k], rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(out["g_texture"][k], c["gtc"][k], rtol=1e-4, atol=2e-6)


def test_bg_step():
    p = "bg/H64/"
    fc = [gt("bg", p + "fc%02d" % i) for i in range(14)]
    out = O.bg_train_step(fc, gt("bg", p + "B"), 5.0, gt("bg", p + "pcs"), gt("bg", p + "z"),
                          gt("bg", p + "gt_depth"), gt("bg", p + "gt_rgb"), gt("bg", p + "labels"),
                          gt("bg", p + "mask_depth"))
    torch.testing.assert_close(out["alpha"], gt("bg", p + "alpha"), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["color_out"], gt("bg", p + "color"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["loss"], gt("bg", p + "loss"), rtol=2e-5, atol=2e-6)
    for i in range(14):
        torch.testing.assert_close(out["g_fc"][i], gt("bg", p + "gfc%02d" % i), rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(out["g_B"], gt("bg", p + "gB"), rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("variant", ["mixed", "nodepth"])
def test_bg_step_h128(variant):
    """The reference's background step at its own size (hidden 128, 1200 rays x (5+9)): oracle vs the reference-written golden."""
    fc = [gt("bg128", "bg128/fc%02d" % i) for i in range(14)]
    p = "bg128/%s/" % variant
    out = O.bg_train_step(fc, gt("bg128", "bg128/B"), 5.0, gt("bg128", p + "pcs"), gt("bg128", p + "z"), gt("bg128", p + "gt_depth"),
                          gt("bg128", p + "gt_rgb"), gt("bg128", p + "labels"), gt("bg128", p + "mask_depth"))
    torch.testing.assert_close(out["loss"], gt("bg128", p + "loss"), rtol=5e-5, atol=2e-6)
    ga = torch.cat([t.reshape(-1) for t in out["g_fc"]]).double(); ge = torch.cat([gt("bg128", p + "gfc%02d" % i).reshape(-1) for i in range(14)]).double()
    assert float((ga - ge).norm() / ge.norm()) < 2e-4           # different summation order over 16800 points; L1 sign flips of ~1-ulp residuals
    assert float((out["g_B"].double() - gt("bg128", p + "gB").double()).norm() / gt("bg128", p + "gB").double().norm()) < 2e-4


@pytest.mark.parametrize("name", ["normal", "empty_depth_cat1"])
def test_loss_and_quirks(name):
    p = "loss/%s/" % name
    a = gt("loss", p + "alpha").requires_grad_(True)
    c = gt("loss", p + "color").requires_grad_(True)
    loss, ld, lcol = O.step_loss(a, c, gt("loss", p + "gt_depth"), gt("loss", p + "gt_color"),
                                 gt("loss", p + "labels"), gt("loss", p + "mask_depth"), gt("loss", p + "z"))
    loss.backward()
    torch.testing.assert_close(loss, gt("loss", p + "loss"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ld["depth"], gt("loss", p + "depth"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ld["color"], gt("loss", p + "colorl"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ld["opacity"], gt("loss", p + "opacity"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a.grad, gt("loss", p + "galpha"), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(c.grad, gt("loss", p + "gcolor"), rtol=1e-4, atol=1e-7)
    if name == "empty_depth_cat1":     # quirk: the depth term is zero for ALL categories
        assert float(ld["depth"].abs().max()) == 0.0
    term = O.termination(torch.sigmoid(gt("loss", p + "alpha").squeeze(-1)))
    torch.testing.assert_close(term, gt("loss", p + "term"), rtol=1e-6, atol=1e-7)


def test_termination_2d():
    t = O.termination(gt("loss", "loss/term2d/occ"))
    torch.testing.assert_close(t, gt("loss", "loss/term2d/term"), rtol=1e-6, atol=1e-7)


def test_closed_form_backward_fp64():
    """Appendix-A closed form vs autograd in fp64, per category (incl. empty-mask category)."""
    p = "loss/normal/"
    K = gt("loss", p + "alpha").shape[0]
    for k in range(K):
        a = gt("loss", p + "alpha")[k].squeeze(-1).double().requires_grad_(True)
        c = gt("loss", p + "color")[k].double().requires_grad_(True)
        args = (gt("loss", p + "gt_depth")[k].double(), gt("loss", p + "gt_color")[k].double(),
                gt("loss", p + "labels")[k], gt("loss", p + "mask_depth")[k], gt("loss", p + "z")[k].double())
        loss, ld, _ = O.step_loss(a[None, ..., None], c[None], args[0][None], args[1][None], args[2][None],
                                  args[3][None], args[4][None])
        loss.backward()
        cf = O.composite_closed_form(a.detach(), c.detach(), args[4], args[0], args[1], args[2], args[3])
        torch.testing.assert_close(cf["loss"], loss.detach(), rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(cf["dalpha"], a.grad, rtol=1e-9, atol=1e-13)
        torch.testing.assert_close(cf["dcolor"], c.grad, rtol=1e-9, atol=1e-13)


def test_curve_first_steps_reproduce_reference_trajectory():
    """The oracle + torch.optim.AdamW reproduce the first optimizer steps of the reference's own training loop (golden_curve in
    tests/golden/make_golden.py: reference sampler / modules / losses, functorch vmap, AdamW with the config's groups), replaying
    the reference's RNG stream for the sampler.  Pins the step-to-step semantics (gradient -> AdamW -> next loss), not only one step."""
    G = golden("curve")
    K, n_obj, R, N, M, L, steps, nb = [int(v) for v in G["curve/cfg"]]
    scale, lr, wd, clr, cwd = [float(v) for v in G["curve/hyper"]]
    traj = G["curve/traj"]
    fc = [gt("curve", "curve/fc%02d" % i).clone().requires_grad_(True) for i in range(28)]
    B = gt("curve", "curve/B").clone().requires_grad_(True)
    sc = [gt("curve", "curve/sc%d" % k).clone().requires_grad_(True) for k in range(K)]
    tc = [gt("curve", "curve/tc%d" % k).clone().requires_grad_(True) for k in range(K)]
    opt = torch.optim.AdamW([{"params": fc + [B], "lr": lr, "weight_decay": wd}, {"params": sc + tc, "lr": clr, "weight_decay": cwd}])
    for it in range(4):
        g = torch.Generator().manual_seed(9000 + it % nb)
        cats = [O.synth_rays(R, n_obj, g) for _ in range(K)]
        rgbs, depth, org, dirs, idx = (torch.stack([c[j] for c in cats]) for j in range(5))
        torch.manual_seed(7000 + it)
        pcs, zz, valid, labels = [], [], [], []
        for k in range(K):
            r = O.draw_sampler_randoms(rgbs[k], depth[k], N, M, 0.1)
            _, _, v, lab, p, z = O.sample_points(rgbs[k], depth[k], org[k], dirs[k], N, M, 0.1, 0.05, r)
            pcs.append(p); zz.append(z); valid.append(v); labels.append(lab)
        st = torch.stack
        out = O.train_step(fc, B, scale, sc, tc, st(pcs), st(zz), depth, rgbs[..., :3], st(labels), st(valid), idx)
        assert abs(float(out["loss"]) - float(traj[it, 0])) <= 2e-4 * abs(float(traj[it, 0])), (it, float(out["loss"]), float(traj[it, 0]))
        np.testing.assert_allclose(out["depth"].numpy(), traj[it, 1:1 + K], rtol=5e-4)
        np.testing.assert_allclose(out["color"].numpy(), traj[it, 1 + K:1 + 2 * K], rtol=5e-4)
        for t, gr in zip(fc + [B] + sc + tc, out["g_fc"] + [out["g_B"]] + out["g_shape"] + out["g_texture"]):
            t.grad = gr
        opt.step()

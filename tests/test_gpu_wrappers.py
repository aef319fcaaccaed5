"""GPU: the reference-facing wrapper classes, called the way the reference's callers call them (never executed).

  * trainer.Trainer built from a cfg stub (src/trainer.py:11-60) -> eval_points (src/trainer.py:125-151; train.py:221,232 meshing)
    for an object category and for the background model, render_rays (BASELINE config 5), against the oracle;
  * sampling.RaySampler.sample_3d_points (scene.sample_3d_points(rgbs, depth, origins, dirs), scene_cateogries.py:453) incl. the
    reference's own RNG order (exact_rng) against the reference golden;
  * render_rays.occupancy_to_termination is differentiable: its backward against autograd of the oracle expression;
  * checkpoint export / import of the fused trainers round-trips through the module-level Trainer."""
import numpy as np
import pytest
import torch

from helpers import SAMPLER_CASES, golden, gt
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Cfg:
    training_device = DEV
    obj_scale = 2.0
    n_unidir_funcs = 5
    hidden_feature_size = 128
    net_hyperparams = {"shape_blocks": 2, "texture_blocks": 1, "W": 32, "latent_dim": 64}


def _oracle_eval(tr, pts, inst_id):
    params = [p.detach().cpu() for p in tr.fc_occ_map.parameters()]
    emb = O.embed(pts[:, None, :], tr.pe.B_layer.weight.detach().cpu(), float(tr.pe.scale))
    if tr.cls_id == 0:
        a, c = O.occmap_forward(params, emb.reshape(-1, 129))
    else:
        P = pts.shape[0]
        row = tr.inst_id_to_index[inst_id]
        sc = tr.shape_codes.weight[row].detach().cpu(); tc = tr.texture_codes.weight[row].detach().cpu()
        a, c = O.codenerf_forward(params, emb, sc.reshape(1, 1, -1).expand(P, 1, -1), tc.reshape(1, 1, -1).expand(P, 1, -1))
    return torch.sigmoid(a.reshape(-1)), c.reshape(-1, 3)


@pytest.mark.parametrize("cls_id", [3, 0])
def test_trainer_eval_points_and_render_rays(cls_id):
    from category_nerf_reconstruction_official_b200 import trainer
    torch.manual_seed(21 + cls_id)
    tr = trainer.Trainer(Cfg, cls_id, [17, 4, 9] if cls_id else [0])
    g = torch.Generator().manual_seed(5)
    P = 128 * 9 + 41
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 3.0
    inst = 4 if cls_id else None
    occ_o, col_o = _oracle_eval(tr, pts, inst)
    for tc_on in (True, False):                     # tensor-core kernels, then the fp32 CUDA-core module path
        tr.use_tensor_cores = tc_on
        tr.fc_occ_map.use_tensor_cores = tc_on
        occ, col = tr.eval_points(pts.to(DEV), inst_id=inst, chunk_size=500)
        assert occ.shape == (P,) and col.shape == (P, 3)
        e_o, e_c = float((occ.cpu() - occ_o).abs().max()), float((col.cpu() - col_o).abs().max())
        print("\n[Trainer.eval_points cls %d tensor_cores=%s] max abs err occupancy %.2e colour %.2e" % (cls_id, tc_on, e_o, e_c))
        assert (e_o < 4e-3 and e_c < 1e-3) if tc_on else (e_o < 2e-5 and e_c < 2e-5)
    # render_rays: dense view through one instance's field
    tr.use_tensor_cores = True
    R, S = 300, 14
    org = (torch.rand(R, 3, generator=g) - 0.5) * 2.0
    dirs = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    z = torch.sort(torch.rand(R, S, generator=g) * 2.5, dim=-1).values
    out = tr.render_rays(org.to(DEV), dirs.to(DEV), z.to(DEV), inst, want_weights=True)
    pcs = (org[:, None, :] + dirs[:, None, :] * z[..., None]).reshape(-1, 3)
    occ_o, col_o = _oracle_eval(tr, pcs, inst)
    w = O.termination(occ_o.reshape(R, S))
    D = (w * z).sum(-1); C = (w[..., None] * col_o.reshape(R, S, 3)).sum(-2)
    assert float((out["weights"].cpu() - w).abs().max()) < 5e-3
    assert float((out["depth"].cpu() - D).abs().max()) < 7e-3 and float((out["rgb"].cpu() - C).abs().max()) < 3e-3


@pytest.mark.parametrize("name", SAMPLER_CASES)
def test_ray_sampler_class_matches_reference_rng_stream(name):
    """RaySampler.sample_3d_points(rgbs, depth, origins, dirs) with exact_rng=True draws torch's randoms with the reference's
    shapes and order; with the draws of the golden run injected (``rand=``) the samples are bit-exact."""
    from category_nerf_reconstruction_official_b200.sampling import RaySampler
    p = "sampler/" + name + "/"
    N, M = [int(v) for v in golden("sampler")[p + "NM"]]
    sc = RaySampler(N, M, 0.1, 0.05, data_device=DEV)
    dv = lambda key: gt("sampler", p + key).to(DEV)
    rand = {k: gt("sampler", p + k) for k in ("u_inv", "u_val", "n_obj", "u_oth")}
    rgb, d, valid, labels, pcs, z = sc.sample_3d_points(dv("rgbs"), dv("depth"), dv("origins"), dv("dirs"), rand=rand)
    assert torch.equal(z.cpu(), gt("sampler", p + "z")) and torch.equal(pcs.cpu(), gt("sampler", p + "pcs"))
    assert torch.equal(valid.cpu(), gt("sampler", p + "valid")) and torch.equal(labels.cpu(), gt("sampler", p + "labels"))
    assert rgb.shape[-1] == 3 and rgb.dtype == torch.uint8
    # own draws (device RNG): right distribution / ranges, 4 syncs as the reference
    rgb, d, valid, labels, pcs, z2 = sc.sample_3d_points(dv("rgbs"), dv("depth"), dv("origins"), dv("dirs"), exact_rng=True)
    assert z2.shape == z.shape and torch.isfinite(z2).all()
    v = valid.cpu()
    if bool(v.any()):
        dd = gt("sampler", p + "depth")[v]
        assert bool((z2.cpu()[v][:, :N] <= (dd - 0.1)[:, None] + 1e-6).all())          # camera -> surface bins end at depth - eps
        assert bool((z2.cpu()[v][:, N:] >= (dd - 0.1)[:, None] - 1e-6).all())          # surface band starts at depth - eps


def test_occupancy_to_termination_is_differentiable():
    from category_nerf_reconstruction_official_b200 import render_rays
    g = torch.Generator().manual_seed(2)
    for shape in ((57, 14), (3, 21, 96), (5, 200)):
        occ = torch.rand(*shape, generator=g)
        occ[..., 3] = 1.0                                       # a saturated sample: f = 1e-10
        gw = torch.randn(*shape, generator=g)
        o0 = occ.clone().requires_grad_(True)
        (O.termination(o0) * gw).sum().backward()
        o1 = occ.to(DEV).requires_grad_(True)
        w = render_rays.occupancy_to_termination(o1, is_batch=len(shape) == 3)
        (w * gw.to(DEV)).sum().backward()
        torch.testing.assert_close(w.detach().cpu(), O.termination(occ), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(o1.grad.cpu(), o0.grad, rtol=1e-4, atol=1e-6)


def test_fused_trainer_checkpoint_round_trip(tmp_path):
    """FusedCategoryTrainer -> reference-shaped checkpoint file -> module-level Trainer (eval_points agrees with the fused
    trainer's own state) -> back into a fresh FusedCategoryTrainer, bit-identical; same for the background trainer."""
    from category_nerf_reconstruction_official_b200 import checkpoint, trainer
    from category_nerf_reconstruction_official_b200.background import FusedBackgroundTrainer
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    g = torch.Generator().manual_seed(8)
    ft = FusedCategoryTrainer(2, [3, 2], latent_dim=64, scale=2.0, device=DEV, generator=g)
    f = checkpoint.save_category(ft, 1, str(tmp_path), cls_id=12, iteration=4000, instance_id_to_index={40: 0, 41: 1})
    assert f.endswith("cls_12_iteration_04000.pth")
    ck = torch.load(f, map_location="cpu", weights_only=False)
    assert ck["shape_code_state_dict"]["weight"].shape == (2, 64) and ck["global_step"] == 4000
    torch.manual_seed(1)
    mt = trainer.Trainer(Cfg, 12, [40, 41])
    assert checkpoint.load_module_checkpoint(mt, ck) == 4000
    for v, p in zip(ft.param_views(1), mt.fc_occ_map.parameters()):
        assert torch.equal(v, p.detach().to(DEV))
    ft2 = FusedCategoryTrainer(2, [3, 2], latent_dim=64, scale=2.0, device=DEV, generator=torch.Generator().manual_seed(99))
    assert checkpoint.load_category_file(ft2, 1, f) == 4000
    assert torch.equal(ft2.params[1], ft.params[1]) and torch.equal(ft2.shape_codes[1, :2], ft.shape_codes[1, :2])
    assert torch.equal(ft2.pe_B[1], ft.pe_B[1]) and not torch.equal(ft2.params[0], ft.params[0])
    bt = FusedBackgroundTrainer(scale=5.0, device=DEV, generator=g)
    bck = checkpoint.export_background(bt, 10)
    mb = trainer.Trainer(Cfg, 0, [0])
    checkpoint.load_module_checkpoint(mb, bck)
    bt2 = FusedBackgroundTrainer(scale=5.0, device=DEV, generator=torch.Generator().manual_seed(3))
    checkpoint.import_background(bt2, checkpoint.module_checkpoint(mb, 0, 10))
    assert torch.equal(bt2.params, bt.params) and torch.equal(bt2.pe_B, bt.pe_B)


def test_host_step_pipeline_orders_results_and_matches_direct_steps():
    """fused.HostStepPipeline (host batches, one step in flight): results come back in submission order, each equal -- up to the
    sampler's random offsets -- to a direct FusedCategoryTrainer.step on the same batch; slots are not overwritten before they are read."""
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer, HostStepPipeline
    from category_nerf_reconstruction_official_b200.synthetic import synth_rays
    g = torch.Generator().manual_seed(11)
    K, n_obj, R = 3, 2, 240
    tr = FusedCategoryTrainer(K, n_obj, latent_dim=32, scale=2.0, device=torch.device(DEV), n_bins_cam2surface=1, n_bins=9, generator=g)

    def batch(kind):
        cats = [synth_rays(R, n_obj, g) for _ in range(K)]
        b = {k: torch.stack([c[i] for c in cats]) for i, k in enumerate(HostStepPipeline.KEYS)}
        # the empty-mask flags out_loss[1:4] (render_rays.py:67-72) are a deterministic signature of a batch
        if kind == "no_depth":
            b["depth"] = torch.zeros_like(b["depth"])
        elif kind == "no_obj":
            b["rgbs"][..., 3] = 0
        elif kind == "no_sem":
            b["rgbs"][..., 3] = 2
        return b

    kinds = ["plain", "no_depth", "no_obj", "no_sem", "plain"]
    flags = {"plain": [0, 0, 0], "no_depth": [1, 0, 0], "no_obj": [1, 1, 0], "no_sem": [0, 0, 1]}
    batches = [batch(k) for k in kinds]
    direct = []
    for b in batches:
        loss = tr.step(*(b[k].to(DEV) for k in HostStepPipeline.KEYS))
        direct.append(float(loss))
    pipe = HostStepPipeline(tr, batches[0])
    got = []
    for i, b in enumerate(batches):
        s = i & 1
        for k in HostStepPipeline.KEYS:
            pipe.host[s][k].copy_(b[k])
        pipe.submit(s)
        if i:
            got.append(pipe.result().clone())
    got.append(pipe.result().clone())
    for o, kind, d in zip(got, kinds, direct):
        assert [int(v) for v in o[1:4]] == flags[kind], (kind, o[:4])
        assert abs(float(o[0]) - d) <= 0.35 * abs(d) + 1e-3, (kind, float(o[0]), d)     # the sampler's offsets differ between the two runs

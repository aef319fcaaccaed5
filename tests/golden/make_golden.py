"""Generate golden vectors from the UNMODIFIED reference modules (run in the authoring container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (``/root/reference``) does not exist on the GPU box, so the vectors produced here are
committed and the tests read only the ``.npz`` files.  Inputs come from ``oracle.cnf_oracle``'s
synthetic generators (pure data) with fixed seeds; every OUTPUT stored here is computed by the
reference's own code: ``src/scene_cateogries.py`` (sampler), ``src/embedding.py``, ``src/model.py``,
``src/render_rays.py``, ``src/loss.py``, batched with functorch exactly as ``train.py:88-89,154-167`` does.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("CNF_REFERENCE", "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


for _n in ["skimage", "skimage.measure", "trimesh", "open3d", "imgviz", "plotly", "plotly.graph_objs",
           "plotly.subplots"]:
    sys.modules[_n] = _Stub(_n)
sys.path.insert(0, os.path.join(REF, "src"))

import embedding as ref_embedding  # noqa: E402
import loss as ref_loss  # noqa: E402
import model as ref_model  # noqa: E402
import render_rays as ref_render  # noqa: E402
import scene_cateogries as ref_scene  # noqa: E402
from functorch import combine_state_for_ensemble, vmap  # noqa: E402

from oracle import cnf_oracle as O  # noqa: E402  (input generators only)


def npy(t):
    return t.detach().cpu().numpy()


def bare_scene(N, M, eps=0.1, stop_eps=0.05):
    sc = ref_scene.sceneCategory.__new__(ref_scene.sceneCategory)
    sc.n_bins_cam2surface = N
    sc.n_bins = M
    sc.surface_eps = eps
    sc.stop_eps = stop_eps
    sc.data_device = "cpu"
    sc.min_bound = 0.0
    sc.this_obj = 1
    return sc


def golden_sampler(out):
    cases = [("mixed_5_9", 96, 5, 9, 0.15, (0.6, 0.3, 0.1)),
             ("mixed_1_9", 80, 1, 9, 0.15, (0.6, 0.3, 0.1)),
             ("allvalid_obj", 40, 5, 9, 0.0, (1.0, 0.0, 0.0)),
             ("allinvalid", 24, 1, 9, 1.0, (0.6, 0.3, 0.1)),
             ("noobj", 33, 5, 9, 0.2, (0.0, 0.7, 0.3))]
    for name, R, N, M, inv, mix in cases:
        g = torch.Generator().manual_seed(100 + R)
        rgbs, depth, origins, dirs, idx = O.synth_rays(R, 4, g, invalid_frac=inv, label_mix=mix)
        sc = bare_scene(N, M)
        torch.manual_seed(7)
        rgb, d, valid, labels, pcs, z = sc.sample_3d_points(rgbs, depth, origins, dirs)
        # replay the global RNG stream to record what the reference drew (same shapes, same order)
        torch.manual_seed(7)
        rand = O.draw_sampler_randoms(rgbs, depth, N, M, 0.1)
        p = "sampler/" + name + "/"
        out[p + "NM"] = np.array([N, M])
        for k, v in dict(rgbs=rgbs, depth=depth, origins=origins, dirs=dirs, valid=valid, labels=labels,
                         pcs=pcs, z=z, **rand).items():
            out[p + k] = npy(v)
    # module-level stratified_bins with tensor min/max and python-float min (category_registration.py:146 use)
    torch.manual_seed(3)
    zz = ref_scene.stratified_bins(0.0, torch.tensor(2.5), 96, 16, device="cpu")
    torch.manual_seed(3)
    u = torch.rand(16, 96)
    out["sampler/strat96/u"] = npy(u)
    out["sampler/strat96/z"] = npy(zz)
    out["sampler/linspace15"] = npy(torch.linspace(0, 1, 15))


def golden_embed(out):
    g = torch.Generator().manual_seed(11)
    for scale in (2.0, 3.0, 5.0):
        pe = ref_embedding.UniDirsEmbed(max_deg=5, scale=scale)
        x = (torch.rand(24, 5, 3, generator=g) * 2 - 1) * scale * 0.9
        Bw = pe.B_layer.weight.data.clone() + 0.01 * torch.randn(21, 3, generator=g)
        pe.B_layer.weight.data = Bw.clone()
        x.requires_grad_(True)
        emb = pe(x)
        gout = torch.randn(emb.shape, generator=g)
        (emb * gout).sum().backward()
        p = "embed/s%g/" % scale
        out[p + "x"] = npy(x); out[p + "B"] = npy(Bw); out[p + "emb"] = npy(emb)
        out[p + "gout"] = npy(gout); out[p + "gB"] = npy(pe.B_layer.weight.grad); out[p + "gx"] = npy(x.grad)
    pe = ref_embedding.UniDirsEmbed(max_deg=5, scale=2.0)
    out["embed/state_keys"] = np.array(sorted(pe.state_dict().keys()))
    out["embed/B_init"] = npy(pe.B_layer.weight.data)


def make_batch(K, n_obj, R, N, M, seed, inv=0.15, mix=(0.6, 0.3, 0.1)):
    g = torch.Generator().manual_seed(seed)
    sc = bare_scene(N, M)
    pcs, zs, gd, grgb, dm, om, idxs, orgs, drs = [], [], [], [], [], [], [], [], []
    for k in range(K):
        rgbs, depth, origins, dirs, idx = O.synth_rays(R, n_obj, g, invalid_frac=inv, label_mix=mix)
        torch.manual_seed(seed * 31 + k)
        rgb, d, valid, labels, pc, z = sc.sample_3d_points(rgbs, depth, origins, dirs)
        pcs.append(pc); zs.append(z); gd.append(d); grgb.append(rgb); dm.append(valid); om.append(labels)
        idxs.append(idx); orgs.append(origins); drs.append(dirs)
    st = torch.stack
    make_batch.last_rays = (st(orgs), st(drs))
    return st(pcs), st(zs), st(gd), st(grgb), st(dm), st(om), st(idxs)


class _FakeTrainer:
    pass


class _FakeCls:
    pass


def golden_train_step(out):
    """Full object-category step as train.py:136-182 (vmap over K categories) for several shapes."""
    cases = [("K2_L32", 2, 3, 48, 5, 9, 32, 2.0), ("K1_L256", 1, 4, 60, 5, 9, 256, 2.0),
             ("K3_L32_S10", 3, 2, 40, 1, 9, 32, 3.0), ("K2_single_inst", 2, 1, 24, 1, 9, 32, 2.0)]
    for name, K, n_obj, R, N, M, L, scale in cases:
        torch.manual_seed(1234 + K + L)
        fcs, pes, scs, tcs = [], [], [], []
        for k in range(K):
            m = ref_model.CodeNeRF(O.EMB_SIZE1, O.EMB_SIZE2, shape_blocks=2, texture_blocks=1, W=32, latent_dim=L)
            m.apply(ref_model.init_weights)
            fcs.append(m)
            pe = ref_embedding.UniDirsEmbed(max_deg=5, scale=scale)
            pe.B_layer.weight.data = pe.B_layer.weight.data + 0.02 * torch.randn(21, 3)
            pes.append(pe)
            scs.append(torch.nn.Parameter(torch.randn(n_obj, L) / np.sqrt(L / 2)))
            tcs.append(torch.nn.Parameter(torch.randn(n_obj, L) / np.sqrt(L / 2)))
        fc_model, fc_param, fc_buffer = combine_state_for_ensemble(fcs)
        [p.requires_grad_() for p in fc_param]
        pe_model, pe_param, pe_buffer = combine_state_for_ensemble(pes)
        [p.requires_grad_() for p in pe_param]
        pcs, z, gd, grgb, dm, om, idx = make_batch(K, n_obj, R, N, M, seed=77 + K)
        sh = torch.stack([scs[k][idx[k]][:, None, :] for k in range(K)])
        tx = torch.stack([tcs[k][idx[k]][:, None, :] for k in range(K)])
        emb = vmap(pe_model)(pe_param, pe_buffer, pcs)
        alpha, color = vmap(fc_model)(fc_param, fc_buffer, emb, sh, tx)
        loss, ld, lcol = ref_loss.step_batch_loss(alpha, color, gd, grgb / 255.0, om, dm, z)
        cls_dict = {}
        for k in range(K):
            c = _FakeCls(); c.trainer = _FakeTrainer(); c.trainer.n_obj = n_obj; c.training_device = "cpu"
            c.trainer.shape_codes = (lambda t: (lambda ids: t[ids]))(scs[k])
            c.trainer.texture_codes = (lambda t: (lambda ids: t[ids]))(tcs[k])
            c.obj_ids = list(range(n_obj))
            cls_dict[k + 1] = c
        rs, rt = ref_loss.step_batch_loss_reg(cls_dict, torch.arange(K))
        loss = loss + 0.0005 * (rs + rt).sum()
        loss.backward()
        p = "step/" + name + "/"
        out[p + "cfg"] = np.array([K, n_obj, R, N, M, L]); out[p + "scale"] = np.array(scale)
        for i, t in enumerate(fc_param):
            out[p + "fc%02d" % i] = npy(t); out[p + "gfc%02d" % i] = npy(t.grad)
        out[p + "B"] = npy(pe_param[0]); out[p + "gB"] = npy(pe_param[0].grad)
        for k in range(K):
            out[p + "sc%d" % k] = npy(scs[k]); out[p + "tc%d" % k] = npy(tcs[k])
            out[p + "gsc%d" % k] = npy(scs[k].grad); out[p + "gtc%d" % k] = npy(tcs[k].grad)
        for kk, v in dict(pcs=pcs, z=z, gt_depth=gd, gt_rgb=grgb, mask_depth=dm, labels=om, indices=idx,
                          alpha=alpha, color=color, loss=loss, depth=ld["depth"], colorl=ld["color"],
                          opacity=ld["opacity"], reg_shape=rs, reg_texture=rt, origins=make_batch.last_rays[0],
                          dirs=make_batch.last_rays[1]).items():
            out[p + kk] = npy(v)
    names = [n for n, _ in fcs[0].named_parameters()]
    out["step/param_names"] = np.array(names)


def golden_bg(out):
    torch.manual_seed(99)
    H = 64
    m = ref_model.OccupancyMap(O.EMB_SIZE1, O.EMB_SIZE2, hidden_size=H)
    m.apply(ref_model.init_weights)
    pe = ref_embedding.UniDirsEmbed(max_deg=5, scale=5.0)
    pcs, z, gd, grgb, dm, om, idx = make_batch(1, 1, 50, 5, 9, seed=5)
    emb = pe(pcs[0])
    a, c = m(emb)
    loss, ld, lcol = ref_loss.step_batch_loss(a[None], c[None], gd, grgb / 255.0, om, dm, z)
    loss.backward()
    p = "bg/H64/"
    for i, (n, t) in enumerate(m.named_parameters()):
        out[p + "fc%02d" % i] = npy(t); out[p + "gfc%02d" % i] = npy(t.grad)
    out[p + "param_names"] = np.array([n for n, _ in m.named_parameters()])
    out[p + "B"] = npy(pe.B_layer.weight); out[p + "gB"] = npy(pe.B_layer.weight.grad)
    for kk, v in dict(pcs=pcs[0], z=z[0], gt_depth=gd[0], gt_rgb=grgb[0], mask_depth=dm[0], labels=om[0],
                      alpha=a, color=c, loss=loss, depth=ld["depth"], colorl=ld["color"],
                      opacity=ld["opacity"]).items():
        out[p + kk] = npy(v)


def golden_bg128(out):
    """The reference's background step at its own size (train.py:172-180: OccupancyMap hidden 128, 1200 rays x (5+9) samples, scale 5),
    un-vmapped, fp32.  Two variants: 'mixed' (the default ray mix, depth term on) and 'nodepth' (every depth invalid, so the depth term
    is switched off by the empty-mask rule and the loss is well conditioned).  Parameters are shared by the variants."""
    torch.manual_seed(4242)
    H, R, N, M = 128, 1200, 5, 9
    m = ref_model.OccupancyMap(O.EMB_SIZE1, O.EMB_SIZE2, hidden_size=H)
    m.apply(ref_model.init_weights)
    pe = ref_embedding.UniDirsEmbed(max_deg=5, scale=5.0)
    pe.B_layer.weight.data = pe.B_layer.weight.data + 0.02 * torch.randn(21, 3)
    out["bg128/cfg"] = np.array([H, R, N, M]); out["bg128/scale"] = np.array(5.0)
    out["bg128/param_names"] = np.array([n for n, _ in m.named_parameters()])
    for i, (n, t) in enumerate(m.named_parameters()):
        out["bg128/fc%02d" % i] = npy(t)
    out["bg128/B"] = npy(pe.B_layer.weight)
    for name, inv in (("mixed", 0.15), ("nodepth", 1.0)):
        m.zero_grad(set_to_none=True); pe.zero_grad(set_to_none=True)
        pcs, z, gd, grgb, dm, om, idx = make_batch(1, 1, R, N, M, seed=500 + int(inv * 10), inv=inv)
        emb = pe(pcs[0])
        a, c = m(emb)
        loss, ld, lcol = ref_loss.step_batch_loss(a[None], c[None], gd, grgb / 255.0, om, dm, z)
        loss.backward()
        p = "bg128/" + name + "/"
        for i, (n, t) in enumerate(m.named_parameters()):
            out[p + "gfc%02d" % i] = npy(t.grad)
        out[p + "gB"] = npy(pe.B_layer.weight.grad)
        for kk, v in dict(pcs=pcs[0], z=z[0], gt_depth=gd[0], gt_rgb=grgb[0], mask_depth=dm[0], labels=om[0],
                          loss=loss, depth=ld["depth"], colorl=ld["color"], opacity=ld["opacity"]).items():
            out[p + kk] = npy(v)
        if name == "mixed":
            out[p + "alpha"] = npy(a).astype(np.float16); out[p + "color"] = npy(c).astype(np.float16)   # coarse forward check only


CURVE = dict(K=2, n_obj=2, R=240, N=1, M=9, L=32, scale=2.0, steps=1000, n_batches=4, lr=1e-3, wd=0.013, code_lr=1e-3, code_wd=0.013)


def curve_batches():
    """The fixed ray batches of the loss-curve fixture (pure data from the oracle's generator; regenerated identically by the test)."""
    c = CURVE
    out = []
    for i in range(c["n_batches"]):
        g = torch.Generator().manual_seed(9000 + i)
        cats = [O.synth_rays(c["R"], c["n_obj"], g) for _ in range(c["K"])]
        out.append(tuple(torch.stack([cat[j] for cat in cats]) for j in range(5)))
    return out


def golden_curve(out):
    """1000 optimizer steps of the reference's training loop (train.py:40,54-64,88-89,123-184) on CPU: reference sampler, functorch vmap
    of the reference PE / CodeNeRF, reference losses + regulariser, torch.optim.AdamW with the config's groups.  Step `it` seeds
    torch's global RNG with 7000 + it before sampling, so the test replays the same random offsets.  Stores the initial state and the
    loss trajectory (total + the three terms per category)."""
    c = CURVE
    K, n_obj, R, N, M, L = c["K"], c["n_obj"], c["R"], c["N"], c["M"], c["L"]
    torch.manual_seed(31337)
    fcs, pes, scs, tcs = [], [], [], []
    for k in range(K):
        m = ref_model.CodeNeRF(O.EMB_SIZE1, O.EMB_SIZE2, shape_blocks=2, texture_blocks=1, W=32, latent_dim=L)
        m.apply(ref_model.init_weights)
        fcs.append(m)
        pes.append(ref_embedding.UniDirsEmbed(max_deg=5, scale=c["scale"]))
        scs.append(torch.nn.Parameter(torch.randn(n_obj, L) / np.sqrt(L / 2)))
        tcs.append(torch.nn.Parameter(torch.randn(n_obj, L) / np.sqrt(L / 2)))
    opt = torch.optim.AdamW([torch.autograd.Variable(torch.tensor(0))], lr=c["lr"], weight_decay=c["wd"])      # train.py:40
    for k in range(K):
        opt.add_param_group({"params": [scs[k]], "lr": c["code_lr"], "weight_decay": c["code_wd"]})                # train.py:63-64
        opt.add_param_group({"params": [tcs[k]], "lr": c["code_lr"], "weight_decay": c["code_wd"]})
    fc_model, fc_param, fc_buffer = combine_state_for_ensemble(fcs)                                                  # utils.py:24-28
    [p.requires_grad_() for p in fc_param]
    opt.add_param_group({"params": fc_param})
    pe_model, pe_param, pe_buffer = combine_state_for_ensemble(pes)
    [p.requires_grad_() for p in pe_param]
    opt.add_param_group({"params": pe_param})
    out["curve/cfg"] = np.array([K, n_obj, R, N, M, L, c["steps"], c["n_batches"]])
    out["curve/hyper"] = np.array([c["scale"], c["lr"], c["wd"], c["code_lr"], c["code_wd"]])
    for i, t in enumerate(fc_param):                 # copies: the optimizer updates these tensors in place
        out["curve/fc%02d" % i] = npy(t).copy()
    out["curve/B"] = npy(pe_param[0]).copy()
    for k in range(K):
        out["curve/sc%d" % k] = npy(scs[k]).copy(); out["curve/tc%d" % k] = npy(tcs[k]).copy()
    batches = curve_batches()
    sc = bare_scene(N, M)
    cls_dict = {}
    for k in range(K):
        cc = _FakeCls(); cc.trainer = _FakeTrainer(); cc.trainer.n_obj = n_obj; cc.training_device = "cpu"
        cc.trainer.shape_codes = (lambda t: (lambda ids: t[ids]))(scs[k])
        cc.trainer.texture_codes = (lambda t: (lambda ids: t[ids]))(tcs[k])
        cc.obj_ids = list(range(n_obj))
        cls_dict[k + 1] = cc
    traj = np.zeros((c["steps"], 1 + 3 * K), dtype=np.float32)
    for it in range(c["steps"]):
        rgbs, depth, org, dirs, idx = batches[it % c["n_batches"]]
        torch.manual_seed(7000 + it)
        pcs, zs, gd, grgb, dm, om = [], [], [], [], [], []
        for k in range(K):
            rgb, d, valid, labels, pc, z = sc.sample_3d_points(rgbs[k], depth[k], org[k], dirs[k])
            pcs.append(pc); zs.append(z); gd.append(d); grgb.append(rgb); dm.append(valid); om.append(labels)
        st = torch.stack
        sh = st([scs[k][idx[k]][:, None, :] for k in range(K)]); tx = st([tcs[k][idx[k]][:, None, :] for k in range(K)])
        emb = vmap(pe_model)(pe_param, pe_buffer, st(pcs))
        alpha, color = vmap(fc_model)(fc_param, fc_buffer, emb, sh, tx)
        loss, ld, _ = ref_loss.step_batch_loss(alpha, color, st(gd), st(grgb) / 255.0, st(om), st(dm), st(zs))
        rs, rt = ref_loss.step_batch_loss_reg(cls_dict, torch.arange(K))
        loss = loss + 0.0005 * (rs + rt).sum()                                                                       # train.py:165-169
        loss.backward()
        opt.step(); opt.zero_grad(set_to_none=True)                                                                  # train.py:183-184
        traj[it, 0] = float(loss)
        traj[it, 1:1 + K] = npy(ld["depth"]); traj[it, 1 + K:1 + 2 * K] = npy(ld["color"]); traj[it, 1 + 2 * K:] = npy(ld["opacity"])
    out["curve/traj"] = traj


def golden_loss(out):
    """step_batch_loss on raw alpha/color incl. the empty-mask quirk (render_rays.py:67-72)."""
    g = torch.Generator().manual_seed(21)
    K, R, S = 3, 32, 14
    alpha = torch.randn(K, R, S, 1, generator=g) * 4
    color = torch.rand(K, R, S, 3, generator=g)
    z = torch.sort(torch.rand(K, R, S, generator=g) * 3 + 0.2, dim=-1).values
    gd = torch.rand(K, R, generator=g) * 3 + 0.3
    gc = torch.rand(K, R, 3, generator=g)
    lab = torch.randint(0, 3, (K, R), generator=g).to(torch.uint8)
    md = torch.rand(K, R, generator=g) > 0.2
    for name, mdk in (("normal", md), ("empty_depth_cat1", md.clone())):
        if name != "normal":
            mdk[1] = False
        a = alpha.clone().requires_grad_(True); c = color.clone().requires_grad_(True)
        loss, ld, lcol = ref_loss.step_batch_loss(a, c, gd, gc, lab, mdk, z)
        loss.backward()
        occ = ref_render.occupancy_activation(a.squeeze(-1))
        term = ref_render.occupancy_to_termination(occ, is_batch=True)
        p = "loss/" + name + "/"
        for kk, v in dict(alpha=alpha, color=color, z=z, gt_depth=gd, gt_color=gc, labels=lab, mask_depth=mdk,
                          loss=loss, depth=ld["depth"], colorl=ld["color"], opacity=ld["opacity"],
                          galpha=a.grad, gcolor=c.grad, term=term).items():
            out[p + kk] = npy(v)
    # 2-D (non-batch) termination used by category_registration.py:152
    occ2 = torch.rand(10, 96, generator=g)
    out["loss/term2d/occ"] = npy(occ2)
    out["loss/term2d/term"] = npy(ref_render.occupancy_to_termination(occ2))


def golden_pool(out):
    """sceneCategory.get_training_samples (scene_cateogries.py:352-451) on bare instances: the per-step pool slices, origin_dirs_O/W
    and the epoch reshuffle are the reference's own code; only the instance's sample_3d_points is replaced by a pass-through so
    that the sampler's INPUTS are recorded (the sampler itself is pinned by golden_sampler)."""
    def passthrough(rgbs, depth, origins, dirs):
        return rgbs, depth, origins, dirs, rgbs[:, :0], depth[:0]

    def record(tag, sc, n, steps, P, seed):
        torch.manual_seed(seed)
        for t in range(steps):
            rgbs, depth, origins, dirs, _, _, idx = sc.get_training_samples(n)
            for kk, v in dict(rgbs=rgbs, depth=depth, origins=origins, dirs=dirs, idx=idx).items():
                out["pool/%s/step%d/%s" % (tag, t, kk)] = npy(v)
        torch.manual_seed(seed)                     # the only RNG consumer was randperm(P) at each epoch end: replay it
        e = 0
        i_batch = 0
        for t in range(steps):
            i_batch += n
            if i_batch >= P - n:
                out["pool/%s/perm%d" % (tag, e)] = npy(torch.randperm(P)); e += 1
                i_batch = 0
        out["pool/%s/meta" % tag] = np.array([P, n, steps, e])

    g = torch.Generator().manual_seed(77)
    P, n_obj = 1000, 4
    pool = O.synth_pool(P, n_obj, 7, g)
    for kk, v in pool.items():
        out["pool/in/" + kk] = npy(v)
    for tag, ids in (("multi", [1, 2, 3, 4]), ("single", [9])):
        sc = bare_scene(1, 9)
        sc.cls_id = 3
        sc.obj_ids = ids
        sc.i_batch = 0
        sc.batch_indices_all = pool["indices"].clone()
        sc.rgbs_batch_all = pool["rgbs"].clone()
        sc.depth_batch_all = pool["depth"].clone()
        sc.ray_dirs_batch_all = pool["ray_dirs"].clone()
        sc.t_co_batch_all = pool["poses"].clone()
        sc.t_wc_batch_all = pool["poses"].clone()
        sc.sample_3d_points = passthrough
        record(tag, sc, 96, 25, P, 5)
    sc = bare_scene(5, 9)
    sc.cls_id = 0
    sc.obj_ids = [0]
    sc.trainer = types.SimpleNamespace(inst_id_to_index={0: 0})
    sc.rgbs_batch_dict = {0: pool["rgbs"].clone()}
    sc.depth_batch_dict = {0: pool["depth"].clone()}
    sc.frame_batch_dict = {0: pool["pose_index"].clone()}
    sc.ray_dirs_batch_dict = {0: pool["ray_dirs"].clone()}
    sc.t_wc_batch_dict = {0: pool["pose_table"].clone()}
    sc.i_batch_dict = {0: 0}
    sc.sample_3d_points = passthrough
    record("background", sc, 120, 20, P, 6)


def main():
    torch.set_num_threads(1)
    groups = dict(sampler=golden_sampler, embed=golden_embed, step=golden_train_step, bg=golden_bg,
                  loss=golden_loss, bg128=golden_bg128, curve=golden_curve, pool=golden_pool)
    only = sys.argv[1:]
    for gname, fn in groups.items():
        if only and gname not in only:
            continue
        out = {}
        fn(out)
        path = os.path.join(HERE, gname + ".npz")
        np.savez_compressed(path, **out)
        print(gname, len(out), "arrays ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

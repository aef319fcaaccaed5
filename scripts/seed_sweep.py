"""Distribution of the bf16-vs-fp32 gradient error of the fused step over weight / data seeds (development aid; backs the
tolerances stated in tests/test_gpu_fused.py): the well-conditioned fixture (depth term off), 12 seeds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cnf_oracle as O
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
from category_nerf_reconstruction_official_b200 import ops
DEV = torch.device("cuda:0")
K, n_obj, R, N, M, L = 2, 4, 1536, 1, 9, 32
names = [n for n, _ in O.codenerf_param_shapes(32, L)]
allw = []
for seed in range(12):
    g = torch.Generator().manual_seed(100 + seed)
    tr = FusedCategoryTrainer(K, n_obj, latent_dim=L, scale=2.0, device=DEV, n_bins_cam2surface=N, n_bins=M, generator=g)
    cats = [O.synth_rays(R, n_obj, g) for _ in range(K)]
    rgbs = torch.stack([c[0] for c in cats]); depth = torch.zeros(K, R)
    org = torch.stack([c[2] for c in cats]); dirs = torch.stack([c[3] for c in cats]); idx = torch.stack([c[4] for c in cats])
    rand = {"u_inv": torch.rand(K, R, N + M, generator=g), "u_val": torch.rand(K, R, N, generator=g),
            "n_obj": torch.randn(K, R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(K, R, M, generator=g)}
    dv = lambda t: t.to(DEV)
    z, _, valid, labels, stats = ops.sample_batched(dv(rgbs), dv(depth), dv(org), dv(dirs), N, M, 0.1, 0.05, {k: dv(v) for k, v in rand.items()},
                                                     max_depth=dv(torch.full((K,), 3.0)), want_pcs=False)
    loss = tr.step_from_samples(dv(org), dv(dirs), z, dv(depth), dv(rgbs), valid.view(torch.uint8), dv(idx).to(torch.int32), stats[:, 4:8])
    torch.cuda.synchronize()
    zc = z.cpu()
    pcs = org[:, :, None, :] + dirs[:, :, None, :] * zc[..., None]
    fc = [torch.stack([tr.param_views(k)[i].cpu() for k in range(K)]) for i in range(28)]
    out = O.train_step(fc, tr.pe_B.cpu(), 2.0, [tr.shape_codes[k].cpu() for k in range(K)], [tr.tex_codes[k].cpu() for k in range(K)],
                       pcs, zc, depth, rgbs[..., :3], labels.cpu(), valid.cpu().bool(), idx)
    worst = (0.0, "", 1.0)
    def chk(a, e, name):
        global worst
        a = a.detach().double().cpu().flatten(); e = e.detach().double().cpu().flatten()
        if float(e.norm()) < 1e-12: return
        rel = float((a - e).norm() / e.norm()); cos = float(torch.nn.functional.cosine_similarity(a, e, dim=0))
        if rel > worst[0]: worst = (rel, name, cos)
    for k in range(K):
        gv = tr.grad_views(k)
        for i in range(28): chk(gv[i], out["g_fc"][i][k], names[i])
        chk(tr.g_B[k], out["g_B"][k], "B_layer"); chk(tr.g_shape[k], out["g_shape"][k], "shape_codes"); chk(tr.g_tex[k], out["g_texture"][k], "texture_codes")
    # whole-gradient error (all tensors of all categories concatenated)
    A = torch.cat([tr.g_params.flatten().cpu().double()]); E = torch.cat([torch.cat([out["g_fc"][i][k].flatten() for i in range(28)]) for k in range(K)]).double()
    tot = float((A - E).norm() / E.norm())
    print("seed %d: loss rel %.1e  whole MLP gradient rel L2 %.4f  worst tensor %s rel %.4f cos %.4f" %
          (seed, abs(float(loss) - float(out["loss"])) / float(out["loss"]), tot, worst[1], worst[0], worst[2]))
    allw.append(worst[0])
print("worst-tensor rel L2 over seeds: median %.4f max %.4f" % (sorted(allw)[len(allw) // 2], max(allw)))

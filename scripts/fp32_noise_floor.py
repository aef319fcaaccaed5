"""How far is the reference's OWN fp32 arithmetic from exact?  The oracle (= the reference's math) is run in fp32 and in fp64 on the
well-conditioned parity fixture of tests/test_gpu_fused.py; the per-tensor relative L2 distance between the two is the noise floor any
reduced-precision path is measured against (CPU only; development aid backing the tolerances stated in DESIGN.md section 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cnf_oracle as O
K, n_obj, R, N, M, L = 2, 4, 1536, 1, 9, 32
names = [n for n, _ in O.codenerf_param_shapes(32, L)]
for seed in range(4):
    g = torch.Generator().manual_seed(100 + seed)
    shapes = O.codenerf_param_shapes(32, L)
    fc = [torch.stack([p for p in ps]) for ps in zip(*[O.init_params(shapes, g) for _ in range(K)])]
    B = O.icosa_matrix().repeat(K, 1, 1)
    sc = [O.init_codes(n_obj, L, g) for _ in range(K)]; tc = [O.init_codes(n_obj, L, g) for _ in range(K)]
    cats = [O.synth_rays(R, n_obj, g) for _ in range(K)]
    rgbs = torch.stack([c[0] for c in cats]); depth = torch.zeros(K, R)
    org = torch.stack([c[2] for c in cats]); dirs = torch.stack([c[3] for c in cats]); idx = torch.stack([c[4] for c in cats])
    z = torch.sort(torch.rand(K, R, N + M, generator=g) * 3.0, dim=-1).values
    pcs = org[:, :, None, :] + dirs[:, :, None, :] * z[..., None]
    labels = rgbs[..., 3]; valid = torch.zeros(K, R, dtype=torch.bool)
    outs = []
    for dt in (torch.float32, torch.float64):
        c = lambda t: t.to(dt)
        outs.append(O.train_step([c(p) for p in fc], c(B), 2.0, [c(s) for s in sc], [c(t) for t in tc], c(pcs), c(z), c(depth),
                                 rgbs[..., :3], labels, valid, idx))
    a, e = outs
    worst = (0.0, "")
    tot_n = tot_d = 0.0
    for i in range(28):
        for k in range(K):
            d = (a["g_fc"][i][k].double() - e["g_fc"][i][k]).norm(); n = e["g_fc"][i][k].norm()
            tot_n += float(d) ** 2; tot_d += float(n) ** 2
            if float(n) > 1e-12 and float(d / n) > worst[0]: worst = (float(d / n), names[i])
    dB = float((a["g_B"].double() - e["g_B"]).norm() / e["g_B"].norm())
    print("seed %d: fp32 vs fp64 oracle: loss rel %.1e, whole MLP gradient rel L2 %.2e, worst tensor %s %.2e, B_layer %.2e"
          % (seed, abs(float(a["loss"]) - float(e["loss"])) / float(e["loss"]), (tot_n / tot_d) ** 0.5, worst[1], worst[0], dB))

"""Repeat the same fused step many times and compare the gradients with the first run (development aid): anything beyond
fp32 atomic-order noise indicates a race in the kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
dev = torch.device("cuda:0")
for name, R, reps in (("replica_room0", None, 300), ("c4", 65536, 60), ("c1", None, 300)):
    wl = dict(WORKLOADS[name])
    if R: wl["R"] = R
    tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"],
                              generator=torch.Generator().manual_seed(5))
    b = synth_batch(wl, seed=3, device=dev)
    from category_nerf_reconstruction_official_b200 import sampling
    rand = sampling.draw_randoms(b["rgbs"], b["depth"], wl["N"], wl["M"], 0.1)
    ref = None; worst = 0.0; bad = 0
    for i in range(reps):
        tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"], rand=rand)
        g = torch.cat([tr.g_params.flatten(), tr.g_B.flatten(), tr.g_shape.flatten(), tr.g_tex.flatten(), tr.out_loss.flatten()]).clone()
        if ref is None:
            ref = g; continue
        d = float((g - ref).abs().max() / ref.abs().max())
        worst = max(worst, d); bad += d > 1e-4
    print("%s: %d repeats, worst max-abs deviation / max |grad| = %.3e, repeats beyond 1e-4: %d, finite: %s" % (name, reps, worst, bad, bool(torch.isfinite(ref).all())))

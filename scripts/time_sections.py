"""Warm (L2-resident) device time of the sections of one training step (development aid): python scripts/time_sections.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200 import ops, sampling
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
name = sys.argv[1] if len(sys.argv) > 1 else "replica_room0"
wl = dict(WORKLOADS[name])
dev = torch.device("cuda:0")
tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"])
b = synth_batch(wl, seed=3, device=dev)
rgbs, depth, org, dirs, idx = b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"].contiguous()
R = rgbs.shape[1]
def ev(): return torch.cuda.Event(enable_timing=True)
acc = {}
for it in range(30):
    e = [ev() for _ in range(6)]
    e[0].record()
    rand = sampling.draw_randoms(rgbs, depth, tr.N, tr.M, tr.eps, tr.min_bound, tr.this_obj)
    e[1].record()
    z, _, valid, labels, stats = ops.sample_batched(rgbs, depth, org, dirs, tr.N, tr.M, tr.eps, tr.stop_eps, rand, tr.min_bound, tr.this_obj, want_pcs=False)
    e[2].record()
    tr.prepare(idx, R)
    e[3].record()
    tr.step_from_samples(org, dirs, z, depth, rgbs, valid.view(torch.uint8), idx, stats[:, 4:8], True, prepared=True)
    e[4].record()
    torch.cuda.synchronize()
    if it >= 10:
        for i, n in enumerate(["rng (4 torch kernels)", "sampler (3 kernels)", "prepare (memset+prep+sort)", "fused + finalize"]):
            acc[n] = acc.get(n, 0.0) + e[i].elapsed_time(e[i + 1]) / 20
for n, v in acc.items():
    print("%-30s %7.1f us" % (n, v * 1e3))
print("(eager launches from Python: sections include launch gaps)")

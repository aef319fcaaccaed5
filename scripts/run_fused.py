"""Run a few fused training steps of one workload (profiling target): python scripts/run_fused.py [workload] [R] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
wl = dict(WORKLOADS[name])
if len(sys.argv) > 2:
    wl["R"] = int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"])
b = synth_batch(wl, seed=3, device=dev)
for i in range(steps):
    loss = tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
torch.cuda.synchronize()
print("loss", float(loss))

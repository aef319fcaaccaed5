"""Small driver for ncu: a few steps of the fused path at a mid-size batch (K categories x R rays)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="replica_room0")
ap.add_argument("--rays", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
wl = dict(WORKLOADS[a.workload])
if a.rays:
    wl["R"] = a.rays
dev = torch.device("cuda:0")
tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev,
                          n_bins_cam2surface=wl["N"], n_bins=wl["M"], generator=torch.Generator().manual_seed(1))
b = synth_batch(wl, seed=3, device=dev)
for _ in range(a.steps):
    loss = tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
torch.cuda.synchronize()
print("loss", float(loss))

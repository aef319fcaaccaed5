"""One background step (FusedBackgroundTrainer, eager) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200.background import FusedBackgroundTrainer
from oracle import cnf_oracle as O
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
tr = FusedBackgroundTrainer(scale=5.0, device=dev, generator=g)
rgbs, depth, org, dirs, _ = [t.to(dev) for t in O.synth_rays(1200, 1, g)]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    tr.step(rgbs, depth, org, dirs)
torch.cuda.synchronize()
print("loss", float(tr.out_loss[0]))

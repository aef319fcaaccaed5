import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200 import _lib
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
wl = dict(WORKLOADS["c4"]); wl["R"] = 65536
dev = torch.device("cuda:0")
tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"])
b = synth_batch(wl, seed=3, device=dev)
dbg = torch.zeros(128, dtype=torch.int64, device=dev)
lib = _lib.load()
for i in range(3):
    tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
_lib.check(lib.cnf_set_debug_buffer(dbg.data_ptr()))
tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
torch.cuda.synchronize()
_lib.check(lib.cnf_set_debug_buffer(None))
d = dbg.cpu()
names = ["start", "pe+oh done(publish)"]
for st in ["L1","L2","L3","L4","L6","L7","L8","B8","B7","B6","B4","B3","B2"]:
    names += [st+" wait-begin", st+" mma-done"]
    if st == "L8":
        names += ["  L8 h3 stored", "  L9 colours done", "  compositing done"]
    names += [st+" epi-done(publish)"]
names += ["B1 wait-begin", "B1 mma-done", "pe-bwd done"]
for h in range(1):
    t = d[h*64:(h+1)*64].tolist()
    n = max(i for i,v in enumerate(t) if v) + 1
    print("half", h, "stamps", n, "tile cycles", t[n-1]-t[0])
    for i in range(1, n):
        print("   %-28s +%6d" % (names[i] if i < len(names) else str(i), t[i]-t[i-1]))

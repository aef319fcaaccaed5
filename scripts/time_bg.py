"""Background OccupancyMap (W=128) training step through the fp32 module path: 1200 rays x (5+9) samples (SURVEY C2 background)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import model, embedding, loss, ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
R, S = 1200, 14
m = model.OccupancyMap(87, 42, hidden_size=128).to(dev); m.use_tensor_cores = len(sys.argv) > 1
pe = embedding.UniDirsEmbed(max_deg=5, scale=5.0).to(dev) if hasattr(embedding, "UniDirsEmbed") else None
pcs = (torch.rand(R, S, 3, device=dev) - 0.5) * 6
z = torch.sort(torch.rand(R, S, device=dev) * 5, dim=-1).values
gt_d = torch.rand(R, device=dev) * 4 + 0.5; gt_c = torch.rand(R, 3, device=dev)
labels = torch.ones(R, dtype=torch.uint8, device=dev); md = torch.ones(R, dtype=torch.bool, device=dev)
params = list(m.parameters()) + list(pe.parameters())
def step():
    for p in params: p.grad = None
    emb = pe(pcs)
    a, c = m(emb)
    l, _, _ = loss.step_batch_loss(a[None], c[None], gt_d[None], gt_c[None], labels[None], md[None], z[None])
    l.backward()
    return l
for _ in range(5): step()
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("bg step eager (module path, fp32): median %.3f ms  min %.3f ms" % (sorted(ts)[10], min(ts)))
# kernel-only: time the two MLP launches
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))

# fused background trainer (C ABI only), eager and CUDA-graph replay
from category_nerf_reconstruction_official_b200.background import FusedBackgroundTrainer
from oracle import cnf_oracle as O
g = torch.Generator().manual_seed(1)
tr = FusedBackgroundTrainer(scale=5.0, device=dev, generator=g)
rgbs, depth, org, dirs, _ = [t.to(dev) for t in O.synth_rays(1200, 1, g)]
for _ in range(3): tr.step(rgbs, depth, org, dirs)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); tr.step(rgbs, depth, org, dirs); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("FusedBackgroundTrainer.step eager: median %.3f ms" % sorted(ts)[10])
replay = tr.capture(rgbs, depth, org, dirs)
for _ in range(3): replay()
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("FusedBackgroundTrainer graph replay: median %.3f ms (%.1f M rays/s)" % (sorted(ts)[10], 1200 / sorted(ts)[10] * 1e-3))

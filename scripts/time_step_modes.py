"""C2 step time as CUDA-graph replays: with / without the L2 flush between replays, side stream on / off (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
wl = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "replica_room0"])
dev = torch.device("cuda:0")
tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"])
b = synth_batch(wl, seed=3, device=dev)
replay = tr.capture(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5): replay()
torch.cuda.synchronize()
ts = []
for _ in range(30):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print("graph replay, L2 flushed before each : %.1f us (median of 30)" % (ts[15] * 1e3))
ts = []
for _ in range(30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print("graph replay, warm L2, one at a time  : %.1f us" % (ts[15] * 1e3))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): replay()
e1.record(); torch.cuda.synchronize()
print("200 back-to-back replays               : %.1f us per step" % (e0.elapsed_time(e1) / 200 * 1e3))

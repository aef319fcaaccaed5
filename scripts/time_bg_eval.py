"""Throughput of the forward-only tensor-core evaluation of the background OccupancyMap (cnf_occmap_tc_eval_points)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import ops, model
dev = torch.device('cuda:0')
tag = sys.argv[1] if len(sys.argv) > 1 else ""
torch.manual_seed(0)
m = model.OccupancyMap(87, 42, hidden_size=128).to(dev); m.apply(model.init_weights)
flat = model.flat_params(m).detach()
B = torch.randn(21, 3, device=dev); sc = torch.tensor(5.0, device=dev)
for P in (1 << 20, 1 << 24):
    pts = (torch.rand(P, 3, device=dev) - 0.5) * 8
    for _ in range(2): ops.occmap_eval_points_fused(pts, flat, B, sc)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.occmap_eval_points_fused(pts, flat, B, sc); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print("[%s] OccupancyMap(128) eval_points P=%d: %.3f ms, %.1f M points/s, %.0f TFLOP/s algorithmic (187392 FLOP/pt) = %.1f%% of bf16 peak"
          % (tag, P, ms, P / ms / 1e3, P * 187392 / ms / 1e9, P * 187392 / ms / 1e9 / 1656.5 * 100))

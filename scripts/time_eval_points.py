"""Throughput of Trainer.eval_points-style inference: tensor-core kernel (cnf_eval_points) vs the fp32 CUDA-core module path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import ops, model, embedding
dev = torch.device("cuda:0")
L = 256
m = model.CodeNeRF(87, 42, W=32, latent_dim=L).to(dev); m.apply(model.init_weights)
pe = embedding.UniDirsEmbed(max_deg=5, scale=2.0).to(dev)
sc = torch.randn(L, device=dev) / (L / 2) ** 0.5; tc = torch.randn(L, device=dev) / (L / 2) ** 0.5
flat = model.flat_params(m).detach()
for P in (1 << 20, 1 << 24):
    pts = (torch.rand(P, 3, device=dev) - 0.5) * 3.0
    for _ in range(2): ops.eval_points_fused(pts, flat, pe.B_layer.weight, pe.scale, sc, tc)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.eval_points_fused(pts, flat, pe.B_layer.weight, pe.scale, sc, tc); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print("tensor-core eval_points  P=%9d: %8.3f ms  %7.1f M points/s  %6.1f TFLOP/s algorithmic (27296 FLOP/pt) = %.1f%% of bf16 peak"
          % (P, ms, P / ms / 1e3, P * 27296 / ms / 1e9, P * 27296 / ms / 1e9 / 1656.5 * 100))
P = 1 << 20
pts = (torch.rand(P, 3, device=dev) - 0.5) * 3.0
with torch.no_grad():
    for _ in range(2): a, c = m(pe(pts), sc, tc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a, c = m(pe(pts), sc, tc); e1.record(); torch.cuda.synchronize()
print("fp32 module path         P=%9d: %8.3f ms  %7.1f M points/s" % (P, e0.elapsed_time(e1), P / e0.elapsed_time(e1) / 1e3))

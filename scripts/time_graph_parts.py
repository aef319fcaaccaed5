"""Graph-replay time of parts of the C2 training step (development aid): which launches are on the critical path?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200 import ops, sampling
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
wl = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "replica_room0"])
dev = torch.device("cuda:0")
tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"])
b = synth_batch(wl, seed=3, device=dev)
rgbs, depth, org, dirs, idx = b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"].contiguous()
R = rgbs.shape[1]
tr.operands_follow_optimizer = True          # as bench.py: the operand half of prepare runs with the optimizer step
tr.refresh_operands(R)
rand0 = sampling.draw_randoms(rgbs, depth, tr.N, tr.M, tr.eps)
z0, _, valid0, _, stats0 = ops.sample_batched(rgbs, depth, org, dirs, tr.N, tr.M, tr.eps, tr.stop_eps, rand0, tr.min_bound, tr.this_obj, want_pcs=False)
vm0 = valid0.view(torch.uint8)
tr.prepare(idx, R)
def full(): tr.step(rgbs, depth, org, dirs, idx)
def rng(): sampling.draw_randoms(rgbs, depth, tr.N, tr.M, tr.eps)
def sampler(): ops.sample_batched(rgbs, depth, org, dirs, tr.N, tr.M, tr.eps, tr.stop_eps, rand0, tr.min_bound, tr.this_obj, want_pcs=False)
def prepare(): tr.prepare(idx, R)
def prepare_full(): tr.prepare(idx, R, what=3)
def fused_fin(): tr.step_from_samples(org, dirs, z0, depth, rgbs, vm0, idx, stats0[:, 4:8], True, prepared=True)
def all_inline():
    tr.prepare(idx, R)
    rand = sampling.draw_randoms(rgbs, depth, tr.N, tr.M, tr.eps)
    z, _, valid, _, stats = ops.sample_batched(rgbs, depth, org, dirs, tr.N, tr.M, tr.eps, tr.stop_eps, rand, tr.min_bound, tr.this_obj, want_pcs=False)
    tr.step_from_samples(org, dirs, z, depth, rgbs, valid.view(torch.uint8), idx, stats[:, 4:8], True, prepared=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn in (("full step (side-stream prepare)", full), ("full step, one stream", all_inline), ("rng only", rng), ("sampler only", sampler),
                 ("prepare (input half) only", prepare), ("prepare (operands + inputs)", prepare_full), ("fused + finalize only", fused_fin)):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    for _ in range(5): g.replay()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("%-36s %7.1f us (median of 20, L2 flushed)" % (name, ts[10] * 1e3))

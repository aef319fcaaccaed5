"""Kernel-only timing of the fused step (CUDA events around fused_step_kernel): python scripts/time_fused.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch, FLOP_PER_POINT
from category_nerf_reconstruction_official_b200 import _lib
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
dev = torch.device("cuda:0")
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = [sys.argv[1] if len(sys.argv) > 1 else "-"]
for name, R in (("c4", 262144), ("replica_room0", None), ("c1", None)):
    wl = dict(WORKLOADS[name])
    if R: wl["R"] = R
    tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"])
    b = synth_batch(wl, seed=3, device=dev)
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e1.record(); torch.cuda.synchronize()
        flush.zero_()
        _lib.check(lib.cnf_set_kernel_events(e0.cuda_event, e1.cuda_event))
        tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[3:]); ms = ts[len(ts) // 2]
    pts = wl["K"] * wl["R"] * (wl["N"] + wl["M"])
    out.append("%s: %.4f ms (%.1f TF/s, %.2f%%)" % (name, ms, pts * FLOP_PER_POINT / ms / 1e9, pts * FLOP_PER_POINT / ms / 1e9 / 1656.5 * 100))
print(" | ".join(out))

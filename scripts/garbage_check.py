"""Uninitialised-read detector (development aid): run the same seeded fused step with the allocator pool pre-filled with
different byte patterns (and after unrelated kernels dirtied shared memory); gradients must agree to atomic-order noise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
from category_nerf_reconstruction_official_b200 import sampling, ops
dev = torch.device("cuda:0")

def dirty(pattern):
    # fill ~2 GB of the caching allocator's pool with a byte pattern, then release it to the pool (not to the driver)
    bufs = [torch.empty(256 << 20, dtype=torch.uint8, device=dev) for _ in range(8)]
    for b in bufs:
        if pattern is None:
            b.random_(0, 256)
        else:
            b.fill_(pattern)
    torch.cuda.synchronize()
    del bufs

def one(name, R, pattern, dirty_smem):
    wl = dict(WORKLOADS[name])
    if R: wl["R"] = R
    dirty(pattern); torch.manual_seed(7)
    tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"],
                              generator=torch.Generator().manual_seed(5))
    b = synth_batch(wl, seed=3, device=dev)
    g = torch.Generator(device=dev).manual_seed(9)
    K, Rr, N, M = wl["K"], wl["R"], wl["N"], wl["M"]
    rand = {"u_inv": torch.rand(K, Rr, N + M, generator=g, device=dev), "u_val": torch.rand(K, Rr, N, generator=g, device=dev),
            "n_obj": torch.randn(K, Rr, M, generator=g, device=dev) * (0.1 / 3), "u_oth": torch.rand(K, Rr, M, generator=g, device=dev)}
    if dirty_smem:
        pts = torch.rand(200000, 3, device=dev)
        ops.eval_points_fused(pts, tr.params[0], tr.pe_B[0], tr.pe_scale[0], tr.shape_codes[0, 0], tr.tex_codes[0, 0])
    tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"], rand=rand)
    torch.cuda.synchronize()
    out = torch.cat([tr.g_params.flatten(), tr.g_B.flatten(), tr.g_shape.flatten(), tr.g_tex.flatten(), tr.out_loss.flatten()]).clone()
    del tr, b, rand
    return out

for name, R in (("c1", None), ("replica_room0", None), ("c4", 65536)):
    ref = one(name, R, 0, False)
    for pattern, ds in ((0, False), (0xff, False), (0x7f, False), (None, False), (0, True), (None, True)):
        o = one(name, R, pattern, ds)
        d = float((o - ref).abs().max() / ref.abs().max())
        print("%s pattern %s dirty_smem %s: max-abs deviation / max |grad| = %.3e finite %s" % (name, pattern, ds, d, bool(torch.isfinite(o).all())))

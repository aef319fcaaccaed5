"""HBM roofline of the standalone kernels (sampler, embedding fwd/bwd, compositing+loss fwd/bwd) on one B200.
Algorithmic bytes per unit from SURVEY 8(d); peak = MEASURED_PEAKS.json hbm_gbs.  Prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import peaks
from category_nerf_reconstruction_official_b200 import ops, sampling
from category_nerf_reconstruction_official_b200.synthetic import synth_rays

dev = torch.device("cuda:0")
quick = "--quick" in sys.argv
R = 65536 if quick else 1 << 20
N, M = 1, 9
S = N + M
g = torch.Generator().manual_seed(0)
base = synth_rays(16384, 8, g)
rep = R // 16384
rgbs, depth, org, dirs, idx = (t.repeat(*([rep] + [1] * (t.dim() - 1)))[None].to(dev) for t in base)
rand = sampling.draw_randoms(rgbs, depth, N, M, 0.1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5 if not quick else 2):
    """Device time of one call, as a CUDA-graph replay (the ops are a few launches of ~100 us: eager launches from Python
    would measure the host, not the kernels); L2 flushed before every timed replay."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e-3


pk = peaks()
out = {"peak_hbm_gbs": pk["hbm"], "peak_source": pk["source"], "rays": R, "S": S}
# sampler (3 launches: count, scan, emit)
t = timeit(lambda: ops.sample_batched(rgbs, depth, org, dirs, N, M, 0.1, 0.05, rand))
b = R * (32 + 20 * S)
out["sampler"] = {"ms": t * 1e3, "alg_bytes": b, "gbs": b / t / 1e9, "frac": b / t / 1e9 / pk["hbm"]}
z, pcs, valid, labels, stats = ops.sample_batched(rgbs, depth, org, dirs, N, M, 0.1, 0.05, rand)
# embedding
P = R * S
x = pcs.reshape(1, P, 3)
Bm = torch.randn(1, 21, 3, device=dev); sc = torch.tensor([2.0], device=dev)
t = timeit(lambda: ops.EmbedFn.apply(x, Bm, sc))
b = P * 528
out["embed_fwd"] = {"ms": t * 1e3, "alg_bytes": b, "gbs": b / t / 1e9, "frac": b / t / 1e9 / pk["hbm"]}
emb = ops.EmbedFn.apply(x, Bm, sc)
gout = torch.randn_like(emb)
lib = ops.lib_for(emb)
gB = torch.zeros_like(Bm)
t = timeit(lambda: ops.check(lib.cnf_embed_bwd(1, P, ops.ptr(x), ops.ptr(Bm), ops.ptr(sc), ops.ptr(gout), ops.ptr(gB), None, ops.stream())))
b = P * (516 + 12)
out["embed_bwd"] = {"ms": t * 1e3, "alg_bytes": b, "gbs": b / t / 1e9, "frac": b / t / 1e9 / pk["hbm"]}
del emb, gout
# compositing + loss
alpha = torch.randn(1, R, S, device=dev) * 3
color = torch.rand(1, R, S, 3, device=dev)
gd = depth.clone(); gc = torch.rand(1, R, 3, device=dev)
md = valid.view(torch.uint8)
counts = ops.mask_counts(labels, md)
ws = torch.empty(lib.cnf_composite_workspace_bytes(1, R), dtype=torch.uint8, device=dev)
rays = torch.empty(1, R, 6, device=dev); ol = torch.empty(7, device=dev)
t = timeit(lambda: ops.check(lib.cnf_composite_loss_fwd(1, R, S, ops.ptr(alpha), ops.ptr(color), ops.ptr(z), ops.ptr(gd), ops.ptr(gc),
                                                        ops.ptr(labels), ops.ptr(md), ops.ptr(counts), 5.0, 10.0, ops.ptr(rays), ops.ptr(ol),
                                                        ops.ptr(ws), ws.numel(), ops.stream())))
b = P * 20 + R * (22 + 24)
out["composite_fwd"] = {"ms": t * 1e3, "alg_bytes": b, "gbs": b / t / 1e9, "frac": b / t / 1e9 / pk["hbm"]}
ga = torch.empty_like(alpha); gcol = torch.empty_like(color)
t = timeit(lambda: ops.check(lib.cnf_composite_loss_bwd(1, R, S, ops.ptr(alpha), ops.ptr(color), ops.ptr(z), ops.ptr(gd), ops.ptr(gc),
                                                        ops.ptr(labels), ops.ptr(md), ops.ptr(counts), 5.0, 10.0, None, ops.ptr(ga),
                                                        ops.ptr(gcol), ops.stream())))
b = P * 36 + R * 22
out["composite_bwd"] = {"ms": t * 1e3, "alg_bytes": b, "gbs": b / t / 1e9, "frac": b / t / 1e9 / pk["hbm"]}
print(json.dumps(out))

"""HBM roofline of the standalone kernels (sampler, embedding fwd/bwd, compositing+loss fwd/bwd) on one B200.
Algorithmic bytes per unit from SURVEY 8(d); peak = MEASURED_PEAKS.json hbm_gbs.  ``measure()`` is also called by bench.py
(``hbm_kernels`` block of the bench line); as a script it prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def measure(dev, R=1 << 20, iters=5, flush=None):
    from bench import peaks
    from category_nerf_reconstruction_official_b200 import ops, sampling
    from category_nerf_reconstruction_official_b200.synthetic import synth_rays
    N, M = 1, 9
    S = N + M
    g = torch.Generator().manual_seed(0)
    base = synth_rays(16384, 8, g)
    rep = R // 16384
    rgbs, depth, org, dirs, idx = (t.repeat(*([rep] + [1] * (t.dim() - 1)))[None].to(dev) for t in base)
    rand = sampling.draw_randoms(rgbs, depth, N, M, 0.1)
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    REPS = 8

    def timeit(fn):
        """Device time of one call: a CUDA graph of REPS back-to-back calls, replay time / REPS (the ops are one or two launches of
        40-100 us, so a single-call replay is 10-15 % graph-launch latency, and eager launches from Python would measure the host).
        L2 is flushed before every timed replay and every op's working set (>= 240 MB at 1 Mi rays) exceeds the 126 MB L2, so
        repeated calls still stream from HBM."""
        if os.environ.get("CNF_PROFILE_ONCE"):        # under ncu: two plain launches per op, no graph, no timing
            fn(); fn(); torch.cuda.synchronize()
            return 1.0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(REPS):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / iters / REPS * 1e-3

    pk = peaks()
    out = {"peak_hbm_gbs": pk["hbm"], "peak_source": pk["source"], "rays": R, "S": S,
           "timing": "CUDA graph of %d back-to-back calls per timed replay, L2 flushed before each replay, working sets > L2" % REPS}
    entry = lambda t, b: {"ms": t * 1e3, "alg_bytes": b, "gbs": b / t / 1e9, "frac": b / t / 1e9 / pk["hbm"]}
    # sampler (count, scan, emit)
    t = timeit(lambda: ops.sample_batched(rgbs, depth, org, dirs, N, M, 0.1, 0.05, rand))
    out["sampler"] = entry(t, R * (32 + 20 * S))
    z, pcs, valid, labels, stats = ops.sample_batched(rgbs, depth, org, dirs, N, M, 0.1, 0.05, rand)
    # embedding
    P = R * S
    x = pcs.reshape(1, P, 3)
    Bm = torch.randn(1, 21, 3, device=dev); sc = torch.tensor([2.0], device=dev)
    t = timeit(lambda: ops.EmbedFn.apply(x, Bm, sc))
    out["embed_fwd"] = entry(t, P * 528)
    emb = ops.EmbedFn.apply(x, Bm, sc)
    gout = torch.randn_like(emb)
    lib = ops.lib_for(emb)
    gB = torch.zeros_like(Bm)
    t = timeit(lambda: ops.check(lib.cnf_embed_bwd(1, P, ops.ptr(x), ops.ptr(Bm), ops.ptr(sc), ops.ptr(gout), ops.ptr(gB), None, ops.stream())))
    out["embed_bwd"] = entry(t, P * (516 + 12))
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
    out["composite_fwd"] = entry(t, P * 20 + R * (22 + 24))
    ga = torch.empty_like(alpha); gcol = torch.empty_like(color)
    t = timeit(lambda: ops.check(lib.cnf_composite_loss_bwd(1, R, S, ops.ptr(alpha), ops.ptr(color), ops.ptr(z), ops.ptr(gd), ops.ptr(gc),
                                                            ops.ptr(labels), ops.ptr(md), ops.ptr(counts), 5.0, 10.0, None, ops.ptr(ga),
                                                            ops.ptr(gcol), ops.stream())))
    out["composite_bwd"] = entry(t, P * 36 + R * 22)
    return out


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    print(json.dumps(measure(torch.device("cuda:0"), 65536 if quick else 1 << 20, 2 if quick else 5)))

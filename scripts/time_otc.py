"""Device time of the tensor-core OccupancyMap forward / backward calls, CUDA-graph replays (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import model, _lib
from category_nerf_reconstruction_official_b200._lib import check
dev = torch.device("cuda:0"); torch.manual_seed(0)
lib = _lib.load()
tag = sys.argv[1] if len(sys.argv) > 1 else ""
m = model.OccupancyMap(87, 42, hidden_size=128).to(dev); m.apply(model.init_weights)
flat = model.flat_params(m).detach().contiguous()
for P in (16800, 148 * 128 * 4):
    emb = torch.randn(P, 129, device=dev)
    ws = torch.empty(lib.cnf_occmap_tc_workspace_bytes(P), dtype=torch.uint8, device=dev)
    alpha = torch.empty(P, device=dev); rgb = torch.empty(P, 3, device=dev)
    ga = torch.randn(P, device=dev); gc = torch.randn(P, 3, device=dev); gp = torch.empty_like(flat); ge = torch.empty(P, 129, device=dev)
    s = torch.cuda.Stream()
    def fwd(): check(lib.cnf_occmap_tc_fwd(P, flat.data_ptr(), emb.data_ptr(), alpha.data_ptr(), rgb.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    def bwd(): check(lib.cnf_occmap_tc_bwd(P, flat.data_ptr(), rgb.data_ptr(), ga.data_ptr(), gc.data_ptr(), ws.data_ptr(), ws.numel(), gp.data_ptr(), ge.data_ptr(), torch.cuda.current_stream().cuda_stream))
    pts = torch.rand(P, 3, device=dev) * 4 - 2; Bm = torch.randn(21, 3, device=dev); sc = torch.tensor([5.0], device=dev); gB = torch.empty(21, 3, device=dev)
    def fwdp(): check(lib.cnf_occmap_tc_fwd_points(P, flat.data_ptr(), Bm.data_ptr(), sc.data_ptr(), pts.data_ptr(), alpha.data_ptr(), rgb.data_ptr(), ws.data_ptr(), ws.numel(), 0, torch.cuda.current_stream().cuda_stream))
    def bwdp(): check(lib.cnf_occmap_tc_bwd_points(P, flat.data_ptr(), pts.data_ptr(), rgb.data_ptr(), ga.data_ptr(), gc.data_ptr(), ws.data_ptr(), ws.numel(), gp.data_ptr(), gB.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
    res = []
    for name, fn in (("fwd", fwd), ("bwd", bwd), ("fwd+bwd", lambda: (fwd(), bwd())), ("fwd_points", fwdp), ("bwd_points", bwdp), ("points fwd+bwd", lambda: (fwdp(), bwdp()))):
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): fn()
        for _ in range(2): g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 10)
        res.append("%s %.1f us" % (name, sorted(ts)[2] * 1e3))
    print("[%s] P=%d: %s" % (tag, P, "  ".join(res)))

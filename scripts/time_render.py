"""Throughput of cnf_render_rays on BASELINE config 5 (dense 512x512 views of one instance, forward only): S = 14 and S = 96."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import ops, _lib
from oracle import cnf_oracle as O   # parameter initialisation only (development script)

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
L = 256
flat = torch.cat([p.reshape(-1) for p in O.init_params(O.codenerf_param_shapes(32, L), g)]).to(dev)
B = O.icosa_matrix().to(dev); scale = torch.tensor(2.0, device=dev)
sc = O.init_codes(1, L, g)[0].to(dev); tc = O.init_codes(1, L, g)[0].to(dev)
res = {}
for S in (14, 96):
    R = 512 * 512
    org = ((torch.rand(R, 3, generator=g) - 0.5) * 2).to(dev)
    dirs = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    z = torch.sort(torch.rand(R, S, generator=g) * 2.5, dim=-1).values.to(dev)
    lib = _lib.load()
    ws = torch.empty(lib.cnf_render_rays_workspace_bytes(R, S), device=dev, dtype=torch.uint8)
    out = torch.empty(R, 6, device=dev)
    flush = torch.empty(64 << 20, device=dev, dtype=torch.float32)
    def run():
        _lib.check(lib.cnf_render_rays(R, S, L, flat.data_ptr(), B.data_ptr(), scale.data_ptr(), sc.data_ptr(), tc.data_ptr(),
                                       org.data_ptr(), dirs.data_ptr(), z.data_ptr(), out.data_ptr(), None, ws.data_ptr(), ws.numel(),
                                       torch.cuda.current_stream().cuda_stream))
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    res["S%d" % S] = {"rays": R, "S": S, "ms": ms, "rays_per_s": R / ms * 1e3, "points_per_s": R * S / ms * 1e3,
                      "tensor_tflops": R * S * 27296 / ms * 1e-9}
    print("S=%d: %.3f ms per 512x512 view, %.1f M rays/s, %.2f G points/s, %.0f TFLOP/s" %
          (S, ms, R / ms * 1e-3, R * S / ms * 1e-6, R * S * 27296 / ms * 1e-9))
json.dump(res, open("gpurun_out/render.json", "w"), indent=1)

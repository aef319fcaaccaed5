import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import dev as _dev; _dev.use_dev_library()   # include/cnf_dev.h hooks live in the development build
from category_nerf_reconstruction_official_b200 import _lib
from category_nerf_reconstruction_official_b200._lib import check, ptr, stream
dev = torch.device("cuda:0")
lib = _lib.lib_for(torch.empty(1, device=dev))
N, Kd = 32, 32
g = torch.Generator().manual_seed(0)
A = torch.randn(64, Kd, generator=g).bfloat16().float(); B = torch.randn(N, Kd, generator=g).bfloat16().float()
ref = A @ B.t()
for lane_off in (0, 16, 32, 64):
    D = torch.zeros(128, N, device=dev)
    try:
        check(lib.cnf_umma_m64_probe(N, Kd, lane_off, ptr(A.to(dev)), ptr(B.to(dev)), ptr(D), stream()))
        torch.cuda.synchronize()
    except Exception as e:
        print("lane_off", lane_off, "ERROR", e); break
    D = D.cpu()
    touched = [l for l in range(128) if not torch.all(D[l] == -777.0)]
    # which reference row does each touched lane hold?
    mapping = {}
    for l in touched:
        d = (ref - D[l][None]).abs().max(dim=1).values
        r = int(d.argmin())
        mapping[l] = (r, float(d.min()))
    print("lane_off", lane_off, "touched lanes:", touched[:8], "...", touched[-8:], "count", len(touched))
    print("   lane->row:", [(l, mapping[l][0]) for l in touched[:6]], [(l, mapping[l][0]) for l in touched[30:36]], "max err", max(v[1] for v in mapping.values()) if mapping else None)

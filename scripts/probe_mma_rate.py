"""tcgen05.mma issue-rate probe (development aid): cycles per M x N x 16 bf16 MMA by operand source / major-ness."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import dev as _dev; _dev.use_dev_library()   # include/cnf_dev.h hooks live in the development build
from category_nerf_reconstruction_official_b200 import _lib
from category_nerf_reconstruction_official_b200._lib import check, ptr, stream
dev = torch.device("cuda:0")
lib = _lib.lib_for(torch.empty(1, device=dev))
names = {0: "SS K-major (fwd)", 1: "TS A in TMEM", 2: "SS MN/MN (wgrad)", 3: "SS B MN-major (dgrad)"}
reps = 1024
for n_cta in (148,):
    print("== CTAs:", n_cta)
    for mode in (0, 1, 2, 3):
        for M in (128, 64):
            for N in (16, 32, 48, 64, 80, 96, 128):
                for ndst in (2,):
                    if ndst * N > 256:
                        continue
                    out = torch.zeros(2 * n_cta, dtype=torch.int64, device=dev)
                    check(lib.cnf_umma_rate_probe(mode, M, N, reps, ndst, n_cta, ptr(out), stream()))
                    torch.cuda.synchronize()
                    o = out.cpu().view(n_cta, 2).double()
                    print("%-24s M=%3d N=%3d ndst=%d  issue %.1f  total %.1f cyc/mma" %
                          (names[mode], M, N, ndst, o[:, 0].mean() / reps, o[:, 1].mean() / reps))

"""Per-job wall time of fused_finalize_kernel (development aid): %globaltimer stamps of every CTA (category, job)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import dev as _dev; _dev.use_dev_library()   # include/cnf_dev.h hooks live in the development build
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200 import _lib
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
name = sys.argv[1] if len(sys.argv) > 1 else "replica_room0"
wl = dict(WORKLOADS[name]); dev = torch.device("cuda:0")
tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"])
b = synth_batch(wl, seed=3, device=dev)
for _ in range(3): tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
torch.cuda.synchronize()
dbg = torch.zeros(32768 + 2 * 4096, dtype=torch.int64, device=dev)
lib = _lib.load(); _lib.check(lib.cnf_set_debug_buffer(dbg.data_ptr()))
tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"]); torch.cuda.synchronize()
_lib.check(lib.cnf_set_debug_buffer(None))
K, J = wl["K"], 28
t = dbg[16384:16384 + 2 * K * J].view(K, J, 2).cpu()
t0 = int(t[..., 0].min())
dur = (t[..., 1] - t[..., 0]).float() / 1e3
end = (t[..., 1] - t0).float() / 1e3
start = (t[..., 0] - t0).float() / 1e3
names = ["0 plain"] + ["1 vd p%d" % i for i in range(3)] + ["2 enc_shape"] + ["3 cond s0"] + ["3 cat p%d" % i for i in range(3)] + ["3 cond s2", "3 cond s3"] + \
        ["4 lat s%d p%d" % (s, p) for s in range(4) for p in range(3)] + ["5 codes p%d" % i for i in range(4)] + ["6 reg+loss"]
print("finalize: first CTA start -> last CTA end %.1f us" % float(end.max()))
for j in range(J):
    print("  job %-12s start %5.1f..%5.1f us  duration mean %5.1f max %5.1f us  end max %5.1f" %
          (names[j], float(start[:, j].min()), float(start[:, j].max()), float(dur[:, j].mean()), float(dur[:, j].max()), float(end[:, j].max())))

# prepare_small_kernel: parts 0..7 weight images (+ constants in part 0), 8..11 latent slots, 12 instance sort, 13.. accumulator clear
JP = 8 + 4 + 1 + 2
tp = dbg[32768:32768 + 2 * K * JP].view(K, JP, 2).cpu()
if int(tp.max()) > 0:
    t0 = int(tp[..., 0].min())
    pn = ["img p%d" % i for i in range(8)] + ["latent s%d" % i for i in range(4)] + ["sort"] + ["clear p%d" % i for i in range(2)]
    print("prepare: first CTA start -> last CTA end %.1f us" % (float((tp[..., 1] - t0).max()) / 1e3))
    for j in range(JP):
        d = (tp[:, j, 1] - tp[:, j, 0]).float() / 1e3
        print("  part %-10s start max %5.1f us  duration mean %5.1f max %5.1f us" % (pn[j], float((tp[:, j, 0] - t0).max()) / 1e3, float(d.mean()), float(d.max())))

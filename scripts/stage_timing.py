"""Stage timestamps of one tile of the fused kernel (development aid): SM clock deltas of forward-role thread 0 and
backward-role thread 0 of CTA (0,0), third tile of the CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import dev as _dev; _dev.use_dev_library()   # include/cnf_dev.h hooks live in the development build
from bench import WORKLOADS, synth_batch
from category_nerf_reconstruction_official_b200 import _lib
from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
name = sys.argv[2] if len(sys.argv) > 2 else "c4"
wl = dict(WORKLOADS[name])
if name == "c4": wl["R"] = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda:0")
tr = FusedCategoryTrainer(wl["K"], wl["n_obj"], latent_dim=wl["L"], scale=wl["scale"], device=dev, n_bins_cam2surface=wl["N"], n_bins=wl["M"])
b = synth_batch(wl, seed=3, device=dev)
dbg = torch.zeros(256 + 8 * 1024, dtype=torch.int64, device=dev)
lib = _lib.load()
for i in range(3):
    tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
_lib.check(lib.cnf_set_debug_buffer(dbg.data_ptr()))
tr.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
torch.cuda.synchronize()
_lib.check(lib.cnf_set_debug_buffer(None))
d = dbg.cpu()
fn = ["start", "PE done (publish)"]
for st in ["L1", "L2", "L3", "L4", "L6", "L7"]:
    fn += [st + " wait-begin", st + " mma-done", st + " epi-done (publish)"]
fn += ["part 1b done (half 0)", "xg | xc copied (tile done)"]
bn = ["start (rays loaded)", "previous tile drained", "full[b] arrived (forward done)", "prologue done (publish)"]
for st in ["B8", "B7", "B6", "B4", "B3", "B2"]:
    bn += [st + " wait-begin", st + " mma-done", st + " epi-done (publish)"]
bn += ["B1 wait-begin", "B1 mma-done", "PE bwd done (publish)"]
for role, names, off in (("forward", fn, 0), ("backward", bn, 64)):
    t = d[off:off + 64].tolist()
    n = max([i for i, v in enumerate(t) if v] + [0]) + 1
    print(role, "stamps", n, "tile cycles", t[n - 1] - t[0])
    for i in range(1, n):
        print("   %-28s +%6d" % (names[i] if i < len(names) else str(i), t[i] - t[i - 1]))
print("forward start -> backward start of the same iteration: %d cycles" % (d[64] - d[0]))

f = d[128:192].tolist(); bk = d[192:256].tolist()
per = lambda x: [x[i + 1] - x[i] for i in range(len(x) - 1) if x[i + 1] and x[i]]
pf, pb = per(f), per(bk)
print("forward tile period (cycles), tiles 0..62:", pf[:6], "... median", sorted(pf)[len(pf) // 2] if pf else None)
print("backward tile period (cycles):", pb[:6], "... median", sorted(pb)[len(pb) // 2] if pb else None)
print("kernel start -> forward loop end %d, -> backward loop end %d, -> after flush %d (first forward tile at +%d)" %
      (d[60] - d[63], d[61] - d[63], d[62] - d[63], f[0] - d[63]))

c = d[256:].view(-1, 8)
c = c[c[:, 0] > 0]
if len(c):
    t0 = int(c[:, 0].min())
    import statistics as st
    f = lambda col: [(int(x) - t0) / 1e3 for x in c[:, col].tolist()]
    print("per-CTA wall clock (us after the first CTA started), %d CTAs: entry max %.1f | setup done median %.1f max %.1f | forward loop end median %.1f max %.1f | backward loop end median %.1f max %.1f | flush done median %.1f max %.1f" %
          (len(c), max(f(0)), st.median(f(1)), max(f(1)), st.median(f(2)), max(f(2)), st.median(f(3)), max(f(3)), st.median(f(4)), max(f(4))))
    print("tiles per CTA: min %d max %d" % (int(c[:, 5].min()), int(c[:, 5].max())))

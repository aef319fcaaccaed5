"""Stage stamps of the tensor-core OccupancyMap kernels (CTA 0, epilogue thread 0), in SM cycles (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import dev as _dev; _dev.use_dev_library()   # include/cnf_dev.h hooks live in the development build
from category_nerf_reconstruction_official_b200 import model, _lib
from category_nerf_reconstruction_official_b200._lib import check
dev = torch.device("cuda:0"); torch.manual_seed(0)
lib = _lib.load()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16800
m = model.OccupancyMap(87, 42, hidden_size=128).to(dev); m.apply(model.init_weights)
flat = model.flat_params(m).detach().contiguous()
emb = torch.randn(P, 129, device=dev)
ws = torch.empty(lib.cnf_occmap_tc_workspace_bytes(P), dtype=torch.uint8, device=dev)
alpha = torch.empty(P, device=dev); rgb = torch.empty(P, 3, device=dev)
ga = torch.randn(P, device=dev); gc = torch.randn(P, 3, device=dev); gp = torch.empty_like(flat); ge = torch.empty(P, 129, device=dev)
st = torch.cuda.current_stream().cuda_stream
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
def fwd(): check(lib.cnf_occmap_tc_fwd(P, flat.data_ptr(), emb.data_ptr(), alpha.data_ptr(), rgb.data_ptr(), ws.data_ptr(), ws.numel(), st))
def bwd(): check(lib.cnf_occmap_tc_bwd(P, flat.data_ptr(), rgb.data_ptr(), ga.data_ptr(), gc.data_ptr(), ws.data_ptr(), ws.numel(), gp.data_ptr(), ge.data_ptr(), st))
if len(sys.argv) > 2 and sys.argv[2] == "pts":      # embedding computed in-kernel
    pts = torch.rand(P, 3, device=dev) * 4 - 2; Bm = torch.randn(21, 3, device=dev); sc = torch.tensor([5.0], device=dev); gB = torch.empty(21, 3, device=dev)
    def fwd(): check(lib.cnf_occmap_tc_fwd_points(P, flat.data_ptr(), Bm.data_ptr(), sc.data_ptr(), pts.data_ptr(), alpha.data_ptr(), rgb.data_ptr(), ws.data_ptr(), ws.numel(), 0, st))
    def bwd(): check(lib.cnf_occmap_tc_bwd_points(P, flat.data_ptr(), pts.data_ptr(), rgb.data_ptr(), ga.data_ptr(), gc.data_ptr(), ws.data_ptr(), ws.numel(), gp.data_ptr(), gB.data_ptr(), 0, st))
for _ in range(3): fwd(); bwd()
check(lib.cnf_occmap_tc_set_debug(dbg.data_ptr()))
FW = ["tile start", "input staged", "L0 mma done", "fc1 written", "L1 mma done", "fc2 written", "L2 mma done", "fc3 written",
      "L3 mma done", "fc4 written", "L4 mma done", "h written"]
BW = ["tile start", "h loaded", "S0 dh written", "S1 dgrad done", "fc4 loaded", "S1 d fc4 written", "S1 wgrad done", "S1 drained",
      "S2 dgrad done", "fc3 loaded", "S2 d fc3 written", "S2 wgrad done", "S2 drained", "S3 dgrad done", "fc2 loaded", "S3 d fc2 written",
      "S3 wgrad done", "S3 drained", "S4 dgrad done", "fc1 loaded", "S4 d fc1 written", "S4 wgrad done", "S4 drained",
      "S5 dgrad done", "S5 wgrad done", "S5 drained"]
for name, fn, labels in (("forward", fwd, FW), ("backward", bwd, BW)):
    dbg.zero_(); fn(); torch.cuda.synchronize()
    v = dbg.cpu().tolist()
    print("%s (P=%d): first tile of CTA 0, %d cycles" % (name, P, v[len(labels) - 1] - v[0]))
    for i in range(1, len(labels)):
        print("   %-20s +%6d" % (labels[i], v[i] - v[i - 1]))
check(lib.cnf_occmap_tc_set_debug(None))

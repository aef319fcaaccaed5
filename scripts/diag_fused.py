"""Diagnostic: per-tensor error of the fused bf16 step vs golden (reference) for every case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import STEP_CASES, step_case
from oracle import cnf_oracle as O
from test_gpu_fused import _trainer_from_case, _run_case

def rel(a, e):
    a = a.double().cpu(); e = e.double().cpu()
    return float((a - e).norm() / (e.norm() + 1e-12)), float(torch.nn.functional.cosine_similarity(a.flatten(), e.flatten(), dim=0))

for name in STEP_CASES:
    c = step_case(name)
    tr = _trainer_from_case(c)
    loss = _run_case(c, tr)
    lt = tr.loss_terms()
    print("==", name, "loss", float(loss), "ref", float(c["loss"]))
    for k_ in ("depth", "color", "opacity"):
        print("   ", k_, lt[k_].cpu().tolist(), c["colorl" if k_ == "color" else k_].tolist())
    names = [n for n, _ in O.codenerf_param_shapes(32, c["L"])]
    for k in range(c["K"]):
        gv = tr.grad_views(k)
        row = []
        for i in range(0, 28, 2):
            r, cs = rel(gv[i], c["gfc"][i][k]); rb, _ = rel(gv[i + 1], c["gfc"][i + 1][k])
            row.append("%s %.3f/%.3f(b %.3f)" % (names[i].split(".")[0][:12], r, cs, rb))
        print("  cat", k, " | ".join(row))
        n = c["n_obj"]
        print("     B %.3f/%.3f  sc %.3f/%.3f  tc %.3f/%.3f" % (*rel(tr.g_B[k], c["gB"][k]), *rel(tr.g_shape[k, :n], c["gsc"][k]),
                                                              *rel(tr.g_tex[k, :n], c["gtc"][k])))

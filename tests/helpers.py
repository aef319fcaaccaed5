"""Shared helpers for the test-suite: golden fixture access."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def golden(group):
    if group not in _cache:
        _cache[group] = np.load(os.path.join(GOLDEN, group + ".npz"), allow_pickle=False)
    return _cache[group]


def gt(group, key):
    """golden array -> torch tensor"""
    a = golden(group)[key]
    return torch.from_numpy(np.array(a))


def step_case(name):
    """Load one 'step/<name>/' case as a dict of tensors + lists."""
    g = golden("step")
    p = "step/" + name + "/"
    K, n_obj, R, N, M, L = [int(v) for v in g[p + "cfg"]]
    d = dict(K=K, n_obj=n_obj, R=R, N=N, M=M, L=L, scale=float(g[p + "scale"]))
    d["fc"] = [gt("step", p + "fc%02d" % i) for i in range(28)]
    d["gfc"] = [gt("step", p + "gfc%02d" % i) for i in range(28)]
    d["B"] = gt("step", p + "B"); d["gB"] = gt("step", p + "gB")
    d["sc"] = [gt("step", p + "sc%d" % k) for k in range(K)]
    d["tc"] = [gt("step", p + "tc%d" % k) for k in range(K)]
    d["gsc"] = [gt("step", p + "gsc%d" % k) for k in range(K)]
    d["gtc"] = [gt("step", p + "gtc%d" % k) for k in range(K)]
    for kk in ["pcs", "z", "gt_depth", "gt_rgb", "mask_depth", "labels", "indices", "alpha", "color", "loss",
               "depth", "colorl", "opacity", "reg_shape", "reg_texture", "origins", "dirs"]:
        d[kk] = gt("step", p + kk)
    return d


STEP_CASES = ["K2_L32", "K1_L256", "K3_L32_S10", "K2_single_inst"]
SAMPLER_CASES = ["mixed_5_9", "mixed_1_9", "allvalid_obj", "allinvalid", "noobj"]


def assert_grad_close(actual, expected, rel_l2=2e-3, elem_rtol=2e-3, elem_atol=2e-5, max_bad_frac=0.1, name=""):
    """Gradient comparison robust to the L1 losses' sign() discontinuities: a ray whose residual is ~1 ulp from
    zero can flip sign between two fp32 evaluation orders and move a handful of gradient entries by a finite
    amount.  Require (a) small relative L2 error over the tensor and (b) elementwise closeness for all but a
    small fraction of entries."""
    a = actual.detach().double().cpu(); e = expected.detach().double().cpu()
    assert a.shape == e.shape, (name, a.shape, e.shape)
    den = float(e.norm()) + 1e-12
    err = float((a - e).norm()) / den
    atol_eff = elem_atol + 1e-3 * float(e.abs().max()) if e.numel() else elem_atol    # scale-aware: tensors span 1e-6..1e3
    bad = ((a - e).abs() > atol_eff + elem_rtol * e.abs()).double().mean().item() if a.numel() else 0.0
    assert err <= rel_l2 or float((a - e).abs().max()) <= elem_atol, "%s: relative L2 error %.3e > %.1e" % (name, err, rel_l2)
    assert bad <= max_bad_frac, "%s: %.2f%% of entries outside tolerance" % (name, 100 * bad)

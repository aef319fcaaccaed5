"""Driver of the UNMODIFIED reference modules staged by ``oracle/make_ref.py`` (or found under /root/reference).

TEST / BASELINE INFRASTRUCTURE ONLY: imported by ``bench.py``'s CPU legs (``--impl reference``, ``cpu_baseline``),
``tests/`` and ``tests/golden/*.py``.  Nothing here computes anything itself: every arithmetic op of the timed step is
executed by the reference's own files (``src/scene_cateogries.py`` sampler, ``src/trainer.py`` model/code construction,
``src/embedding.py``, ``src/model.py``, ``src/loss.py``, ``src/render_rays.py``, ``src/utils.py:24-28 update_vmap``)
and by ``functorch.vmap`` / ``combine_state_for_ensemble`` -- the sequence of calls is train.py:40,88-89,123-182.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(HERE, "_ref", "src")
_REF_NAMES = ["render_rays", "embedding", "model", "loss", "vis", "trainer", "utils", "scene_cateogries"]
_loaded = None


class _Stub(types.ModuleType):
    """Stand-in for the visualisation / mesh dependencies the hot path never calls (open3d, trimesh, skimage, imgviz, plotly)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


def find_ref_src():
    """Directory holding the reference's src/*.py: the staged copy (travels to the GPU box) or the read-only original."""
    if os.path.exists(os.path.join(_STAGED, "MANIFEST.json")):
        return _STAGED
    ref = os.path.join(os.environ.get("CNF_REFERENCE", "/root/reference"), "src")
    return ref if os.path.isdir(ref) else None


def load_ref():
    """Import the reference's modules (cached).  Returns a namespace of modules or None if no reference is available."""
    global _loaded
    if _loaded is not None:
        return _loaded
    src = find_ref_src()
    if src is None:
        return None
    warnings.filterwarnings("ignore")
    for n in ["skimage", "skimage.measure", "trimesh", "open3d", "imgviz", "plotly", "plotly.graph_objs", "plotly.subplots"]:
        if n not in sys.modules:
            try:
                importlib.import_module(n)
            except Exception:
                sys.modules[n] = _Stub(n)
    sys.path.insert(0, src)
    try:
        mods = {n: importlib.import_module(n) for n in _REF_NAMES}
    finally:
        sys.path.remove(src)
    for n, m in mods.items():       # make sure nothing of this repo was picked up under the same module name
        assert os.path.dirname(os.path.abspath(m.__file__)) == os.path.abspath(src), (n, m.__file__)
    import functorch
    ns = types.SimpleNamespace(src=src, vmap=functorch.vmap, **mods)
    _loaded = ns
    return ns


class _Cfg:
    """The attributes of src/cfg.py's Config that Trainer reads (trainer.py:12-27)."""

    def __init__(self, scale, latent_dim, device="cpu"):
        self.training_device = device
        self.obj_scale = scale
        self.n_unidir_funcs = 5
        self.hidden_feature_size = 32
        self.net_hyperparams = {"shape_blocks": 2, "texture_blocks": 1, "W": 32, "latent_dim": latent_dim}


class ReferenceStep:
    """K object categories built and stepped by the reference's own code on ``device`` (CPU by default).

    batch: dict(rgbs u8 [K,R,4], depth [K,R], origins [K,R,3], dirs [K,R,3], indices [K,R]) -- the rays
    ``get_training_samples`` would hand to ``sample_3d_points`` (scene_cateogries.py:421-436)."""

    def __init__(self, K, n_obj, N, M, L, scale, batch, seed=0, device="cpu", eps=0.1, stop_eps=0.05):
        R = load_ref()
        if R is None:
            raise RuntimeError("no reference available (neither oracle/_ref/src nor /root/reference)")
        self.R = R
        torch.manual_seed(seed)
        cfg = _Cfg(scale, L, device)
        self.K, self.n_rays = K, batch["depth"].shape[1]
        self.cls_dict = {}
        self.optimizer = torch.optim.AdamW([torch.autograd.Variable(torch.tensor([1]))], lr=1e-3, weight_decay=0.013)   # train.py:40
        fc_models, pe_models = [], []
        n_objs = [n_obj] * K if isinstance(n_obj, int) else list(n_obj)
        for k in range(K):
            sc = R.scene_cateogries.sceneCategory.__new__(R.scene_cateogries.sceneCategory)     # no dataset: bare instance
            sc.n_bins_cam2surface, sc.n_bins, sc.surface_eps, sc.stop_eps = N, M, eps, stop_eps
            sc.data_device = sc.training_device = device
            sc.min_bound, sc.this_obj = 0.0, 1
            sc.obj_ids = list(range(n_objs[k]))
            sc.trainer = R.trainer.Trainer(cfg, k + 1, sc.obj_ids)                              # scene_cateogries.py:349
            self.optimizer.add_param_group({"params": sc.trainer.shape_codes.parameters(), "lr": 1e-3, "weight_decay": 0.013})
            self.optimizer.add_param_group({"params": sc.trainer.texture_codes.parameters(), "lr": 1e-3, "weight_decay": 0.013})
            fc_models.append(sc.trainer.fc_occ_map)
            pe_models.append(sc.trainer.pe)
            self.cls_dict[k + 1] = sc
        self.fc_model, self.fc_param, self.fc_buffer = R.utils.update_vmap(fc_models, self.optimizer)   # train.py:88
        self.pe_model, self.pe_param, self.pe_buffer = R.utils.update_vmap(pe_models, self.optimizer)   # train.py:89
        self.batch = {k_: v.to(device) for k_, v in batch.items()}
        self.device = device

    def step(self, backward=True, optimizer_step=False):
        """train.py:123-184 for the object categories (no background, no logging)."""
        R, b = self.R, self.batch
        gd, grgb, dm, om, pcs, zs, shs, txs, cls_ids = [], [], [], [], [], [], [], [], []
        for k, (cls_id, cls_k) in enumerate(self.cls_dict.items()):
            gt_rgb, gt_depth, depth_mask, obj_mask, input_pcs, sampled_z = cls_k.sample_3d_points(
                b["rgbs"][k], b["depth"][k], b["origins"][k], b["dirs"][k])                     # scene_cateogries.py:435-436
            indices = b["indices"][k].long()
            gd.append(gt_depth); grgb.append(gt_rgb); dm.append(depth_mask); om.append(obj_mask)
            pcs.append(input_pcs); zs.append(sampled_z)
            cls_ids.append(np.array([cls_id], dtype=np.int32))
            shs.append(cls_k.trainer.shape_codes(indices)[:, None, :])                          # train.py:136-137
            txs.append(cls_k.trainer.texture_codes(indices)[:, None, :])
        st = torch.stack
        pcs, gd, grgb, dm, om, zs, shs, txs = st(pcs), st(gd), st(grgb) / 255., st(dm), st(om), st(zs), st(shs), st(txs)
        cls_ids = np.concatenate(cls_ids, axis=0)
        emb = R.vmap(self.pe_model)(self.pe_param, self.pe_buffer, pcs)                         # train.py:154
        alpha, color = R.vmap(self.fc_model)(self.fc_param, self.fc_buffer, emb, shs, txs)     # train.py:155
        loss, loss_dict, loss_col = R.loss.step_batch_loss(alpha, color, gd.detach(), grgb.detach(), om.detach(), dm.detach(),
                                                            zs.detach())                        # train.py:158-162
        rs, rt = R.loss.step_batch_loss_reg(self.cls_dict, torch.from_numpy(cls_ids).to(self.device))   # train.py:166
        loss = loss + 0.0005 * (rs + rt).sum()
        if backward:
            loss.backward()                                                                     # train.py:182
            if optimizer_step:
                self.optimizer.step()                                                           # train.py:183
            self.optimizer.zero_grad(set_to_none=True)                                          # train.py:184
        return loss.detach(), loss_dict


def time_reference(wl, batch, steps, warmup, seed=0, threads=None):
    """rays/s of the reference's own CPU path on ``wl`` (bench.WORKLOADS entry): (rays_per_s, s_per_step, threads)."""
    import time
    torch.set_num_threads(threads or os.cpu_count() or 1)
    rs = ReferenceStep(wl["K"], wl["n_obj"], wl["N"], wl["M"], wl["L"], wl["scale"], batch, seed)
    for _ in range(warmup):
        rs.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        rs.step()
    dt = (time.perf_counter() - t0) / steps
    return wl["K"] * wl["R"] / dt, dt, torch.get_num_threads()

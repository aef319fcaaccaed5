"""Checkpoint and log-tag interop with the reference.

The reference saves one file per category, ``<path>/cls_<cls_id>_iteration_<iter:05d>.pth``, holding a dict with the keys written
by ``sceneCategory.save_checkpoints`` (src/scene_cateogries.py:548-571) and read back by ``load_checkpoints`` (:573-597):

    global_step, PE_state_dict {B_layer.weight [21,3], scale 0-d}, FC_state_dict (nn.Linear names of src/model.py), cls_id,
    instance_id_to_index, obj_scale, bound, and for object categories obj_tensor_dict, [extent_dict], shape_code_state_dict {weight
    [n_obj, L]}, texture_code_state_dict {weight [n_obj, L]}.

The fused trainers keep their state as flat fp32 buffers in named_parameters() order, so a category's state dicts are views of those
buffers reshaped to the reference's tensors: ``export_category`` / ``import_category`` convert in both directions without touching
the values, and ``save_category`` / ``load_category_file`` write / read files the reference's own ``load_checkpoints`` /
``save_checkpoints`` accept.  ``log_scalars`` names the TensorBoard scalars as src/loss.py:76-102 does.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Optional

import torch

# named_parameters() order of src/model.py; shapes depend on (W, L) / hidden size only
CODENERF_KEYS = ["encoding_xyz.0", "shape_latent_layer_1.0", "shape_layer_1.0", "shape_latent_layer_2.0", "shape_layer_2.0",
                 "cat_layer.0", "cat_latent_layer.0", "encoding_shape", "sigma.0", "encoding_viewdir.0", "texture_layer_1.0",
                 "texture_latent_layer_1.0", "rgb.0", "rgb.2"]
OCCMAP_KEYS = ["in_layer.0", "mid1.0.0", "cat_layer.0", "mid2.0.0", "out_alpha", "color_linear.0", "out_color"]


def checkpoint_path(path: str, cls_id: int, iteration: int) -> str:
    """File name of src/scene_cateogries.py:549."""
    return path + "/cls_" + str(cls_id) + "_iteration_{:05d}.pth".format(iteration)


def _state_dict(keys, views) -> "OrderedDict[str, torch.Tensor]":
    assert len(views) == 2 * len(keys)
    sd = OrderedDict()
    for i, k in enumerate(keys):
        sd[k + ".weight"] = views[2 * i].detach().clone().cpu()
        sd[k + ".bias"] = views[2 * i + 1].detach().clone().cpu()
    return sd


def _pe_state_dict(B: torch.Tensor, scale) -> "OrderedDict[str, torch.Tensor]":
    """UniDirsEmbed.state_dict(): the persistent buffer ``scale`` first, then the trainable projection (src/embedding.py:75-80)."""
    sd = OrderedDict()
    sd["scale"] = torch.as_tensor(scale, dtype=torch.float32).detach().reshape(()).clone().cpu()
    sd["B_layer.weight"] = B.detach().clone().cpu()
    return sd


def export_category(trainer, k: int, cls_id: int, iteration: int, instance_id_to_index: Optional[Dict] = None, extras: Optional[Dict] = None) -> Dict:
    """The reference's checkpoint dict of category ``k`` of a FusedCategoryTrainer.  ``extras``: the geometry entries that belong to
    the scene object, not to the trainer (obj_tensor_dict, extent_dict, bound), passed through unchanged."""
    n = trainer.n_obj_list[k]
    d = {"global_step": iteration,
         "PE_state_dict": _pe_state_dict(trainer.pe_B[k], trainer.pe_scale[k]),
         "FC_state_dict": _state_dict(CODENERF_KEYS, trainer.param_views(k)),
         "cls_id": cls_id,
         "instance_id_to_index": dict(instance_id_to_index) if instance_id_to_index is not None else {i: i for i in range(n)},
         "obj_scale": float(trainer.pe_scale[k]),
         "shape_code_state_dict": OrderedDict(weight=trainer.shape_codes[k, :n].detach().clone().cpu()),
         "texture_code_state_dict": OrderedDict(weight=trainer.tex_codes[k, :n].detach().clone().cpu())}
    d.update(extras or {})
    d.setdefault("obj_tensor_dict", {})
    d.setdefault("bound", None)
    return d


def import_category(trainer, k: int, ckpt: Dict) -> int:
    """Load a reference checkpoint dict (object category) into slot ``k`` of a FusedCategoryTrainer; returns its global_step."""
    fc = ckpt["FC_state_dict"]
    params = []
    for key in CODENERF_KEYS:
        params += [fc[key + ".weight"], fc[key + ".bias"]]
    views = trainer.param_views(k)
    for v, p in zip(views, params):
        if tuple(v.shape) != tuple(p.shape):
            raise ValueError("checkpoint tensor %s does not fit this trainer (%s)" % (tuple(p.shape), tuple(v.shape)))
    sc, tc = ckpt["shape_code_state_dict"]["weight"], ckpt["texture_code_state_dict"]["weight"]
    if sc.shape[0] != trainer.n_obj_list[k] or sc.shape[1] != trainer.L:
        raise ValueError("checkpoint code table %s does not fit category %d (%d instances, latent %d)" % (tuple(sc.shape), k, trainer.n_obj_list[k], trainer.L))
    trainer.load_category(k, params, ckpt["PE_state_dict"]["B_layer.weight"], sc, tc)
    trainer.pe_scale[k] = float(ckpt["PE_state_dict"]["scale"])
    return int(ckpt.get("global_step", 0))


def export_background(trainer, iteration: int, extras: Optional[Dict] = None) -> Dict:
    """The reference's checkpoint dict of the background model (cls 0) of a FusedBackgroundTrainer."""
    d = {"global_step": iteration,
         "PE_state_dict": _pe_state_dict(trainer.pe_B, trainer.pe_scale[0]),
         "FC_state_dict": _state_dict(OCCMAP_KEYS, trainer.param_views()),
         "cls_id": 0, "instance_id_to_index": {0: 0}, "obj_scale": float(trainer.pe_scale[0])}
    d.update(extras or {})
    d.setdefault("bound", None)
    return d


def import_background(trainer, ckpt: Dict) -> int:
    fc = ckpt["FC_state_dict"]
    params = []
    for key in OCCMAP_KEYS:
        params += [fc[key + ".weight"], fc[key + ".bias"]]
    trainer.load(params, ckpt["PE_state_dict"]["B_layer.weight"])
    trainer.pe_scale[0] = float(ckpt["PE_state_dict"]["scale"])
    return int(ckpt.get("global_step", 0))


def module_checkpoint(trainer_module, cls_id: int, iteration: int, extras: Optional[Dict] = None) -> Dict:
    """Same dict from a module-level ``trainer.Trainer`` (drop-in of src/trainer.py)."""
    d = {"global_step": iteration, "PE_state_dict": trainer_module.pe.state_dict(), "FC_state_dict": trainer_module.fc_occ_map.state_dict(),
         "cls_id": cls_id, "instance_id_to_index": trainer_module.inst_id_to_index, "obj_scale": trainer_module.obj_scale}
    if cls_id != 0:
        d["shape_code_state_dict"] = trainer_module.shape_codes.state_dict()
        d["texture_code_state_dict"] = trainer_module.texture_codes.state_dict()
        d.setdefault("obj_tensor_dict", {})
    d.update(extras or {})
    d.setdefault("bound", None)
    return d


def load_module_checkpoint(trainer_module, ckpt: Dict) -> int:
    """What sceneCategory.load_checkpoints does to its trainer (src/scene_cateogries.py:581-596)."""
    trainer_module.fc_occ_map.load_state_dict(ckpt["FC_state_dict"])
    trainer_module.pe.load_state_dict(ckpt["PE_state_dict"])
    if ckpt["cls_id"] != 0:
        trainer_module.shape_codes.load_state_dict(ckpt["shape_code_state_dict"])
        trainer_module.texture_codes.load_state_dict(ckpt["texture_code_state_dict"])
    trainer_module.inst_id_to_index = ckpt["instance_id_to_index"]
    trainer_module.obj_scale = ckpt["obj_scale"]
    return int(ckpt["global_step"])


def save_category(trainer, k: int, path: str, cls_id: int, iteration: int, **kw) -> str:
    os.makedirs(path, exist_ok=True)
    f = checkpoint_path(path, cls_id, iteration)
    torch.save(export_category(trainer, k, cls_id, iteration, **kw), f)
    return f


def load_category_file(trainer, k: int, ckpt_file: str) -> int:
    return import_category(trainer, k, torch.load(ckpt_file, map_location="cpu", weights_only=False))


def log_scalars(loss_terms: Dict[str, torch.Tensor], cls_ids, background_terms: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, float]:
    """The TensorBoard scalars of the reference's logging step, by tag (src/loss.py:76-102: log_loss + log_psnr), from a
    FusedCategoryTrainer's ``loss_terms()`` (and optionally a FusedBackgroundTrainer's): ``cls_<id>/depth``, ``cls_<id>/color``,
    ``cls_<id>/opacity``, ``cls_<id>/reg_shape``, ``cls_<id>/reg_texture``, ``cls_<id>/psnr`` = -10 log10(colour loss), and
    ``background/<key>`` / ``background/psnr``.  ONE device->host read for everything (the reference does one ``.item()`` per
    scalar).  Feed the result to ``SummaryWriter.add_scalar(tag, value, iteration)``."""
    import math
    keys = [k for k in ("depth", "color", "opacity", "reg_shape", "reg_texture") if k in loss_terms]
    rows = [loss_terms[k].detach().reshape(-1).float() for k in keys]
    nb = 0
    if background_terms is not None:
        bkeys = [k for k in ("depth", "color", "opacity") if k in background_terms]
        rows += [background_terms[k].detach().reshape(-1).float()[:1].expand(rows[0].shape[0]) if rows else background_terms[k].detach().reshape(-1).float()
                 for k in bkeys]
        nb = len(bkeys)
    host = torch.stack(rows).cpu()
    out = {}
    for j, cid in enumerate(cls_ids):
        for i, k in enumerate(keys):
            out["cls_%s/%s" % (cid, k)] = float(host[i, j])
        if "color" in keys:
            c = float(host[keys.index("color"), j])
            out["cls_%s/psnr" % cid] = -10.0 * math.log(c) / math.log(10.0) if c > 0 else float("inf")
    for i in range(nb):
        v = float(host[len(keys) + i, 0])
        out["background/" + bkeys[i]] = v
        if bkeys[i] == "color":
            out["background/psnr"] = -10.0 * math.log(v) / math.log(10.0) if v > 0 else float("inf")
    return out

"""Trainer -- drop-in for the model/PE/code construction and point evaluation of the reference's
src/trainer.py:11-60,125-151 (meshing/visualisation stays with the reference: CPU post-processing, out of scope)."""
import math

import numpy as np
import torch
import torch.nn as nn

from . import embedding, model, render_rays


class Trainer:
    """Per-category model state with the attribute names the reference's callers and checkpoints use (src/trainer.py:11-60):
    ``fc_occ_map`` (CodeNeRF for object categories, OccupancyMap for cls 0), ``pe`` (UniDirsEmbed), ``shape_codes`` /
    ``texture_codes`` (nn.Embedding, one row per instance, N(0, 2/latent_dim) init), ``inst_id_to_index``, ``n_obj``, ``obj_scale``,
    ``bound_extent`` (0.995 background / 0.9 objects), ``extent_dict``, ``scale_template``."""

    EMB_GEOMETRY = 3 + 21 * 4          # raw point + 4 octaves of the 21 directions (input of the geometry branch)
    EMB_TOTAL = 3 + 21 * 6             # + 2 more octaves for the colour branch

    def __init__(self, cfg, cls_id, inst_ids):
        inst_ids = list(inst_ids)
        self.cls_id, self.n_obj = cls_id, len(inst_ids)
        self.inst_id_to_index = {inst: row for row, inst in enumerate(inst_ids)}
        if len(self.inst_id_to_index) != self.n_obj:
            raise ValueError("duplicate instance ids in %r" % (inst_ids,))
        self.device = cfg.training_device
        self.obj_scale, self.n_unidir_funcs = cfg.obj_scale, cfg.n_unidir_funcs
        self.emb_size1, self.emb_size2 = self.EMB_GEOMETRY, self.EMB_TOTAL - self.EMB_GEOMETRY
        is_background = cls_id == 0
        if is_background:
            self.hidden_feature_size = cfg.hidden_feature_size
            self.load_NeRF()
        else:
            self.net_hyperparams = cfg.net_hyperparams
            self.load_codeNeRF()
            self.load_codes()
        self.bound_extent = 0.995 if is_background else 0.9
        self.extent_dict = self.scale_template = None
        self.use_tensor_cores = True      # eval_points / render_rays through the tcgen05 kernels (False: fp32 CUDA-core modules)

    def _make_pe(self):
        return embedding.UniDirsEmbed(max_deg=self.n_unidir_funcs, scale=self.obj_scale).to(self.device)

    def load_codeNeRF(self):
        net = model.CodeNeRF(self.emb_size1, self.emb_size2, **self.net_hyperparams)
        net.apply(model.init_weights)                  # xavier-normal weights, default-uniform biases (model.py:4-6)
        self.fc_occ_map, self.pe = net.to(self.device), self._make_pe()

    def load_NeRF(self):
        net = model.OccupancyMap(self.emb_size1, self.emb_size2, hidden_size=self.hidden_feature_size)
        net.apply(model.init_weights)
        self.fc_occ_map, self.pe = net.to(self.device), self._make_pe()

    def load_codes(self):
        dim = self.net_hyperparams["latent_dim"]
        inv_std = math.sqrt(dim / 2)
        # RNG consumption as the reference (trainer.py:52-60), so that the same seed gives the same initial codes: both tables are
        # constructed first (nn.Embedding draws its own, discarded, init), then the N(0, 2/dim) weights, shape before texture
        tables = {name: nn.Embedding(self.n_obj, dim) for name in ("shape_codes", "texture_codes")}
        for name, table in tables.items():
            table.weight = nn.Parameter(torch.randn(self.n_obj, dim) / inv_std)
            setattr(self, name, table.to(self.device))

    def eval_points(self, points, inst_id=None, chunk_size=500000):
        """points [P,3] -> (occupancy [P], colour [P,3]) with one instance's codes (trainer.py:125-151).
        Outputs are written chunk by chunk into preallocated tensors (no Python list of 0-d tensors)."""
        P = points.shape[0]
        alpha = torch.empty(P, device=points.device)
        color = torch.empty(P, 3, device=points.device)
        if self.cls_id != 0:
            obj_idx = torch.from_numpy(np.array(self.inst_id_to_index[inst_id])).to(self.device)
            shape_code, texture_code = self.shape_codes(obj_idx), self.texture_codes(obj_idx)
        if self.cls_id != 0 and points.is_cuda and getattr(self.fc_occ_map, "W", 32) == 32 and self.use_tensor_cores:
            # tensor-core forward (bf16 operands, fp32 accumulation): one launch for all points, no chunking needed
            from . import model as _model, ops
            with torch.no_grad():
                occ, color = ops.eval_points_fused(points, _model.flat_params(self.fc_occ_map), self.pe.B_layer.weight, self.pe.scale,
                                                   shape_code, texture_code)
            if occ.max() == 0:
                print("no occ")
                return None
            return (occ, color)
        if self.cls_id == 0 and points.is_cuda and getattr(self.fc_occ_map, "hidden_size", 0) == 128 and self.use_tensor_cores:
            # background model on tensor cores (embedding computed in-kernel): one launch for all points
            from . import model as _model, ops
            with torch.no_grad():
                occ, color = ops.occmap_eval_points_fused(points, _model.flat_params(self.fc_occ_map), self.pe.B_layer.weight, self.pe.scale)
            if occ.max() == 0:
                print("no occ")
                return None
            return (occ, color)
        with torch.no_grad():
            for s in range(0, P, chunk_size):
                emb = self.pe(points[s:s + chunk_size, ...])
                if self.cls_id == 0:
                    a, c = self.fc_occ_map(emb)
                else:
                    a, c = self.fc_occ_map(emb, shape_code, texture_code)
                alpha[s:s + chunk_size] = a.reshape(-1)
                color[s:s + chunk_size] = c
        occ = render_rays.occupancy_activation(alpha).detach()
        if occ.max() == 0:
            print("no occ")
            return None
        return (occ, color)

    def render_rays(self, origins, dirs, z_vals, inst_id, want_weights=False):
        """Forward-only rendering of rays through one instance's field (extension; the reference renders views with un-shipped
        scripts and builds its uncertainty fields inline, category_registration.py:146-152): origins/dirs [R,3], z_vals [R,S] ->
        dict(depth, var, rgb, opacity[, weights = termination probabilities]).  CodeNeRF categories on a CUDA device only
        (tensor-core kernel cnf_render_rays; there is no CPU path)."""
        from . import model as _model, ops
        if self.cls_id == 0:
            if getattr(self.fc_occ_map, "hidden_size", 0) != 128:
                raise NotImplementedError("render_rays: the tensor-core background path needs hidden_size == 128")
            with torch.no_grad():
                return ops.occmap_render_rays_fused(origins, dirs, z_vals, _model.flat_params(self.fc_occ_map), self.pe.B_layer.weight,
                                                    self.pe.scale, want_weights)
        obj_idx = torch.from_numpy(np.array(self.inst_id_to_index[inst_id])).to(self.device)
        with torch.no_grad():
            return ops.render_rays_fused(origins, dirs, z_vals, _model.flat_params(self.fc_occ_map), self.pe.B_layer.weight,
                                         self.pe.scale, self.shape_codes(obj_idx), self.texture_codes(obj_idx), want_weights)

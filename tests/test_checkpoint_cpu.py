"""CPU: checkpoint / log-tag interop with the REFERENCE's own classes.  The reference's modules come from
oracle/_ref (staged by oracle/make_ref.py in the authoring container; travels to the GPU box) or /root/reference; when neither is
present the reference-side half is skipped and only the schema is checked."""
import os

import pytest
import torch

from category_nerf_reconstruction_official_b200 import checkpoint, trainer
from oracle import ref_driver


class Cfg:
    training_device = "cpu"
    obj_scale = 2.0
    n_unidir_funcs = 5
    hidden_feature_size = 128
    net_hyperparams = {"shape_blocks": 2, "texture_blocks": 1, "W": 32, "latent_dim": 32}


REF_KEYS_OBJ = {"global_step", "PE_state_dict", "FC_state_dict", "cls_id", "instance_id_to_index", "obj_scale", "obj_tensor_dict",
                "shape_code_state_dict", "texture_code_state_dict", "bound"}       # scene_cateogries.py:550-567 (+ extent_dict if > 1 instance)


def test_trainer_constructor_fields_and_init():
    """trainer.Trainer builds what src/trainer.py:11-60 builds: attribute names, shapes, code init N(0, 2/L) drawn shape-first."""
    torch.manual_seed(3)
    t = trainer.Trainer(Cfg, 4, [11, 5, 8])
    assert t.inst_id_to_index == {11: 0, 5: 1, 8: 2} and t.n_obj == 3 and t.cls_id == 4
    assert (t.emb_size1, t.emb_size2) == (87, 42) and t.bound_extent == 0.9 and t.extent_dict is None and t.scale_template is None
    assert sum(p.numel() for p in t.fc_occ_map.parameters()) == 18116 and t.pe.B_layer.weight.shape == (21, 3)
    assert t.shape_codes.weight.shape == (3, 32) and abs(float(t.shape_codes.weight.std()) - (1 / 16) ** 0.5) < 0.06
    bg = trainer.Trainer(Cfg, 0, [0])
    assert bg.bound_extent == 0.995 and sum(p.numel() for p in bg.fc_occ_map.parameters()) == 94340 and not hasattr(bg, "shape_codes")
    R = ref_driver.load_ref()
    if R is not None:                       # identical RNG consumption: same seed -> same initial state as the reference's Trainer
        torch.manual_seed(3)
        r = R.trainer.Trainer(Cfg, 4, [11, 5, 8])
        for (n1, p1), (n2, p2) in zip(t.fc_occ_map.named_parameters(), r.fc_occ_map.named_parameters()):
            assert n1 == n2 and torch.equal(p1, p2), n1
        assert torch.equal(t.shape_codes.weight, r.shape_codes.weight) and torch.equal(t.texture_codes.weight, r.texture_codes.weight)
        assert torch.equal(t.pe.B_layer.weight, r.pe.B_layer.weight)


def test_checkpoint_round_trip_through_reference_classes(tmp_path):
    """A checkpoint written from this repo's Trainer loads (strict) into the reference's Trainer modules and back, bit-identically;
    the dict has the reference's keys and file name."""
    torch.manual_seed(0)
    t = trainer.Trainer(Cfg, 7, [3, 9])
    d = checkpoint.module_checkpoint(t, 7, 2000, extras={"obj_tensor_dict": {3: torch.zeros(10)}, "bound": None})
    assert REF_KEYS_OBJ <= set(d.keys())
    f = checkpoint.checkpoint_path(str(tmp_path), 7, 2000)
    assert os.path.basename(f) == "cls_7_iteration_02000.pth"
    torch.save(d, f)
    ck = torch.load(f, map_location="cpu", weights_only=False)
    assert list(ck["PE_state_dict"].keys()) == ["scale", "B_layer.weight"]
    assert list(ck["shape_code_state_dict"].keys()) == ["weight"]
    R = ref_driver.load_ref()
    if R is None:
        pytest.skip("no reference modules available (oracle/_ref not staged)")
    torch.manual_seed(99)
    r = R.trainer.Trainer(Cfg, 7, [3, 9])                                   # different random state ...
    r.fc_occ_map.load_state_dict(ck["FC_state_dict"])                        # ... overwritten as load_checkpoints does (:581-590)
    r.pe.load_state_dict(ck["PE_state_dict"])
    r.shape_codes.load_state_dict(ck["shape_code_state_dict"]); r.texture_codes.load_state_dict(ck["texture_code_state_dict"])
    for (n1, p1), (n2, p2) in zip(t.fc_occ_map.named_parameters(), r.fc_occ_map.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)
    assert torch.equal(t.pe.B_layer.weight, r.pe.B_layer.weight) and float(r.pe.scale) == 2.0
    # and the other way: a dict built the way the reference's save_checkpoints builds it loads into this repo's Trainer
    ref_dict = {"global_step": 4000, "PE_state_dict": r.pe.state_dict(), "FC_state_dict": r.fc_occ_map.state_dict(), "cls_id": 7,
                "instance_id_to_index": r.inst_id_to_index, "obj_scale": r.obj_scale, "obj_tensor_dict": {},
                "shape_code_state_dict": r.shape_codes.state_dict(), "texture_code_state_dict": r.texture_codes.state_dict(), "bound": None}
    torch.manual_seed(5)
    t2 = trainer.Trainer(Cfg, 7, [3, 9])
    assert checkpoint.load_module_checkpoint(t2, ref_dict) == 4000
    assert torch.equal(t2.shape_codes.weight, r.shape_codes.weight) and torch.equal(t2.fc_occ_map.sigma[0].weight, r.fc_occ_map.sigma[0].weight)


def test_log_scalar_tags_match_reference_logging():
    """Tags of src/loss.py:76-102 (log_loss / log_psnr) recorded through a fake SummaryWriter against checkpoint.log_scalars."""
    K = 3
    terms = {"loss": torch.tensor(9.0), "depth": torch.tensor([1.0, 2.0, 3.0]), "color": torch.tensor([0.1, 0.2, 0.3]),
             "opacity": torch.tensor([0.01, 0.02, 0.03]), "reg_shape": torch.tensor([4.0, 5.0, 6.0]), "reg_texture": torch.tensor([7.0, 8.0, 9.0])}
    bg = {"depth": torch.tensor([0.5]), "color": torch.tensor([0.25]), "opacity": torch.tensor([0.125])}
    mine = checkpoint.log_scalars(terms, [4, 7, 9], background_terms=bg)
    assert mine["cls_7/color"] == pytest.approx(0.2) and mine["background/opacity"] == pytest.approx(0.125)
    R = ref_driver.load_ref()
    if R is None:
        pytest.skip("no reference modules available")
    import numpy as np

    class W:
        def __init__(self): self.s = {}
        def add_scalar(self, tag, v, it): self.s[tag] = float(v)
    w = W()
    d = {k: v for k, v in terms.items() if k != "loss"}
    d["cls_ids"] = np.array([4, 7, 9]); d["background"] = {k: v[0] for k, v in bg.items()}
    R.loss.log_loss(w, d, 100)
    R.loss.log_psnr(w, np.array([4, 7, 9]), terms["color"], 100, bg_loss_col=bg["color"][0])
    assert set(w.s) == set(mine), set(w.s) ^ set(mine)
    for tag, v in w.s.items():
        assert mine[tag] == pytest.approx(v, rel=1e-6), tag

"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/cnf.h declares
(no compute calls -- there is no GPU here).  Also: the product refuses to run without CUDA (no fallback)."""
import os
import re

import pytest
import torch

import category_nerf_reconstruction_official_b200 as cnf
from category_nerf_reconstruction_official_b200 import _lib, model, embedding, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return set(re.findall(r"\b(cnf_[a-z0-9_]+)\s*\(", hdr))


def test_build_and_load_exports_every_declared_symbol():
    cnf.build()
    lib = cnf.load()
    declared = _declared("cnf.h")
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cnf_version() >= 100
    if os.environ.get("CNF_LIB_PATH") is None:
        for name in _lib.DEV_SIGNATURES:        # the product library exports no development hook
            assert not hasattr(lib, name), name


def test_development_build_exports_product_and_dev_symbols():
    """include/cnf_dev.h (self-tests, probes, clock stamps) lives only in libcnf_sm100_dev.so = the same sources with -DCNF_DEV."""
    cnf.build_dev()
    lib = cnf.load_dev()
    declared = _declared("cnf_dev.h")
    assert declared == set(_lib.DEV_SIGNATURES), (declared ^ set(_lib.DEV_SIGNATURES))
    for name in declared | set(_lib.SIGNATURES):
        assert hasattr(lib, name), name


def test_param_counts_match_reference_layout():
    lib = cnf.load()
    assert lib.cnf_codenerf_param_count(32, 256) == 46788     # SURVEY 2 (#3)
    assert lib.cnf_codenerf_param_count(32, 32) == 18116
    assert lib.cnf_occmap_param_count(128) == 94340           # SURVEY 2 (#4)
    m = model.CodeNeRF(87, 42, W=32, latent_dim=32)
    assert sum(p.numel() for p in m.parameters()) == 18116
    assert model.flat_params(m).shape == (18116,)


def test_no_cpu_fallback():
    pe = embedding.UniDirsEmbed(max_deg=5, scale=2.0)
    with pytest.raises(cnf.CnfError):
        pe(torch.zeros(4, 3))
    m = model.CodeNeRF(87, 42, W=32, latent_dim=32)
    with pytest.raises(cnf.CnfError):
        m(torch.zeros(4, 2, 129), torch.zeros(4, 1, 32), torch.zeros(4, 1, 32))


def test_no_cpu_fallback_tensor_core_paths():
    """The tensor-core entry points (background model, evaluation / rendering, ray set-up) refuse CPU tensors loudly."""
    from category_nerf_reconstruction_official_b200 import ops, sampling
    bg = model.OccupancyMap(87, 42, hidden_size=128)
    bg.use_tensor_cores = True
    with pytest.raises(cnf.CnfError):
        bg(torch.zeros(4, 129))
    flat = model.flat_params(bg).detach()
    with pytest.raises(cnf.CnfError):
        ops.occmap_eval_points_fused(torch.zeros(4, 3), flat, torch.zeros(21, 3), torch.tensor(5.0))
    with pytest.raises(cnf.CnfError):
        ops.occmap_render_rays_fused(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 14), flat, torch.zeros(21, 3), torch.tensor(5.0))
    with pytest.raises(cnf.CnfError):
        sampling.origin_dirs_O(torch.eye(4).repeat(3, 1, 1), torch.zeros(3, 3))
    with pytest.raises((cnf.CnfError, RuntimeError, AssertionError)):
        from category_nerf_reconstruction_official_b200.background import FusedBackgroundTrainer
        FusedBackgroundTrainer(device="cpu")


def test_tensor_core_kernels_use_bulk_copies_and_tmem():
    """The OccupancyMap kernels carry the TMA-engine bulk copies (UBLKCP) next to the tensor-core instructions."""
    import subprocess
    cnf.build()
    out = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3cnf3otc11eval_kernelILb0EEEviPKhPKfPfS6_iS5_S5_", _lib.LIB_PATH],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    if not out.strip():      # mangled name changed: scan the whole library
        out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "UTCHMMA" in out and "UBLKCP" in out and "LDTM" in out


def test_state_dict_keys_match_reference():
    from helpers import golden
    names = [str(s) for s in golden("step")["step/param_names"]]
    m = model.CodeNeRF(87, 42, W=32, latent_dim=32)
    assert [n for n, _ in m.named_parameters()] == names
    names = [str(s) for s in golden("bg")["bg/H64/param_names"]]
    o = model.OccupancyMap(87, 42, hidden_size=64)
    assert [n for n, _ in o.named_parameters()] == names
    pe = embedding.UniDirsEmbed(max_deg=5, scale=2.0)
    assert sorted(pe.state_dict().keys()) == [str(s) for s in golden("embed")["embed/state_keys"]]
    assert torch.equal(pe.B_layer.weight.data, torch.from_numpy(golden("embed")["embed/B_init"]))


def test_sass_contains_tcgen05():
    """The built library really carries Blackwell tensor-core code (UTCHMMA / LDTM), not a legacy path."""
    import subprocess
    cnf.build()
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "UTCHMMA" in out and "LDTM" in out
    assert "HMMA.16816" not in out       # no mma.sync fallback

"""B200-native (sm_100a) implementation of the neural-field training hot path of
Taekbum/category-nerf-reconstruction-official, behind the reference's own Python API.

    from category_nerf_reconstruction_official_b200 import embedding, model, render_rays, loss, sampling
    from category_nerf_reconstruction_official_b200.ensemble import update_vmap, vmap
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer

All compute goes through ``csrc/libcnf_sm100.so`` (C ABI in ``include/cnf.h``); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import CnfError, build, build_dev, load, load_dev  # noqa: F401

__all__ = ["build", "load", "build_dev", "load_dev", "CnfError"]

"""Category batching -- replaces the reference's functorch ensemble (src/utils.py:24-28 update_vmap;
train.py:88-89,154-155 ``vmap(fmodel)(params, buffers, ...)``).

``update_vmap(models, optimiser)`` returns ``(fmodel, params, buffers)`` exactly like the reference: ``params`` is
a tuple of leaf tensors stacked ``[K, ...]`` in named_parameters() order (added to the optimiser as one param
group), so ``train.py:196-201``'s copy-back by index keeps working.  ``vmap(fmodel)`` returns a callable that
runs ALL K categories in one category-batched kernel sequence (blockIdx.z = category)."""
import torch

from . import ops
from .embedding import UniDirsEmbed
from .model import CodeNeRF, OccupancyMap, codenerf_apply


class _Ensemble:
    def __init__(self, kind, meta):
        self.kind = kind
        self.meta = meta

    def batched(self, params, buffers, *args):
        K = params[0].shape[0]
        if self.kind == "pe":
            (x,) = args                                    # [K, R, S, 3]
            scale = buffers[1].reshape(K).to(x.device)
            emb = ops.embed(x.reshape(K, -1, 3), params[0], scale)
            return emb.reshape(*x.shape[:-1], ops.EMB_DIM)
        flat = torch.cat([p.reshape(K, -1) for p in params], dim=1)
        if self.kind == "codenerf":
            emb, sh, tx = args                             # [K,R,S,129], [K,R,1,L], [K,R,1,L]
            return codenerf_apply(emb, sh, tx, flat, self.meta["W"], batched=True)
        if self.kind == "occmap":
            (emb,) = args
            lead = emb.shape[1:-1]
            a, c = ops.OccMapFn.apply(emb.reshape(K, -1, ops.EMB_DIM), flat, self.meta["H"])
            return a.reshape(K, *lead, 1), c.reshape(K, *lead, 3)
        raise TypeError(self.kind)


def combine_state_for_ensemble(models):
    """Stack K identically-shaped modules: (fmodel, params tuple [K,...], buffers tuple [K,...])."""
    m0 = models[0]
    if isinstance(m0, CodeNeRF):
        ens = _Ensemble("codenerf", {"W": m0.W, "L": m0.latent_dim})
    elif isinstance(m0, OccupancyMap):
        ens = _Ensemble("occmap", {"H": m0.hidden_size})
    elif isinstance(m0, UniDirsEmbed):
        ens = _Ensemble("pe", {})
    else:
        raise TypeError("unsupported module for the category ensemble: %s" % type(m0).__name__)
    names = [n for n, _ in m0.named_parameters()]
    params = tuple(torch.stack([dict(m.named_parameters())[n].detach() for m in models]) for n in names)
    bnames = [n for n, _ in m0.named_buffers()]
    buffers = tuple(torch.stack([dict(m.named_buffers())[n] for m in models]) for n in bnames)
    return ens, params, buffers


def update_vmap(models, optimiser):
    fmodel, params, buffers = combine_state_for_ensemble(models)
    [p.requires_grad_() for p in params]
    optimiser.add_param_group({"params": params})
    return (fmodel, params, buffers)


def vmap(fmodel):
    """``vmap(fmodel)(params, buffers, *batched_args)`` -- only ensembles built by update_vmap are supported."""
    if not isinstance(fmodel, _Ensemble):
        raise TypeError("vmap here batches category ensembles from update_vmap(); use torch.func.vmap for anything else")
    return fmodel.batched

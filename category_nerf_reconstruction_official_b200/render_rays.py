"""Drop-in for the reference's src/render_rays.py:3-95 (same function names and signatures).

The per-ray scan (occupancy_to_termination) runs in libcnf_sm100.so; the one-line reducers are kept as the
thin tensor expressions they are in the reference.  The training hot path does not come through here: it
uses loss.step_batch_loss (one fused composite+loss kernel) or the fused tcgen05 step."""
import torch

from . import ops


def occupancy_activation(alpha, distances=None):
    return torch.sigmoid(alpha)


class _TerminationFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, occupancy):
        ctx.save_for_backward(occupancy)
        return ops.termination(occupancy)

    @staticmethod
    def backward(ctx, g):
        (occupancy,) = ctx.saved_tensors
        return ops.termination_bwd(occupancy, g.contiguous())


def occupancy_to_termination(occupancy, is_batch=False):
    """[n_rays, S] (or [K, n_rays, S] with is_batch=True) -> termination weights of the same shape."""
    return _TerminationFn.apply(occupancy)


def ray_uncertainty(termination):
    """Uncertainty fields of category_registration.py:152-162 from termination weights [n_rays, S]:
    (entropy [n_rays], termination mass [n_rays], heuristic = mass * exp(-0.5 entropy) [n_rays])."""
    from ._lib import check, lib_for, ptr, stream
    w = termination.detach().float().contiguous()
    lib = lib_for(w)
    n, S = w.shape
    out = torch.empty(n, 3, device=w.device, dtype=torch.float32)
    check(lib.cnf_ray_uncertainty(n, S, ptr(w), ptr(out), stream()))
    return out[:, 0], out[:, 1], out[:, 2]


def render(termination, vals, dim=-1):
    """Weighted sum of per-sample values along the sample axis (render_rays.py:46-50)."""
    return torch.sum(termination * vals, dim=dim)


def render_loss(render, gt, loss="L1", normalise=False):
    """Per-ray residual magnitude (render_rays.py:52-64); callers on the hot path only use L1."""
    kinds = {"L1": torch.abs, "L2": torch.square}
    if loss not in kinds:
        raise ValueError("loss type {} not implemented!".format(loss))
    out = kinds[loss](render - gt)
    return out / gt if normalise else out


def reduce_batch_loss(loss_mat, var=None, avg=True, mask=None, loss_type="L1"):
    """Masked mean per category with optional 1/(sqrt(var)+1e-4) weighting (render_rays.py:66-95).
    Quirk kept: if ANY category has an empty mask the term is zero for ALL categories.
    Deviation: a term above 1e5 raises FloatingPointError instead of calling exit(-1)."""
    n_valid = mask.sum(dim=-1)
    if bool((n_valid == 0).any()):
        z = torch.zeros_like(loss_mat)
        return z.mean(dim=-1) if avg else z
    if var is not None:
        spread = var if loss_type == "L2" else var.sqrt()
        loss_mat = loss_mat * (1.0 / (spread + 1e-4))
    if not avg:
        return loss_mat
    if mask is None:
        return loss_mat.mean(dim=-1).sum()
    out = loss_mat.sum(dim=-1) / (n_valid + 1e-10)
    if bool((out > 1e5).any()):
        raise FloatingPointError("loss explode")
    return out

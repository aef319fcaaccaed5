"""Drop-in for the reference's src/loss.py:5-74: step_batch_loss and step_batch_loss_reg, same signatures and
return values, computed by one fused compositing+loss kernel pair (cnf_composite_loss_fwd/bwd)."""
import torch

from . import ops


def step_batch_loss(alpha, color, gt_depth, gt_color, sem_labels, mask_depth, z_vals,
                    color_scaling=5.0, opacity_scaling=10.0):
    """alpha [K,R,S,1], color [K,R,S,3], gt_depth [K,R], gt_color [K,R,3] in 0..1, sem_labels u8 [K,R],
    mask_depth bool [K,R], z_vals [K,R,S] -> (loss, {'depth','color','opacity'} each [K], loss_col [K])."""
    K = alpha.shape[0]
    a = alpha.squeeze(dim=-1)
    loss, out, _rays = ops.CompositeLossFn.apply(a, color, z_vals, gt_depth, gt_color, sem_labels, mask_depth,
                                                 color_scaling, opacity_scaling)
    loss_dict = {'depth': out[4:4 + K], 'color': out[4 + K:4 + 2 * K], 'opacity': out[4 + 2 * K:4 + 3 * K]}
    return loss, loss_dict, loss_dict['color']


def step_batch_loss_reg(cls_dict, cls_ids):
    """Per category: sum over ALL instances of ||shape_code||_2 and ||texture_code||_2 (0 if one instance)."""
    regs_s, regs_t = [], []
    for idx, cls_k in enumerate(cls_dict.values()):
        tr = cls_k.trainer
        regs_s.append(ops.CodeRegFn.apply(tr.shape_codes.weight))
        regs_t.append(ops.CodeRegFn.apply(tr.texture_codes.weight))
    return torch.stack(regs_s), torch.stack(regs_t)

"""CodeNeRF and OccupancyMap -- drop-ins for the reference's src/model.py:22-154.

Module/parameter names (hence state_dict keys and named_parameters() order) are the
reference's; forward/backward run in libcnf_sm100.so (cnf_codenerf_* / cnf_occmap_*)."""
import torch
import torch.nn as nn

from . import ops


def init_weights(m, init_fn=torch.nn.init.xavier_normal_):
    if type(m) == torch.nn.Linear:
        init_fn(m.weight)


def fc_block(in_f, out_f):
    return torch.nn.Sequential(torch.nn.Linear(in_f, out_f), torch.nn.ReLU(out_f))


def flat_params(module: nn.Module) -> torch.Tensor:
    """[n_param] concatenation of the module's parameters in named_parameters() order (autograd-tracked)."""
    return torch.cat([p.reshape(-1) for p in module.parameters()])


class CodeNeRF(nn.Module):
    def __init__(self, emb_size1, emb_size2, shape_blocks=2, texture_blocks=1, W=32, latent_dim=32):
        super().__init__()
        if (emb_size1, emb_size2) != (ops.EMB1, ops.EMB2) or shape_blocks != 2 or texture_blocks != 1:
            raise NotImplementedError("kernels implement the shipped architecture: emb 87/42, shape_blocks=2, "
                                      "texture_blocks=1 (every reference config)")
        self.shape_blocks = shape_blocks
        self.texture_blocks = texture_blocks
        self.embedding_size1 = emb_size1
        self.embedding_size2 = emb_size2
        self.W = W
        self.latent_dim = latent_dim
        self.encoding_xyz = nn.Sequential(nn.Linear(emb_size1, W), nn.ReLU())
        for j in range(shape_blocks):
            setattr(self, f"shape_latent_layer_{j+1}", nn.Sequential(nn.Linear(latent_dim, W), nn.ReLU()))
            setattr(self, f"shape_layer_{j+1}", nn.Sequential(nn.Linear(W, W), nn.ReLU()))
        self.cat_layer = nn.Sequential(nn.Linear(W + emb_size1, W), nn.ReLU())
        self.cat_latent_layer = nn.Sequential(nn.Linear(latent_dim, W), nn.ReLU())
        self.encoding_shape = nn.Linear(W, W)
        self.sigma = nn.Sequential(nn.Linear(W, 1))
        self.encoding_viewdir = nn.Sequential(nn.Linear(W + emb_size2, W), nn.ReLU())
        for j in range(texture_blocks):
            setattr(self, f"texture_layer_{j+1}", nn.Sequential(nn.Linear(W, W), nn.ReLU()))
            setattr(self, f"texture_latent_layer_{j+1}", nn.Sequential(nn.Linear(latent_dim, W), nn.ReLU()))
        self.rgb = nn.Sequential(nn.Linear(W, W // 2), nn.ReLU(), nn.Linear(W // 2, 3), nn.Sigmoid())

    def forward(self, x, shape_latent, texture_latent, noise_std=None, do_cat=True):
        """x [..., 129]; latents broadcastable to x[..., :1, L]-style rows: either [R,1,L] for x [R,S,129]
        (train.py:136-137,155) or a single code [L] for x [P,129] (trainer.py:140)."""
        if not do_cat:
            raise NotImplementedError("do_cat=False is never used by the reference's callers")
        sig, rgb = codenerf_apply(x, shape_latent, texture_latent, flat_params(self)[None], self.W, batched=False)
        if noise_std is not None:
            sig = sig + torch.randn(sig.shape, device=x.device) * noise_std * 10.
        return sig, rgb


def codenerf_apply(x, shape_latent, texture_latent, params, W, batched):
    """Shared shape plumbing. batched=True: leading K on every argument (vmap semantics)."""
    if not batched:
        x, shape_latent, texture_latent = x[None], shape_latent[None], texture_latent[None]
    K = x.shape[0]
    L = shape_latent.shape[-1]
    lead = x.shape[1:-1]
    P = 1
    for d in lead:
        P *= d
    if shape_latent.dim() == 2:                      # [K, L]: one code for all points (eval_points)
        NR, ppr = 1, P
    else:                                            # [K, R, 1, L] rows broadcast over the S samples
        assert shape_latent.shape[-2] == 1 and len(lead) >= 1
        NR = shape_latent[0].numel() // L
        assert P % NR == 0, "latent rows must divide the number of points"
        ppr = P // NR
    sig, rgb = ops.CodeNeRFFn.apply(x.reshape(K, P, ops.EMB_DIM), shape_latent.reshape(K, NR, L),
                                    texture_latent.reshape(K, NR, L), params, W, ppr)
    sig = sig.reshape(K, *lead, 1)
    rgb = rgb.reshape(K, *lead, 3)
    return (sig, rgb) if batched else (sig[0], rgb[0])


class OccupancyMap(torch.nn.Module):
    # True: hidden_size == 128 models (the reference's background model) run on tcgen05 tensor cores (bf16 operands, fp32
    # accumulation; gradients match a bf16-rounding emulation to 1e-3, the fp32 path to ~6 %).  False (default): fp32 CUDA cores.
    use_tensor_cores = False

    def __init__(self, emb_size1, emb_size2, hidden_size=256, do_color=True, hidden_layers_block=1):
        super().__init__()
        if (emb_size1, emb_size2) != (ops.EMB1, ops.EMB2) or hidden_layers_block != 1 or not do_color:
            raise NotImplementedError("kernels implement the shipped architecture: emb 87/42, one hidden block, colour on")
        self.do_color = do_color
        self.embedding_size1 = emb_size1
        self.embedding_size2 = emb_size2
        self.hidden_size = hidden_size
        self.in_layer = fc_block(emb_size1, hidden_size)
        self.mid1 = torch.nn.Sequential(*[fc_block(hidden_size, hidden_size) for _ in range(hidden_layers_block)])
        self.cat_layer = fc_block(hidden_size + emb_size1, hidden_size)
        self.mid2 = torch.nn.Sequential(*[fc_block(hidden_size, hidden_size) for _ in range(hidden_layers_block)])
        self.out_alpha = torch.nn.Linear(hidden_size, 1)
        self.color_linear = fc_block(emb_size2 + hidden_size, hidden_size)
        self.out_color = torch.nn.Linear(hidden_size, 3)
        self.sigmoid = torch.sigmoid

    def forward(self, x, noise_std=None, do_alpha=True, do_color=True, do_cat=True):
        if not do_cat:
            raise NotImplementedError("do_cat=False is never used by the reference's callers")
        lead = x.shape[:-1]
        if self.use_tensor_cores and self.hidden_size == 128 and x.numel() > 0:
            alpha, rgb = ops.OccMapTCFn.apply(x.reshape(-1, ops.EMB_DIM), flat_params(self))     # bf16 tcgen05 path
        else:
            alpha, rgb = ops.OccMapFn.apply(x.reshape(1, -1, ops.EMB_DIM), flat_params(self)[None], self.hidden_size)
        alpha = alpha.reshape(*lead, 1)
        if noise_std is not None:
            alpha = alpha + torch.randn(alpha.shape, device=x.device) * noise_std * 10.
        return (alpha if do_alpha else None), (rgb.reshape(*lead, 3) if do_color else None)

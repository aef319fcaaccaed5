"""UniDirsEmbed -- drop-in for the reference's src/embedding.py:43-92 (same constructor, parameters, buffers and
state_dict keys {'B_layer.weight', 'scale'}); forward/backward run in libcnf_sm100.so (cnf_embed_fwd/bwd)."""
import torch

from . import ops

# 21 icosahedron directions: initial value of the trainable projection B_layer.weight (embedding.py:51-73)
_PHI_A, _PHI_B = 0.8506508, 0.5257311
_DIRS = [
    (_PHI_A, 0, _PHI_B), (0.809017, 0.5, 0.309017), (_PHI_B, _PHI_A, 0), (1, 0, 0), (0.809017, 0.5, -0.309017),
    (_PHI_A, 0, -_PHI_B), (0.309017, 0.809017, -0.5), (0, _PHI_B, -_PHI_A), (0.5, 0.309017, -0.809017), (0, 1, 0),
    (-_PHI_B, _PHI_A, 0), (-0.309017, 0.809017, -0.5), (0, _PHI_B, _PHI_A), (-0.309017, 0.809017, 0.5),
    (0.309017, 0.809017, 0.5), (0.5, 0.309017, 0.809017), (0.5, -0.309017, 0.809017), (0, 0, 1),
    (-0.5, 0.309017, 0.809017), (-0.809017, 0.5, 0.309017), (-0.809017, 0.5, -0.309017),
]


class UniDirsEmbed(torch.nn.Module):
    def __init__(self, min_deg=0, max_deg=2, scale=2.):
        super().__init__()
        self.min_deg = min_deg
        self.max_deg = max_deg
        self.n_freqs = max_deg - min_deg + 1
        if min_deg != 0 or self.n_freqs != 6:
            raise NotImplementedError("the sm_100a kernels implement the shipped configuration min_deg=0, max_deg=5 "
                                      "(n_unidir_funcs=5 in every reference config); got %d..%d" % (min_deg, max_deg))
        self.tensor_scale = torch.tensor(scale, requires_grad=False)
        self.B_layer = torch.nn.Linear(3, 21, bias=False)
        self.B_layer.weight.data = torch.tensor(_DIRS, dtype=torch.float32)
        frequency_bands = 2.0 ** torch.linspace(self.min_deg, self.max_deg, self.n_freqs)
        self.register_buffer("frequency_bands", frequency_bands, persistent=False)
        self.register_buffer("scale", self.tensor_scale, persistent=True)

    def forward(self, x):
        lead = x.shape[:-1]
        emb = ops.embed(x.reshape(1, -1, 3), self.B_layer.weight[None], self.scale.reshape(1).to(x.device))
        return emb.reshape(*lead, ops.EMB_DIM)

"""Synthetic ray batches with the value distributions of SURVEY 8(d) (pure data generation on the CPU; used by
bench.py and the smoke test so that the CPU baseline and the GPU path see identical bits)."""
import torch


def synth_rays(R: int, n_obj: int, gen: torch.Generator, invalid_frac: float = 0.15, label_mix=(0.60, 0.30, 0.10)):
    """One category: rgbs u8 [R,4] (last column = label: 1 this object, 0 other, 2 unknown), depth [R] (0 = invalid),
    origins [R,3], dirs [R,3] (camera rays through a random rotation, scaled like a sim3), indices int64 [R]."""
    depth = torch.rand(R, generator=gen) * 3.0 + 0.5
    depth = torch.where(torch.rand(R, generator=gen) < invalid_frac, torch.zeros_like(depth), depth)
    rgb = torch.randint(0, 256, (R, 3), generator=gen, dtype=torch.int64).to(torch.uint8)
    r = torch.rand(R, generator=gen)
    lab = torch.where(r < label_mix[0], 1, torch.where(r < label_mix[0] + label_mix[1], 0, 2)).to(torch.uint8)
    rgbs = torch.cat([rgb, lab[:, None]], dim=1)
    u = torch.rand(R, generator=gen) * 1200
    v = torch.rand(R, generator=gen) * 680
    dc = torch.stack([(u - 599.5) / 600.0, (v - 339.5) / 600.0, torch.ones(R)], -1)
    q = torch.randn(4, generator=gen)
    w_, x_, y_, z_ = (q / q.norm()).tolist()
    rot = torch.tensor([[1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_)],
                        [2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_)],
                        [2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_)]])
    s = torch.rand(1, generator=gen).item() * 0.7 + 0.3
    dirs = (dc @ rot.t()) / (s * 4.0)
    o = torch.randn(R, 3, generator=gen)
    origins = o / o.norm(dim=-1, keepdim=True) * (torch.rand(R, 1, generator=gen) * 0.5 + 0.5)
    idx = (torch.arange(R) % n_obj)[torch.randperm(R, generator=gen)].to(torch.int64)
    return rgbs, depth, origins.contiguous(), dirs.contiguous(), idx

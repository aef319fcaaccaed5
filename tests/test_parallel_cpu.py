"""CPU (gloo, world_size=2): host-side multi-GPU logic -- category sharding, ray slicing, count/max/gradient
all-reduces -- and the additivity the ray-split mode relies on, checked with the CPU oracle as the compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from category_nerf_reconstruction_official_b200 import parallel
from oracle import cnf_oracle as O


def test_shard_categories_balanced():
    costs = [3600.0] * 20 + [16800.0 * 6.9]      # 20 object categories + a heavy background-like one
    sh = parallel.shard_categories(costs, 8)
    assert sorted(k for r in sh for k in r) == list(range(21))
    loads = [sum(costs[k] for k in r) for r in sh]
    assert max(loads) == pytest.approx(16800.0 * 6.9)          # the heavy one sits alone
    rest = sorted(loads)[:-1]
    assert max(rest) - min(rest) <= 3600.0
    assert parallel.shard_categories([1.0] * 5, 1) == [[0, 1, 2, 3, 4]]


def test_split_rays_covers_everything():
    for n, w in ((524288, 8), (361, 4), (7, 8)):
        seen = []
        for r in range(w):
            s = parallel.split_rays(n, w, r)
            seen += list(range(n))[s]
        assert seen == list(range(n))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(5)
    K, n_obj, R, N, M, L = 1, 3, 40, 1, 9, 32
    fc = [p[None] for p in O.init_params(O.codenerf_param_shapes(32, L), g)]
    B = O.icosa_matrix()[None]
    sc = [O.init_codes(n_obj, L, g)]; tc = [O.init_codes(n_obj, L, g)]
    rgbs, depth, org, dirs, idx = O.synth_rays(R, n_obj, g)
    rand = O.draw_sampler_randoms(rgbs, depth, N, M, 0.1, g)
    # single-process reference: the full batch
    _, _, valid, labels, pcs, z = O.sample_points(rgbs, depth, org, dirs, N, M, 0.1, 0.05, rand)
    full = O.train_step(fc, B, 2.0, sc, tc, pcs[None], z[None], depth[None], rgbs[None, :, :3], labels[None], valid[None], idx[None])
    # ray split: this rank's slice; global counts / max depth via the collectives under test
    sl = parallel.split_rays(R, world, rank)
    md = parallel.allreduce_max_depth(depth[None, sl])
    assert float(md) == float(depth.max())
    lab = labels[sl]; va = valid[sl]
    local_counts = torch.tensor([[int((va & (lab != 0)).sum()), int((lab != 0).sum()), int((lab != 2).sum()), 0]], dtype=torch.int32)
    counts = parallel.allreduce_counts(local_counts)
    assert counts[0, :3].tolist() == [int((valid & (labels != 0)).sum()), int((labels != 0).sum()), int((labels != 2).sum())]
    # the ONE-message variant RaySplitStep uses: all-gather of [K,5] int32 (counts + max-depth bits), reduced locally
    msg = parallel.pack_counts_and_max(local_counts, depth[None, sl].max(dim=-1).values)
    gathered = torch.empty(world * 1, 5, dtype=torch.int32)
    dist.all_gather_into_tensor(gathered, msg)
    c2, m2 = parallel.reduce_counts_and_max(gathered.view(world, 1, 5))
    assert torch.equal(c2, counts) and float(m2) == float(depth.max())
    # local step with GLOBAL denominators: scale the local masked sums by (local count / global count)
    fcl = [p.clone().requires_grad_(True) for p in fc]
    Bl = B.clone().requires_grad_(True)
    scl = [c.clone().requires_grad_(True) for c in sc]; tcl = [c.clone().requires_grad_(True) for c in tc]
    emb = O.embed(pcs[None, sl], Bl, 2.0)
    a, c = O.codenerf_forward(fcl, emb, scl[0][idx[sl]][None, :, None, :], tcl[0][idx[sl]][None, :, None, :])
    cf_mo = (lab != 0); cf_ms = (lab != 2); cf_md = va & cf_mo
    occ = torch.sigmoid(a.squeeze(-1)); w = O.termination(occ)
    D = (w * z[None, sl]).sum(-1); var = (w * (z[None, sl] - D[..., None]) ** 2).sum(-1).detach()
    C = (w[..., None] * c).sum(-2); Oo = w.sum(-1)
    gt = rgbs[sl, :3].float() / 255.
    ld = (torch.abs(D - depth[None, sl]) * cf_md / (torch.sqrt(var) + 1e-4)).sum() / (counts[0, 0] + 1e-10)
    lc = (torch.abs(C - gt[None]).sum(-1) * cf_mo).sum() / (counts[0, 1] + 1e-10)
    lo = (torch.abs(Oo - cf_mo.float()) * cf_ms).sum() / (counts[0, 2] + 1e-10)
    loss = ld + 5 * lc + 10 * lo
    if rank == 0:
        rs, rt = O.reg_loss([(scl[0], tcl[0])])
        loss = loss + 0.0005 * (rs + rt).sum()
    loss.backward()
    grads = [p.grad for p in fcl] + [Bl.grad, scl[0].grad, tcl[0].grad, loss.detach().reshape(1).clone()]
    parallel.allreduce_gradients(grads)
    ok = True
    for gsum, ref in zip(grads[:28], full["g_fc"]):
        ok &= bool(torch.allclose(gsum, ref, rtol=1e-3, atol=1e-5))
    ok &= bool(torch.allclose(grads[28], full["g_B"], rtol=1e-3, atol=1e-4))
    ok &= bool(torch.allclose(grads[29], full["g_shape"][0], rtol=1e-3, atol=1e-6))
    ok &= bool(torch.allclose(grads[30], full["g_texture"][0], rtol=1e-3, atol=1e-6))
    ok &= bool(torch.allclose(grads[31][0], full["loss"], rtol=1e-4, atol=1e-5))
    ret[rank] = ok
    dist.destroy_process_group()


def test_ray_split_gradients_are_additive_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)

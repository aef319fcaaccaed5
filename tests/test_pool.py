"""Ray pools: per-step slicing, ray set-up and epoch reshuffle of sceneCategory.get_training_samples
(src/scene_cateogries.py:352-451).  tests/golden/pool.npz holds what the reference's own get_training_samples returned on bare
instances (flat pool with several instances -> origin_dirs_O, flat pool with one instance -> origin_dirs_W, background pool with a
per-frame pose table) over 20-25 steps (two epoch reshuffles each), and the permutations it drew.

CPU: the oracle's RayPool against the golden.  GPU: the device-resident pools (cnf_pool_gather / cnf_pool_compose) against the
golden: ray data and instance indices bit-exact, origins / directions to fp32 rounding of the 4x4 inverse."""
import pytest
import torch

from oracle import cnf_oracle as O
from helpers import golden, gt

CASES = ["multi", "single", "background"]
# origins / directions: the reference inverts with LAPACK (partial pivoting), the kernel by cofactors; both are fp32
POSE_RTOL, POSE_ATOL = 2e-5, 2e-6


def _pool_inputs():
    return {k: gt("pool", "pool/in/" + k) for k in ("rgbs", "depth", "ray_dirs", "indices", "pose_table", "pose_index", "poses")}


def _perms(tag):
    P, n, steps, n_perm = [int(v) for v in golden("pool")["pool/%s/meta" % tag]]
    return P, n, steps, [gt("pool", "pool/%s/perm%d" % (tag, e)) for e in range(n_perm)]


def _check(tag, t, got, exact_pose):
    p = "pool/%s/step%d/" % (tag, t)
    rgbs, depth, origins, dirs, idx = [g.cpu() for g in got]
    assert torch.equal(rgbs, gt("pool", p + "rgbs")), (tag, t)
    assert torch.equal(depth, gt("pool", p + "depth")), (tag, t)
    assert torch.equal(idx.to(torch.int64), gt("pool", p + "idx")), (tag, t)
    if exact_pose:
        assert torch.equal(origins, gt("pool", p + "origins")) and torch.equal(dirs, gt("pool", p + "dirs"))
    else:
        torch.testing.assert_close(origins, gt("pool", p + "origins"), rtol=POSE_RTOL, atol=POSE_ATOL)
        torch.testing.assert_close(dirs, gt("pool", p + "dirs"), rtol=POSE_RTOL, atol=POSE_ATOL)


@pytest.mark.parametrize("tag", CASES)
def test_oracle_pool_vs_reference(tag):
    pin = _pool_inputs()
    P, n, steps, perms = _perms(tag)
    assert len(perms) >= 2
    if tag == "background":
        pool = O.RayPool(pin["rgbs"], pin["depth"], pin["ray_dirs"], pin["pose_table"], None, frames=pin["pose_index"], invert=False)
    else:
        pool = O.RayPool(pin["rgbs"], pin["depth"], pin["ray_dirs"], pin["poses"], pin["indices"], invert=(tag == "multi"))
    e = 0
    for t in range(steps):
        will_shuffle = pool.i_batch + n >= P - n
        got = pool.get_batch(n, perms[e] if will_shuffle else None)
        e += int(will_shuffle)
        _check(tag, t, got, exact_pose=True)
    assert e == len(perms)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_device_pool_vs_reference(tag):
    from category_nerf_reconstruction_official_b200 import pool as P_
    pin = _pool_inputs()
    P, n, steps, perms = _perms(tag)
    dev = "cuda:0"
    if tag == "background":
        sc = P_.BackgroundRayPool(pin["rgbs"], pin["depth"], pin["pose_index"], pin["ray_dirs"], pin["pose_table"], data_device=dev)
    else:
        sc = P_.CategoryRayPool(pin["rgbs"], pin["depth"], pin["ray_dirs"], pin["poses"], pin["indices"], multi_instance=(tag == "multi"),
                                data_device=dev)
    e = 0
    for t in range(steps):
        will_shuffle = sc.pools.i_batch[0] + n >= P - n
        got = sc.pools.next_batch_of(0, n, perms[e] if will_shuffle else None)
        e += int(will_shuffle)
        _check(tag, t, got, exact_pose=False)
    assert e == len(perms) and sc.pools.epochs[0] == e


@pytest.mark.gpu
def test_device_pools_batched_categories_and_sampler():
    """K categories with pools of different sizes in one launch == the per-category oracle pools, over several epochs with the
    device-drawn permutations read back; then get_training_samples (pool slice + sampler) against the oracle sampler on the slice."""
    from category_nerf_reconstruction_official_b200 import pool as P_
    g = torch.Generator().manual_seed(3)
    sizes = [700, 1000, 513]
    pools = [O.synth_pool(p, 3, 5, g) for p in sizes]
    dp = P_.DeviceRayPools([dict(rgbs=p["rgbs"], depth=p["depth"], ray_dirs=p["ray_dirs"], poses=p["poses"], indices=p["indices"])
                            for p in pools], invert=True, device="cuda:0")
    ora = [O.RayPool(p["rgbs"], p["depth"], p["ray_dirs"], p["poses"], p["indices"], invert=True) for p in pools]
    n = 120
    for t in range(14):
        shuf = [o.i_batch + n >= s - n for o, s in zip(ora, sizes)]
        rand_idx = [torch.randperm(s, generator=g) if sh else None for s, sh in zip(sizes, shuf)]
        rgbs, depth, origins, dirs, idx = [v.cpu() for v in dp.next_batch(n, rand_idx)]
        for k, o in enumerate(ora):
            r, d, og, di, ix = o.get_batch(n, rand_idx[k])
            assert torch.equal(rgbs[k], r) and torch.equal(depth[k], d) and torch.equal(idx[k], ix)
            torch.testing.assert_close(origins[k], og, rtol=POSE_RTOL, atol=POSE_ATOL)
            torch.testing.assert_close(dirs[k], di, rtol=POSE_RTOL, atol=POSE_ATOL)
        assert dp.i_batch == [o.i_batch for o in ora]
    assert min(dp.epochs) >= 1
    # default path: permutation drawn on the device; every pool row is still served exactly once per epoch
    dp2 = P_.DeviceRayPools([dict(rgbs=pools[0]["rgbs"], depth=pools[0]["depth"], ray_dirs=pools[0]["ray_dirs"], poses=pools[0]["poses"],
                                  indices=pools[0]["indices"])], device="cuda:0")
    for t in range(9):
        dp2.next_batch(70)
    assert sorted(dp2.perm.cpu().tolist()) == list(range(700)) and dp2.epochs[0] == 1
    # drop-in get_training_samples: the sampler sees exactly the pool slice
    p = pools[1]
    sc = P_.CategoryRayPool(p["rgbs"], p["depth"], p["ray_dirs"], p["poses"], p["indices"], n_bins_cam2surface=1, n_bins=9,
                            data_device="cuda:0")
    R, N, M = 200, 1, 9
    rand = {"u_inv": torch.rand(R, N + M, generator=g), "u_val": torch.rand(R, N, generator=g),
            "n_obj": torch.randn(R, M, generator=g) * (0.1 / 3), "u_oth": torch.rand(R, M, generator=g)}
    gt_rgb, gt_depth, valid, labels, pcs, z, inst = sc.get_training_samples(R, rand={k: v.cuda() for k, v in rand.items()})
    o = O.RayPool(p["rgbs"], p["depth"], p["ray_dirs"], p["poses"], p["indices"], invert=True)
    r, d, og, di, ix = o.get_batch(R)
    # sampler inputs: the device origins/dirs (the 4x4 inverse differs from LAPACK in the last bits), so that z / pcs are bit-exact
    sc2 = P_.CategoryRayPool(p["rgbs"], p["depth"], p["ray_dirs"], p["poses"], p["indices"], data_device="cuda:0")
    _, _, og_d, di_d, _ = sc2.pools.next_batch_of(0, R)
    rgb_o, d_o, valid_o, lab_o, pcs_o, z_o = O.sample_points(r, d, og_d.cpu(), di_d.cpu(), N, M, 0.1, 0.05, rand)
    assert torch.equal(z.cpu(), z_o) and torch.equal(pcs.cpu(), pcs_o)
    assert torch.equal(valid.cpu(), valid_o) and torch.equal(labels.cpu(), lab_o) and torch.equal(inst.cpu(), ix)
    assert torch.equal(gt_rgb.cpu(), r[:, :3]) and torch.equal(gt_depth.cpu(), d)

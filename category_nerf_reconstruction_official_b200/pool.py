"""Device-resident ray pools -- the per-step batch slicing of the reference's sceneCategory.get_training_samples
(src/scene_cateogries.py:352-451) without the host in the loop.

The reference keeps every category's rays in shuffled CPU pools (rgbs u8 [P,4], depth [P], camera-frame directions [P,3], a 4x4
pose per ray or a frame index into a pose table, instance indices), copies a slice of each to the device every step, inverts one
4x4 matrix per ray and physically re-permutes the pools at the end of an epoch.  Here the pools of ALL categories live on the GPU
as one structure of arrays; a step's batch for K categories is ONE launch (`cnf_pool_gather`: slice + pose inverse + ray
transform), and the epoch reshuffle `pool = pool[rand_idx]` is a composition of permutations over pool rows (`cnf_pool_compose`,
4 B per ray instead of ~100).  `i_batch` stays a host integer exactly as in the reference: it is a function of the call sequence
only, so there is no device->host synchronisation anywhere.

Extension for parity tests: `rand_idx=` injects the epoch permutation (the reference draws `torch.randperm(P)` on the CPU)."""
from typing import List, Optional, Sequence

import torch

from . import sampling
from ._lib import check, lib_for, ptr, stream


class DeviceRayPools:
    """Ray pools of K categories on one device.  Each entry of ``pools`` is a dict with
    rgbs u8 [P,4], depth [P], ray_dirs [P,3], poses [P,4,4] (per ray) or [F,4,4] with ``pose_index`` int [P],
    optional ``indices`` int64 [P] (instance index of each ray; zeros when absent).
    ``invert`` True: poses are object->camera T_CO and rays go to the object frame (origin_dirs_O, multi-instance categories);
    False: poses are camera->world T_WC (origin_dirs_W, background / single-instance categories)."""

    def __init__(self, pools: Sequence[dict], invert: bool = True, device="cuda:0"):
        self.device = torch.device(device)
        self.K = len(pools)
        self.invert = bool(invert)
        sizes = [int(p["depth"].shape[0]) for p in pools]
        self.sizes = sizes
        dev = self.device
        f32 = dict(device=dev, dtype=torch.float32)
        self.rgbs = torch.cat([p["rgbs"].reshape(-1, 4).to(torch.uint8) for p in pools]).to(dev).contiguous()
        self.depth = torch.cat([p["depth"].reshape(-1) for p in pools]).to(**f32).contiguous()
        self.ray_dirs = torch.cat([p["ray_dirs"].reshape(-1, 3) for p in pools]).to(**f32).contiguous()
        self.indices = torch.cat([p["indices"].reshape(-1).to(torch.int64) if p.get("indices") is not None
                                  else torch.zeros(n, dtype=torch.int64) for p, n in zip(pools, sizes)]).to(dev).contiguous()
        per_ray = [p.get("pose_index") is None for p in pools]
        assert all(per_ray) or not any(per_ray), "either every pool has one pose per ray or every pool has a pose table"
        poses, pidx, base = [], [], 0
        for p, n in zip(pools, sizes):
            T = p["poses"].reshape(-1, 4, 4)
            if p.get("pose_index") is None:
                assert T.shape[0] == n, "per-ray poses: one 4x4 matrix per pool row"
            else:
                pidx.append(p["pose_index"].reshape(-1).to(torch.int64) + base)
            poses.append(T)
            base += T.shape[0]
        self.poses = torch.cat(poses).to(**f32).contiguous()
        self.pose_index = torch.cat(pidx).to(dev, torch.int32).contiguous() if pidx else None
        seg = [0]
        for n in sizes:
            seg.append(seg[-1] + n)
        assert seg[-1] < 2 ** 31
        self._seg_host = seg
        self.seg = torch.tensor(seg[:-1], dtype=torch.int64, device=dev)
        self.perm = torch.arange(seg[-1], dtype=torch.int32, device=dev)      # pool row at each (shuffled) position
        self._perm_tmp = torch.empty_like(self.perm)
        self.i_batch = [0] * self.K                                             # scene_cateogries.py: self.i_batch
        self.epochs = [0] * self.K

    # -------------------------------------------------------------------------------------------------------------------------
    def _gather(self, cats: Sequence[int], n: int):
        lib = lib_for(self.depth)
        K = len(cats)
        dev = self.device
        # i_batch of the categories: a few bytes of pageable host memory (staged by the driver at call time, so the host may go on
        # and change i_batch before the copy has run on the stream)
        seg = self.seg if list(cats) == list(range(self.K)) else self.seg[list(cats)].contiguous()
        start = torch.tensor([self.i_batch[k] for k in cats], dtype=torch.int64).to(dev)
        rgbs = torch.empty(K, n, 4, dtype=torch.uint8, device=dev)
        depth = torch.empty(K, n, dtype=torch.float32, device=dev)
        origins = torch.empty(K, n, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(K, n, 3, dtype=torch.float32, device=dev)
        inst = torch.empty(K, n, dtype=torch.int64, device=dev)
        check(lib.cnf_pool_gather(K, n, int(self.invert), ptr(seg), ptr(start), ptr(self.perm), ptr(self.rgbs), ptr(self.depth),
                                  ptr(self.ray_dirs), ptr(self.poses), ptr(self.pose_index), ptr(self.indices), ptr(rgbs), ptr(depth),
                                  ptr(origins), ptr(dirs), ptr(inst), stream()))
        return rgbs, depth, origins, dirs, inst

    def _advance(self, k: int, n: int, rand_idx: Optional[torch.Tensor]):
        """`self.i_batch += n`, then the reference's epoch test and reshuffle (scene_cateogries.py:437-449)."""
        P = self.sizes[k]
        self.i_batch[k] += n
        if self.i_batch[k] >= P - n:
            if rand_idx is None:
                rand_idx = torch.randperm(P, device=self.device)
            rand_idx = rand_idx.to(self.device, torch.int64).contiguous()
            assert rand_idx.numel() == P
            lo, hi = self._seg_host[k], self._seg_host[k + 1]
            lib = lib_for(self.depth)
            check(lib.cnf_pool_compose(P, ptr(self.perm[lo:hi]), ptr(rand_idx), ptr(self._perm_tmp[lo:hi]), stream()))
            self.perm[lo:hi].copy_(self._perm_tmp[lo:hi])
            self.i_batch[k] = 0
            self.epochs[k] += 1

    def next_batch(self, n: int, rand_idx: Optional[List[Optional[torch.Tensor]]] = None):
        """This step's rays of every category: (rgbs u8 [K,n,4], depth [K,n], origins [K,n,3], dirs [K,n,3], indices i64 [K,n]);
        then advances every category's i_batch (and reshuffles the categories whose epoch ended).  One launch."""
        for k in range(self.K):
            assert self.i_batch[k] + n <= self.sizes[k], "category %d: pool of %d rays is smaller than the batch" % (k, self.sizes[k])
        out = self._gather(range(self.K), n)
        for k in range(self.K):
            self._advance(k, n, None if rand_idx is None else rand_idx[k])
        return out

    def next_batch_of(self, k: int, n: int, rand_idx: Optional[torch.Tensor] = None):
        """One category's slice (un-batched shapes [n,...]), as one call of the reference's get_training_samples consumes it."""
        assert self.i_batch[k] + n <= self.sizes[k]
        rgbs, depth, origins, dirs, inst = self._gather([k], n)
        self._advance(k, n, rand_idx)
        return rgbs[0], depth[0], origins[0], dirs[0], inst[0]


class CategoryRayPool(sampling.RaySampler):
    """The sampling-facing half of the reference's sceneCategory for one (non-background) category: `get_training_samples(n)`
    with the reference's return tuple (scene_cateogries.py:421-451, the flat-pool branch), pools resident on the device.
    ``multi_instance`` False = the single-object case (`t_wc_batch_all`, origin_dirs_W, :430-431)."""

    def __init__(self, rgbs_batch_all, depth_batch_all, ray_dirs_batch_all, pose_batch_all, batch_indices_all, multi_instance=True,
                 n_bins_cam2surface=1, n_bins=9, surface_eps=0.1, stop_eps=0.05, data_device="cuda:0", min_bound=0.0, this_obj=1):
        super().__init__(n_bins_cam2surface, n_bins, surface_eps, stop_eps, data_device, min_bound, this_obj)
        self.pools = DeviceRayPools([dict(rgbs=rgbs_batch_all, depth=depth_batch_all, ray_dirs=ray_dirs_batch_all, poses=pose_batch_all,
                                          indices=batch_indices_all)], invert=multi_instance, device=data_device)

    @property
    def i_batch(self):
        return self.pools.i_batch[0]

    def get_training_samples(self, n_samples, rand=None, rand_idx=None, exact_rng=False):
        rgbs, depth, origins, dirs, inst = self.pools.next_batch_of(0, n_samples, rand_idx)
        gt_rgb, gt_depth, valid, labels, pcs, z = self.sample_3d_points(rgbs, depth, origins, dirs, rand=rand, exact_rng=exact_rng)
        return gt_rgb, gt_depth, valid, labels, pcs, z, inst


class BackgroundRayPool(sampling.RaySampler):
    """`get_training_samples` of the background category (cls_id 0, scene_cateogries.py:353-418 with one pool): rays carry a frame
    index into the per-frame camera->world table `t_wc_batch_dict[0]`."""

    def __init__(self, rgbs_batch, depth_batch, frame_batch, ray_dirs_batch, t_wc_batch, n_bins_cam2surface=5, n_bins=9, surface_eps=0.1,
                 stop_eps=0.05, data_device="cuda:0", min_bound=0.0, this_obj=1):
        super().__init__(n_bins_cam2surface, n_bins, surface_eps, stop_eps, data_device, min_bound, this_obj)
        self.pools = DeviceRayPools([dict(rgbs=rgbs_batch, depth=depth_batch, ray_dirs=ray_dirs_batch, poses=t_wc_batch,
                                          pose_index=frame_batch)], invert=False, device=data_device)

    def get_training_samples(self, n_samples, rand=None, rand_idx=None, exact_rng=False):
        rgbs, depth, origins, dirs, inst = self.pools.next_batch_of(0, n_samples, rand_idx)
        gt_rgb, gt_depth, valid, labels, pcs, z = self.sample_3d_points(rgbs, depth, origins, dirs, rand=rand, exact_rng=exact_rng)
        return gt_rgb, gt_depth, valid, labels, pcs, z, inst          # indices: index 0 repeated (:390)

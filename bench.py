"""bench.py -- train rays/sec (fwd+bwd, device-timed) of the neural-field training hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload replica_room0|c1|scannet0066|c4] [--impl reference]

Contract (see the task statement): one JSON line on rank 0.  A "step" is one pass of the hot path over one
synthetic batch: the sampler's random draws (2 torch launches), depth-guided sampling (3 launches) and the fused tcgen05 step
(prep + instance sort on a side stream, fused fwd+bwd, finalize incl. the loss outputs).  `value` = rays/s with inputs resident in HBM (CUDA events around each step, L2 flushed
between steps, max over ranks); `e2e` = the same through FusedCategoryTrainer.step with pinned HOST buffers
(one H2D copy of the step's inputs + the captured step + D2H of the loss terms, wall clock incl. the stream sync).  `roofline` = the fused kernel against the
measured bf16 tensor peak (MEASURED_PEAKS.json), timed live with CUDA events recorded around that kernel only
(cnf_set_kernel_events).  `cpu_baseline` / `--impl reference` = the reference's OWN modules (staged unmodified under
oracle/_ref by oracle/make_ref.py, driven with functorch.vmap as train.py does) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_POINT = 81888.0          # CodeNeRF fwd+bwd, algorithmic: 3 x 2 x 13648 MAC

WORKLOADS = {
    # name: K categories, instances per category, rays per category, N, M, L, scale
    "replica_room0": dict(K=20, n_obj=3, R=360, N=1, M=9, L=256, scale=2.0,
                          desc="Replica room_0-shaped synthetic: 20 categories x 3 instances, 120 rays/instance, 1+9 samples"),
    "c1": dict(K=1, n_obj=4, R=480, N=5, M=9, L=256, scale=2.0,
               desc="single category (chair), 4 instances x 120 rays, 5+9 samples"),
    "scannet0066": dict(K=8, n_obj=2, R=210, N=1, M=9, L=32, scale=3.0,
                        desc="ScanNet scene0066_00-shaped synthetic: 8 categories, 14 objects, L=32"),
    "c4": dict(K=1, n_obj=64, R=524288, N=1, M=9, L=256, scale=2.0,
               desc="large single category: 64 instances x 8192 rays, 1+9 samples"),
}


def config_of(name, wl):
    """The workload description both arms print (identical keys and values, so the driver can match them)."""
    return {"workload": name, "desc": wl["desc"], "categories_per_gpu": wl["K"], "instances_per_category": wl["n_obj"],
            "rays_per_category": wl["R"], "samples_per_ray": wl["N"] + wl["M"], "latent_dim": wl["L"],
            "l2": "b200 arm: flushed between timed steps (256 MB write)",
            "operands": "b200 arm: fp16 operand images of the parameters are refreshed by the optimizer step (with_optimizer block), not per pass"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def synth_batch(wl, seed, device=None, pinned=False):
    """Synthetic ray batch for K categories (SURVEY 8(d) distributions), generated on the CPU with a fixed seed."""
    from category_nerf_reconstruction_official_b200.synthetic import synth_rays
    g = torch.Generator().manual_seed(seed)
    K, R = wl["K"], wl["R"]
    if R > 100000:      # large batch: generate one category cheaply by tiling a 16k-ray block
        base = synth_rays(16384, wl["n_obj"], g)
        rep = R // 16384
        cats = [tuple(t.repeat(*([rep] + [1] * (t.dim() - 1))) for t in base)]
        cats = cats * K
    else:
        cats = [synth_rays(R, wl["n_obj"], g) for _ in range(K)]
    out = dict(rgbs=torch.stack([c[0] for c in cats]), depth=torch.stack([c[1] for c in cats]),
               origins=torch.stack([c[2] for c in cats]), dirs=torch.stack([c[3] for c in cats]),
               indices=torch.stack([c[4] for c in cats]).to(torch.int32))
    if pinned:
        out = {k: v.contiguous().pin_memory() for k, v in out.items()}
    if device is not None:
        out = {k: v.to(device) for k, v in out.items()}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_arm(wl, steps, warmup, seed=0):
    """Fallback CPU arm when the reference's files are not staged (oracle/_ref absent): the oracle PORT of the reference's
    PyTorch path (same ATen ops, one bmm per layer like functorch.vmap), fp32, all host threads."""
    from oracle import cnf_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(seed)
    K, R, N, M, L, n_obj = wl["K"], wl["R"], wl["N"], wl["M"], wl["L"], wl["n_obj"]
    batch = synth_batch(wl, seed)
    fc = [torch.stack(t) for t in zip(*[O.init_params(O.codenerf_param_shapes(32, L), g) for _ in range(K)])]
    B = O.icosa_matrix().repeat(K, 1, 1)
    sc = [O.init_codes(n_obj, L, g) for _ in range(K)]
    tc = [O.init_codes(n_obj, L, g) for _ in range(K)]
    idx = batch["indices"].long()

    def one_step():
        pcs, zz, valid, labels = [], [], [], []
        for k in range(K):      # the reference samples category by category (train.py:123-125)
            rand = O.draw_sampler_randoms(batch["rgbs"][k], batch["depth"][k], N, M, 0.1)
            _, _, v, lab, p, z = O.sample_points(batch["rgbs"][k], batch["depth"][k], batch["origins"][k], batch["dirs"][k],
                                                 N, M, 0.1, 0.05, rand)
            pcs.append(p); zz.append(z); valid.append(v); labels.append(lab)
        st = torch.stack
        return O.train_step(fc, B, wl["scale"], sc, tc, st(pcs), st(zz), batch["depth"], batch["rgbs"][..., :3], st(labels),
                            st(valid), idx)

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    return K * R / dt, dt, torch.get_num_threads()


def cpu_reference_arm(wl, steps, warmup, seed=0):
    """The reference's OWN CPU implementation of the path on this box's host cores: the unmodified modules staged under
    oracle/_ref (oracle/make_ref.py), driven as train.py:123-182 -- sample_3d_points per category, code gather, functorch
    vmap(PE), vmap(CodeNeRF), step_batch_loss, regulariser, backward; fp32, all host threads, no optimizer step.
    Falls back to the oracle port (kind "port") only when no staged reference exists.  Returns (rays/s, s/step, threads, kind)."""
    from oracle import ref_driver
    if ref_driver.find_ref_src() is not None:
        v, dt, cores = ref_driver.time_reference(wl, synth_batch(wl, seed), steps, warmup, seed)
        return v, dt, cores, "reference"
    v, dt, cores = cpu_port_arm(wl, steps, warmup, seed)
    return v, dt, cores, "port"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="replica_room0", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large", action="store_true", help="skip the extra large-batch (c4) roofline measurement")
    ap.add_argument("--with-optimizer", action="store_true", help="include the fused AdamW update in the timed step")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--categories", type=int, default=0, help="override the workload's categories per GPU (K); SURVEY 8(d): K is not derivable from the reference repo")
    ap.add_argument("--instances", type=int, default=0, help="override the workload's instances per category (R stays 120 rays per instance)")
    ap.add_argument("--ray-split", action="store_true",
                    help="strong scaling: ONE copy of the workload, each rank takes R/N rays + NCCL all-reduce of counts/grads")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.categories or args.instances:
        rpi = wl["R"] // wl["n_obj"]
        wl["K"] = args.categories or wl["K"]
        wl["n_obj"] = args.instances or wl["n_obj"]
        wl["R"] = rpi * wl["n_obj"]
        wl["desc"] += " [overridden: %d categories x %d instances]" % (wl["K"], wl["n_obj"])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    metric = "train rays/sec (fwd+bwd, device-timed)"

    if args.impl == "reference":
        if rank != 0:
            return
        cpu_steps = max(1, args.steps if wl["K"] * wl["R"] <= 20000 else min(args.steps, 3))     # bounded sample for the large batch
        v, dt, cores, kind = cpu_reference_arm(wl, cpu_steps, warmup if wl["K"] * wl["R"] <= 20000 else 1)
        cb = {"value": v, "unit": "rays/s", "cores": cores, "kind": kind,
              "sample": "%d steps of the full %s batch (%d rays/step)" % (cpu_steps, args.workload, wl["K"] * wl["R"])}
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "rays/s", "n_gpus": args.gpus,
                          "steps": cpu_steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config_of(args.workload, wl), "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import category_nerf_reconstruction_official_b200 as cnf
    from category_nerf_reconstruction_official_b200 import _lib
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cnf.load()
    K, R, S = wl["K"], wl["R"], wl["N"] + wl["M"]

    def make_trainer(w):
        t = FusedCategoryTrainer(w["K"], w["n_obj"], latent_dim=w["L"], scale=w["scale"], device=dev,
                                 n_bins_cam2surface=w["N"], n_bins=w["M"],
                                 generator=torch.Generator().manual_seed(1234 + rank))
        # the fp16 operand images / per-instance latent biases are functions of the parameters: refreshed by the optimizer step
        # (inside the timed step with --with-optimizer / the with_optimizer block), not rebuilt by every forward+backward pass
        t.operands_follow_optimizer = True
        t.refresh_operands(w["R"])
        return t

    # weak scaling: every rank owns its own K categories (categories shard across GPUs; no data-path collective)
    if args.ray_split and world > 1:
        ray_split_arm(args, wl, make_trainer, dev, rank, world, warmup, metric)
        dist.destroy_process_group()
        return
    tr = make_trainer(wl)
    batch = synth_batch(wl, seed=100 + rank, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def run_step(t, b, kernel_events=None):
        if kernel_events is not None:
            _lib.check(_lib.load().cnf_set_kernel_events(kernel_events[0].cuda_event, kernel_events[1].cuda_event))
        loss = t.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
        if args.with_optimizer:
            t.adamw_step()
        return loss

    opt_post = (lambda t: (lambda: t.adamw_step())) if args.with_optimizer else (lambda t: None)

    def timed(t, b, steps, warm, graphed=False):
        replay = t.capture(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"], post=opt_post(t)) if graphed else None
        for _ in range(warm):
            replay() if graphed else run_step(t, b)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b_ in kev:       # cuda_event handles exist only after a first record
            a.record(); b_.record()
        torch.cuda.synchronize()
        for i in range(steps):
            flush.zero_()                          # L2 flush between timed iterations (outside the events)
            ev[i][0].record()
            if graphed:
                replay()
            else:
                run_step(t, b, kev[i])
            ev[i][1].record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        step_ms = sum(a.elapsed_time(b_) for a, b_ in ev) / steps
        kern_ms = sum(a.elapsed_time(b_) for a, b_ in kev) / steps
        return step_ms, kern_ms

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    use_graph = not args.no_graph
    # pass 1 (eager launches): times the fused kernel alone via events recorded around it inside cnf_train_step
    eager_ms, kern_ms = timed(tr, batch, args.steps, warmup)
    # pass 2: the step as a user runs it in a training loop -- one CUDA-graph replay per step
    step_ms = timed(tr, batch, args.steps, warmup, graphed=True)[0] if use_graph else eager_ms
    clk = clocks.stop() if rank == 0 else None
    loss_val = float(tr.out_loss[0])

    # ---- e2e through the public API with HOST (pinned) buffers: fused.HostStepPipeline -- every step's inputs are copied from pinned
    # host memory (one H2D per step), the step runs as one CUDA-graph replay and its loss terms are copied back to the host and READ
    # by the loop; two slots keep one step in flight, so the copies of step i+1 overlap the kernels of step i (a training loop that
    # logs the loss one step late).  All copies and reads are inside the timed region.
    from category_nerf_reconstruction_official_b200.fused import HostStepPipeline
    hb_src = synth_batch(wl, seed=100 + rank)
    pipe = HostStepPipeline(tr, hb_src, post=opt_post(tr))
    h2d = pipe.bytes_per_step
    d2h = pipe.out_host[0].numel() * 4

    def e2e_loop(n, with_flush):
        seen = 0.0
        for i in range(n):
            if with_flush:
                flush.zero_()                                              # L2 flush between steps (its time is measured and removed below)
            pipe.submit(i & 1)
            if i:
                seen += float(pipe.result()[0])                           # the host reads step i-1's loss
        seen += float(pipe.result()[0])
        return seen

    e2e_loop(warmup, True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps, True)
    torch.cuda.synchronize()
    e2e_s_total = time.perf_counter() - t0
    # the flush is not part of what a user runs: time the same number of flushes alone and remove them
    t1 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
    torch.cuda.synchronize()
    flush_s = time.perf_counter() - t1
    e2e_ms = max(1e-6, (e2e_s_total - flush_s)) / args.steps * 1e3
    # the same loop without overlap (submit, then wait for the result) for comparison
    t2 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        pipe.submit(i & 1)
        pipe.result()
    torch.cuda.synchronize()
    e2e_serial_ms = max(1e-6, (time.perf_counter() - t2 - flush_s)) / args.steps * 1e3

    # ---- max over ranks
    if dist is not None:
        tt = torch.tensor([step_ms, kern_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms, kern_ms, e2e_ms = [float(x) for x in tt]

    rays_per_step = K * R * world
    value = rays_per_step / (step_ms * 1e-3)
    pk = peaks()
    pts = K * R * S
    flops = pts * FLOP_PER_POINT
    ach = flops / (kern_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None    # DRAM bytes of the fused kernel per launch, from the committed ncu capture of this workload
    for tname in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath) and not (args.categories or args.instances):
            traffic = json.load(open(tpath)).get(args.workload, {}).get("bytes_per_launch")
            traffic_src = "profiles/%s (ncu --set full capture of this workload; bench.py cannot run ncu itself)" % tname
            break
    roof = {"bound": "tensor", "kernel": "fused_step_kernel", "achieved": ach, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
            "frac": ach / pk["bf16_burst"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["source"] + " (bf16 burst)",
            "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / step_ms,
            "algorithmic_flop_per_point": FLOP_PER_POINT, "points_per_launch": pts}

    extra = {}
    if rank == 0 and not args.no_large and args.workload != "c4":
        # the large-batch regime (BASELINE config 4 shape): where the tensor-pipe fraction is meaningful
        wl4 = dict(WORKLOADS["c4"])
        try:
            tr4 = make_trainer(wl4)
            b4 = synth_batch(wl4, seed=7, device=dev)
            s4, k4 = timed_single(tr4, b4, run_step, flush, steps=5, warm=3)
            p4 = wl4["K"] * wl4["R"] * (wl4["N"] + wl4["M"])
            a4 = p4 * FLOP_PER_POINT / (k4 * 1e-3) / 1e12
            extra["large_batch_c4"] = {"rays_per_s": wl4["K"] * wl4["R"] / (s4 * 1e-3), "ms_per_step": s4, "kernel_ms": k4,
                                       "tensor_tflops": a4, "tensor_frac_of_burst_peak": a4 / pk["bf16_burst"],
                                       "config": wl4["desc"]}
            del tr4, b4
        except Exception as e:      # never let the extra measurement break the contract line
            extra["large_batch_c4"] = {"error": repr(e)}

    if rank == 0 and not args.no_large:
        # inference sweep (BASELINE config 5 / Trainer.eval_points): forward-only tensor-core kernel on a 256^3 grid
        try:
            from category_nerf_reconstruction_official_b200 import ops as _ops
            P = 256 ** 3
            pts = (torch.rand(P, 3, device=dev) - 0.5) * 3.0
            fl = tr.params[0]; sc0 = tr.shape_codes[0, 0]; tc0 = tr.tex_codes[0, 0]
            for _ in range(2):
                _ops.eval_points_fused(pts, fl, tr.pe_B[0], tr.pe_scale[0], sc0, tc0)
            torch.cuda.synchronize()
            tms = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); _ops.eval_points_fused(pts, fl, tr.pe_B[0], tr.pe_scale[0], sc0, tc0); e1.record()
                torch.cuda.synchronize(); tms.append(e0.elapsed_time(e1))
            ms5 = sorted(tms)[2]
            extra["eval_points_256cube"] = {"points": P, "ms": ms5, "points_per_s": P / (ms5 * 1e-3),
                                            "tensor_tflops": P * 27296.0 / (ms5 * 1e-3) / 1e12,
                                            "tensor_frac_of_burst_peak": P * 27296.0 / (ms5 * 1e-3) / 1e12 / pk["bf16_burst"],
                                            "note": "forward-only CodeNeRF (27296 FLOP/point) incl. the operand-image prep launch; inputs larger than L2"}
            del pts
        except Exception as e:
            extra["eval_points_256cube"] = {"error": repr(e)}

    if rank == 0 and not args.no_large:
        # background model inference (Trainer.eval_points for cls 0): OccupancyMap W=128 forward on tensor cores, 256^3 grid
        try:
            from category_nerf_reconstruction_official_b200 import ops as _ops, model as _model
            mb = _model.OccupancyMap(87, 42, hidden_size=128).to(dev); mb.apply(_model.init_weights)
            flb = _model.flat_params(mb).detach()
            P = 256 ** 3
            pts = (torch.rand(P, 3, device=dev) - 0.5) * 8.0
            for _ in range(2):
                _ops.occmap_eval_points_fused(pts, flb, tr.pe_B[0], tr.pe_scale[0])
            torch.cuda.synchronize()
            tms = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); _ops.occmap_eval_points_fused(pts, flb, tr.pe_B[0], tr.pe_scale[0]); e1.record()
                torch.cuda.synchronize(); tms.append(e0.elapsed_time(e1))
            ms6 = sorted(tms)[2]
            tf6 = P * 187392.0 / (ms6 * 1e-3) / 1e12
            extra["occmap_eval_points_256cube"] = {"points": P, "ms": ms6, "points_per_s": P / (ms6 * 1e-3), "tensor_tflops": tf6,
                                                   "tensor_frac_of_burst_peak": tf6 / pk["bf16_burst"],
                                                   "note": "forward-only OccupancyMap W=128 (187392 FLOP/point), embedding in-kernel, two CTAs per SM"}
            del pts, mb
        except Exception as e:
            extra["occmap_eval_points_256cube"] = {"error": repr(e)}

    if rank == 0 and not args.no_large:
        # dense view rendering (BASELINE config 5): 512x512 rays of one instance, forward only, S = 14 and S = 96
        try:
            from category_nerf_reconstruction_official_b200 import ops as _ops
            Rv = 512 * 512
            fl = tr.params[0]; sc0 = tr.shape_codes[0, 0]; tc0 = tr.tex_codes[0, 0]
            org = (torch.rand(Rv, 3, device=dev) - 0.5) * 2.0
            dr = torch.nn.functional.normalize(torch.randn(Rv, 3, device=dev), dim=-1)
            for Sv in (14, 96):
                zv = torch.sort(torch.rand(Rv, Sv, device=dev) * 2.5, dim=-1).values
                for _ in range(2):
                    _ops.render_rays_fused(org, dr, zv, fl, tr.pe_B[0], tr.pe_scale[0], sc0, tc0)
                torch.cuda.synchronize()
                tms = []
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); _ops.render_rays_fused(org, dr, zv, fl, tr.pe_B[0], tr.pe_scale[0], sc0, tc0); e1.record()
                    torch.cuda.synchronize(); tms.append(e0.elapsed_time(e1))
                msv = sorted(tms)[2]
                tf = Rv * Sv * 27296.0 / (msv * 1e-3) / 1e12
                extra["render_512x512_S%d" % Sv] = {"rays": Rv, "samples_per_ray": Sv, "ms": msv, "rays_per_s": Rv / (msv * 1e-3),
                                                    "points_per_s": Rv * Sv / (msv * 1e-3), "tensor_tflops": tf,
                                                    "tensor_frac_of_burst_peak": tf / pk["bf16_burst"]}
                del zv
            del org, dr
        except Exception as e:
            extra["render_512x512"] = {"error": repr(e)}

    if rank == 0 and not args.no_large:
        # background model step (train.py:172-180): 1200 rays x (5+9) samples, OccupancyMap W=128 on tensor cores, CUDA-graph replay
        try:
            from category_nerf_reconstruction_official_b200.background import FusedBackgroundTrainer
            from category_nerf_reconstruction_official_b200.synthetic import synth_rays as _sr
            gb = torch.Generator().manual_seed(11)
            trb = FusedBackgroundTrainer(scale=5.0, device=dev, generator=gb)
            rb = [t.to(dev) for t in _sr(1200, 1, gb)][:4]
            replay_b = trb.capture(*rb)
            for _ in range(3):
                replay_b()
            torch.cuda.synchronize()
            tms = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); replay_b(); e1.record(); torch.cuda.synchronize(); tms.append(e0.elapsed_time(e1))
            msb = sorted(tms)[5]
            extra["background_step_1200x14"] = {"rays": 1200, "samples_per_ray": 14, "ms_per_step": msb, "rays_per_s": 1200 / (msb * 1e-3),
                                                "algorithmic_flop_per_point": 562176.0,
                                                "tensor_tflops_whole_step": 16800 * 562176.0 / (msb * 1e-3) / 1e12,
                                                "note": "sampler + embed + OccupancyMap(128) fwd/bwd on tcgen05 + composite/loss, graph replay, L2 flushed"}
            # whole iteration of the reference: all object categories + the background model, one graph, two parallel branches
            from category_nerf_reconstruction_official_b200.background import capture_iteration
            tri = make_trainer(wl)
            bi = synth_batch(wl, seed=5, device=dev)
            replay_i = capture_iteration(tri, (bi["rgbs"], bi["depth"], bi["origins"], bi["dirs"], bi["indices"]), trb, tuple(rb))
            for _ in range(3):
                replay_i()
            torch.cuda.synchronize()
            tms = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); replay_i(); e1.record(); torch.cuda.synchronize(); tms.append(e0.elapsed_time(e1))
            msi = sorted(tms)[5]
            extra["iteration_objects_plus_background"] = {"rays": K * R + 1200, "ms_per_iteration": msi, "rays_per_s": (K * R + 1200) / (msi * 1e-3),
                                                          "note": "C2 object step and background step as parallel branches of one CUDA graph (train.py:123-184 minus the optimizer)"}
            del trb, rb, tri, bi
        except Exception as e:
            extra["background_step_1200x14"] = {"error": repr(e)}

    if rank == 0 and not args.no_large and not args.with_optimizer:
        # the same step with the fused AdamW update of all four parameter groups inside the captured graph (report
        # with / without the optimizer; train.py:183)
        try:
            tro = make_trainer(wl)
            bo = synth_batch(wl, seed=100 + rank, device=dev)
            replay_o = tro.capture(bo["rgbs"], bo["depth"], bo["origins"], bo["dirs"], bo["indices"], post=lambda: tro.adamw_step())
            for _ in range(warmup):
                replay_o()
            torch.cuda.synchronize()
            tms = []
            for _ in range(args.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); replay_o(); e1.record(); torch.cuda.synchronize(); tms.append(e0.elapsed_time(e1))
            mso = sum(tms) / len(tms)
            extra["with_optimizer"] = {"ms_per_step": mso, "rays_per_s": K * R / (mso * 1e-3), "optimizer_steps_taken": tro.opt_steps,
                                       "note": "step + cnf_adamw_groups (one launch, device-side step count) in one CUDA graph"}
            del tro, bo
        except Exception as e:
            extra["with_optimizer"] = {"error": repr(e)}

    if rank == 0 and not args.no_large and world == 1:
        # HBM-bound standalone kernels of the module-level API against the measured copy bandwidth (SURVEY 8(d) algorithmic bytes;
        # ncu DRAM counters of the same kernels: profiles/r02_hbm_ncu.txt)
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("bench_standalone", os.path.join(ROOT, "scripts", "bench_standalone.py"))
            mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
            hk = mod.measure(dev, 1 << 20, 5, flush)
            extra["hbm_kernels"] = {k_: ({"bytes": v["alg_bytes"], "ms": v["ms"], "gbs": v["gbs"], "frac": v["frac"]} if isinstance(v, dict) else v)
                                    for k_, v in hk.items()}
        except Exception as e:
            extra["hbm_kernels"] = {"error": repr(e)}

    if world > 1:
        # (collective blocks: every rank takes part)
        try:
            extra["ray_split_c4"] = ray_split_c4_block(make_trainer, dev, rank, world, flush, pk)
        except Exception as e:
            extra["ray_split_c4"] = {"error": repr(e)}
        try:
            extra["scannet0066_sharded"] = category_sharded_block(dev, rank, world, flush)
        except Exception as e:
            extra["scannet0066_sharded"] = {"error": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cb = None
    if not args.no_cpu_baseline and world == 1:
        cpu_steps = 10 if K * R <= 20000 else 1
        v, dt, cores, kind = cpu_reference_arm(wl, cpu_steps, 2)
        cb = {"value": v, "unit": "rays/s", "cores": cores, "kind": kind, "ms_per_step": dt * 1e3,
              "sample": "%d steps of the full %s batch (%d rays/step), %s" % (
                  cpu_steps, args.workload, K * R,
                  "unmodified reference modules (oracle/_ref) under functorch.vmap" if kind == "reference"
                  else "oracle port of the reference PyTorch CPU path")}

    line = {"metric": metric, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
            "data": "synthetic",
            "config": config_of(args.workload, wl),
            "run": {"optimizer_in_step": bool(args.with_optimizer),
                    "cuda_graph": bool(use_graph), "eager_ms_per_step": eager_ms,
                    "parallelism": "categories sharded, %d per GPU, no collective in the step" % K},
            "clocks": clk, "roofline": roof, "cpu_baseline": cb,
            "e2e": {"value": rays_per_step / (e2e_ms * 1e-3), "unit": "rays/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "fused.HostStepPipeline: per step one H2D copy from pinned host memory, one graph replay, one D2H copy of "
                           "the loss terms read by the host; one step in flight (copies overlap the previous step)",
                    "ms_per_step_no_overlap": e2e_serial_ms},
            "gpu_launches": (tr.launches_per_step + (1 if args.with_optimizer else 0)) * args.steps,
            "loss": loss_val}
    line.update(extra)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def ray_split_arm(args, wl, make_trainer, dev, rank, world, warmup, metric):
    """Strong scaling of one workload: rays split across ranks, identical model replicas, NCCL all-reduces."""
    import torch.distributed as dist
    from category_nerf_reconstruction_official_b200.parallel import RaySplitStep, split_rays
    gen_seed = 1234            # same replica on every rank
    tr = make_trainer(wl)
    for t_ in (tr.params, tr.shape_codes, tr.tex_codes):
        dist.broadcast(t_, src=0)
    full = synth_batch(wl, seed=100, device=None)
    sl = split_rays(wl["R"], world, rank)
    b = {k: v[:, sl].contiguous().to(dev) for k, v in full.items()}
    rs = RaySplitStep(tr)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        rs.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        rs.step(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
        ev[i][1].record()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([sum(a.elapsed_time(c) for a, c in ev) / args.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": metric, "value": wl["K"] * wl["R"] / (float(ms) * 1e-3), "unit": "rays/s", "n_gpus": world,
                          "steps": args.steps, "warmup": warmup, "ms_per_step": float(ms), "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": args.workload, "desc": wl["desc"], "parallelism": "ray split %d-way + NCCL all-reduce" % world,
                                     "l2": "flushed between timed steps"}, "loss": float(tr.out_loss[0]),
                          "gpu_launches": 8 * args.steps}))


def _event_ms(fn, flush, steps, warm, dist=None, dev=None):
    """Mean device time of fn() over `steps` calls (L2 flushed before each), max over ranks when `dist` is given."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def ray_split_c4_block(make_trainer, dev, rank, world, flush, pk, steps=10, warm=3):
    """BASELINE config 4 on `world` GPUs: ONE large category (64 instances x 8192 rays, R = 524288) with its ray batch split over the
    ranks; identical model replicas; one all-gather of the mask counts / max depth before the sampler and ONE NCCL all-reduce of the
    flat gradient buffer after the fused step (parallel.RaySplitStep), the whole step captured in a CUDA graph.
    Also the NCCL parity check a 1-GPU test box cannot run: the N-rank reduced gradients of one step against rank 0's own
    full-batch gradients on the same samples."""
    import torch.distributed as dist
    from category_nerf_reconstruction_official_b200 import ops
    from category_nerf_reconstruction_official_b200.parallel import RaySplitStep, split_rays
    wl = dict(WORKLOADS["c4"])
    K, R, N, M = wl["K"], wl["R"], wl["N"], wl["M"]
    tr = make_trainer(wl)
    for t_ in (tr.params, tr.pe_B, tr.shape_codes, tr.tex_codes):             # identical replicas
        dist.broadcast(t_, src=0)
    full = synth_batch(wl, seed=100, device=dev)                              # same seed on every rank: same full batch
    sl = split_rays(R, world, rank)
    loc = {k_: v[:, sl].contiguous() for k_, v in full.items()}
    rs = RaySplitStep(tr)
    # ---- parity: samples of the FULL batch (bit-exact sampler, same seeded offsets on every rank), each rank steps its slice
    g = torch.Generator(device=dev).manual_seed(77)
    rand = {"u_inv": torch.rand(K, R, N + M, device=dev, generator=g), "u_val": torch.rand(K, R, N, device=dev, generator=g),
            "n_obj": torch.randn(K, R, M, device=dev, generator=g) * (0.1 / 3), "u_oth": torch.rand(K, R, M, device=dev, generator=g)}
    z, _, valid, labels, stats = ops.sample_batched(full["rgbs"], full["depth"], full["origins"], full["dirs"], N, M, 0.1, 0.05, rand,
                                                     want_pcs=False)
    counts_full = stats[:, 4:8].contiguous()
    vm = valid.view(torch.uint8)
    local_counts = ops.mask_counts(loc["rgbs"][..., 3].contiguous(), vm[:, sl].contiguous())
    counts_g, _ = rs.global_counts_and_max(local_counts, loc["depth"].max(dim=-1).values)
    counts_ok = bool(torch.equal(counts_g[:, :3], counts_full[:, :3]))
    # the path RaySplitStep.step takes: the sampler's count pass on the local rays, all-gather of the stats rows, cnf_reduce_stats
    counts_s, max_s = rs.global_stats(ops.sample_stats(loc["rgbs"], loc["depth"]))
    counts_ok = counts_ok and bool(torch.equal(counts_s[:, :3], counts_full[:, :3])) and \
        bool(torch.equal(max_s, full["depth"].max(dim=-1).values))
    tr.prepare(loc["indices"], loc["indices"].shape[1])
    rs.step_from_samples(loc["origins"], loc["dirs"], z[:, sl].contiguous(), loc["depth"], loc["rgbs"], vm[:, sl].contiguous(),
                         loc["indices"], counts_g, prepared=True)
    reduced = tr.grad_flat.clone()
    reduced_loss = float(tr.out_loss[0])
    err = None
    if rank == 0:
        tr.step_from_samples(full["origins"], full["dirs"], z, full["depth"], full["rgbs"], vm, full["indices"], counts_full)
        err = float((reduced.double() - tr.grad_flat.double()).norm() / tr.grad_flat.double().norm())
        loss_err = abs(reduced_loss - float(tr.out_loss[0])) / abs(float(tr.out_loss[0]))
    del z, valid, labels, rand, vm
    torch.cuda.synchronize(); dist.barrier()
    # ---- timing: the ray-split step as a graph replay (sampler with locally drawn offsets + both collectives inside)
    try:
        replay = rs.capture(loc["rgbs"], loc["depth"], loc["origins"], loc["dirs"], loc["indices"])
        graphed = True
    except Exception:
        replay = lambda: rs.step(loc["rgbs"], loc["depth"], loc["origins"], loc["dirs"], loc["indices"])
        graphed = False
    ms = _event_ms(replay, flush, steps, warm, dist, dev)
    ar_ms = _event_ms(rs.allreduce_outputs, flush, steps, warm, dist, dev)     # the gradient all-reduce alone (eager launch)
    # ---- 1-GPU baseline of the same workload on rank 0 (the others wait): strong-scaling efficiency
    t1 = None
    if rank == 0:
        rep1 = tr.capture(full["rgbs"], full["depth"], full["origins"], full["dirs"], full["indices"])
        t1 = _event_ms(rep1, flush, 5, 2)
    torch.cuda.synchronize(); dist.barrier()
    if rank != 0:
        return None
    return {"config": wl["desc"] + "; rays split %d-way" % world, "rays": K * R, "ms_per_step": ms, "rays_per_s": K * R / (ms * 1e-3),
            "one_gpu_ms_per_step": t1, "strong_efficiency": t1 / (world * ms), "allreduce_ms": ar_ms, "allreduce_share": ar_ms / ms,
            "allreduce_bytes": tr.grad_flat.numel() * 4, "cuda_graph": graphed, "collectives_per_step": 2,
            "grad_rel_err_vs_1gpu": err, "loss_rel_err_vs_1gpu": loss_err, "global_counts_match": counts_ok}


def category_sharded_block(dev, rank, world, flush, steps=20, warm=3):
    """BASELINE config 3 (ScanNet scene0066_00-shaped: 8 categories, L = 32, partial-observation depth masks) sharded by category:
    parallel.shard_categories assigns the 8 categories to the ranks, every rank steps its own, no collective in the step."""
    import torch.distributed as dist
    from category_nerf_reconstruction_official_b200.fused import FusedCategoryTrainer
    from category_nerf_reconstruction_official_b200.parallel import shard_categories
    wl = dict(WORKLOADS["scannet0066"])
    mine = shard_categories([float(wl["R"] * (wl["N"] + wl["M"]))] * wl["K"], world)[rank]
    ms = 0.0
    if mine:
        w = dict(wl); w["K"] = len(mine)
        tr = FusedCategoryTrainer(w["K"], w["n_obj"], latent_dim=w["L"], scale=w["scale"], device=dev, n_bins_cam2surface=w["N"],
                                  n_bins=w["M"], generator=torch.Generator().manual_seed(77 + rank))
        b = synth_batch(w, seed=300 + rank, device=dev)
        replay = tr.capture(b["rgbs"], b["depth"], b["origins"], b["dirs"], b["indices"])
    else:
        replay = lambda: None
    ms = _event_ms(replay, flush, steps, warm, dist, dev)
    if rank != 0:
        return None
    return {"config": wl["desc"], "categories_per_rank": [len(x) for x in shard_categories([1.0] * wl["K"], world)],
            "ms_per_step": ms, "rays_per_s": wl["K"] * wl["R"] / (ms * 1e-3), "collectives_per_step": 0}


def timed_single(t, b, run_step, flush, steps, warm):
    from category_nerf_reconstruction_official_b200 import _lib
    for _ in range(warm):
        run_step(t, b)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b_ in kev:
        a.record(); b_.record()
    torch.cuda.synchronize()
    for i in range(steps):
        flush.zero_()
        ev[i][0].record()
        run_step(t, b, kev[i])
        ev[i][1].record()
    torch.cuda.synchronize()
    return (sum(a.elapsed_time(b_) for a, b_ in ev) / steps, sum(a.elapsed_time(b_) for a, b_ in kev) / steps)


if __name__ == "__main__":
    main()

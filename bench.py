#!/usr/bin/env python3
"""Benchmark of the NR-NeRF render hot path on B200 (driver contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...   # reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): one training step of configs/example_sequence.txt --
N_rand = 1024 rays per GPU, 64 coarse + 128 fine network evaluations per ray, 8x256 MLP, ray bending on,
perturb = 1, raw_noise_std = 1, offsets / rigidity / divergence regularisers on, backward, Adam --
on synthetic rays shaped like the example sequence (there is no dataset on the box).
Metric: rays/sec (whole job, all ranks).  Weak scaling: the per-GPU ray batch is fixed.
"""
import argparse
import json
import os
import subprocess
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rays/sec (64c+128f samples, 8x256 MLP), example_sequence training step"
N_RAND = 1024
N_SAMPLES, N_IMPORTANCE = 64, 64
FLOP_PER_POINT = 1_016_320          # forward, per point evaluation (NeRF + bender)
POINTS_PER_RAY = N_SAMPLES + (N_SAMPLES + N_IMPORTANCE)


def make_args():
    a = types.SimpleNamespace()
    a.chunk, a.N_samples, a.N_importance, a.N_iters = 32768, N_SAMPLES, N_IMPORTANCE, 200000
    a.offsets_loss_weight, a.divergence_loss_weight, a.rigidity_loss_weight = 60.0, 3.0, 0.0005
    a.ray_bending_latent_size = 32
    return a


def synth_batch(rs, n, n_images=86):
    """Rays like get_rays_np on the 384x512 example frames + uniform targets + (image, y, x) indices."""
    H, W, focal = 384, 512, 256.61
    img = rs.randint(0, n_images, size=n)
    y, x = rs.randint(0, H, size=n), rs.randint(0, W, size=n)
    dirs = np.stack([(x - W * 0.5) / focal, -(y - H * 0.5) / focal, -np.ones(n)], -1).astype(np.float32)
    ang = (img.astype(np.float32) / n_images - 0.5) * 0.6
    c, s = np.cos(ang), np.sin(ang)
    rays_d = np.stack([c * dirs[:, 0] + s * dirs[:, 2], dirs[:, 1], -s * dirs[:, 0] + c * dirs[:, 2]], -1).astype(np.float32)
    rays_o = np.stack([0.3 * s, np.zeros(n), 0.4 + 0.0 * s], -1).astype(np.float32)
    target = rs.uniform(0, 1, size=(n, 3)).astype(np.float32)
    idx = np.stack([img, y, x], -1).astype(np.int64)
    return rays_o, rays_d, target, idx


def read_peaks():
    """(sustained bf16 TFLOP/s, HBM GB/s, source).  Sustained figures: the kernels are timed inside a long step."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), float(d.get("hbm_gbs", 6500.0)),
                "measured (MEASURED_PEAKS.json: sustained bf16 cuBLAS, copy bandwidth)")
    return 1400.0, 6500.0, "fallback (B200_PROFILING.md figures)"


class ClockSampler:
    def __init__(self, dev_index):
        self.path = f"/tmp/nrn_clocks_{os.getpid()}.csv"
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", f"--id={dev_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.remove(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (a PyTorch-on-CPU restatement of the reference algorithm; /root/reference does not
# exist on the GPU box) -- the only place bench.py executes anything under oracle/
# ---------------------------------------------------------------------------------------------
def cpu_training_rate(n_rays, steps, warmup, threads):
    import oracle.nrnerf_oracle as O
    torch.set_num_threads(threads)
    cp, fp, bp = O.clone_params(O.make_nerf_params(1, 5, 30.0), True), O.clone_params(O.make_nerf_params(2, 5, 30.0), True), \
        O.clone_params(O.make_bender_params(3), True)
    params = O.flat_param_list(cp) + O.flat_param_list(fp) + O.flat_param_list(bp)
    opt = torch.optim.Adam(params, lr=5e-4, betas=(0.9, 0.999))
    times = []
    for it in range(warmup + steps):
        r = O.make_rays(100 + it, n_rays)
        rnd = O.make_randomness(100 + it, n_rays, N_SAMPLES, N_IMPORTANCE)
        lat = r["latents"].clone().requires_grad_(True)
        t0 = time.perf_counter()
        ret = O.render_rays(cp, fp, bp, r["rays_o"], r["rays_d"], r["near"], r["far"], lat, N_SAMPLES, N_IMPORTANCE, perturb=True,
                            raw_noise_std=1.0, rnd=rnd)
        loss = O.training_loss(ret, r["target"], 60.0, 0.0005, 0.01)
        loss = loss + 3.0 * 0.01 * O.divergence_loss(bp, ret, lat, n_rays, N_SAMPLES)
        opt.zero_grad()
        loss.mean().backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return n_rays / float(np.median(times)), float(np.median(times))


def run_reference_arm(args, rank):
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 32)   # see cpu_baseline: more threads make the small-tensor ops slower
    sample = 256
    rate, sec = cpu_training_rate(sample, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "example_sequence training step (N_rand=1024/GPU, 64c+128f, ray bending on, regularisers, Adam)",
                       "note": "reference algorithm on the host CPU cores via the oracle port (the reference checkout is not on the box)"},
            "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port",
                             "sample": f"{sample}-ray slices of the 1024-ray step, median of {args.steps} steps"},
            "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n-rand", type=int, default=N_RAND, help="rays per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    from nonrigid_nerf_b200 import _lib, parallel, run_nerf_helpers as H

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    # ---- models exactly as create_nerf builds them (train.py:556-721), default inits, bender output layers re-drawn
    torch.manual_seed(0)
    embed_fn, input_ch = H.get_embedder(10, 0)
    bender = H.ray_bending(input_ch, 32, "simple_neural", embed_fn).to(dev)
    with torch.no_grad():
        bender.network[-1].weight.normal_(0, 0.01)
        bender.rigidity_network[-1].weight.normal_(0, 0.1)
    kw = dict(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=0, use_viewdirs=False, ray_bender=bender,
              ray_bending_latent_size=32)
    coarse = H.NeRF(num_ray_samples=N_SAMPLES, **kw).to(dev)
    fine = H.NeRF(num_ray_samples=N_SAMPLES + N_IMPORTANCE, **kw).to(dev)
    n_images = 86
    latents = [torch.zeros(32, device=dev).normal_(0, 0.1).requires_grad_(True) for _ in range(n_images)]
    grad_vars = latents + list(bender.parameters()) + list(coarse.parameters()) + list(fine.parameters())
    from nonrigid_nerf_b200 import optim
    optimizer = optim.Adam(grad_vars, lr=5e-4, betas=(0.9, 0.999))   # train.py:656-658; one launch over a flat parameter buffer
    render_kwargs_train = {"network_query_fn": None, "perturb": 1.0, "N_importance": N_IMPORTANCE, "network_fine": fine,
                           "N_samples": N_SAMPLES, "network_fn": coarse, "ray_bender": bender, "use_viewdirs": False,
                           "white_bkgd": False, "raw_noise_std": 1.0, "ndc": False, "lindisp": False, "near": 0.0022, "far": 1.0024}
    targs = make_args()
    dataset_extras = {"imageid_to_timestepid": list(range(n_images))}
    train_fn = parallel.get_parallelized_training_function(coarse, latents, fine_model=fine, ray_bender=bender)

    n_global = args.n_rand * world
    rs = np.random.RandomState(1234)     # identical on every rank: same global batch, sliced by rank inside train_fn
    pool = 8
    host = [synth_batch(rs, n_global, n_images) for _ in range(pool)]
    pinned = [[torch.from_numpy(a).pin_memory() for a in b] for b in host]

    def to_device_packed(batch):
        """One flat device buffer per batch; the four tensors are views into it, so that the multi-GPU input
        broadcast is ONE collective."""
        sizes = [t.numel() * t.element_size() for t in batch]
        offs = [0]
        for sz in sizes:
            offs.append((offs[-1] + sz + 15) // 16 * 16)
        flat = torch.empty(offs[-1], dtype=torch.uint8, device=dev)
        views = []
        for t, o, sz in zip(batch, offs, sizes):
            v = flat[o:o + sz].view(t.dtype).view(t.shape)
            v.copy_(t, non_blocking=True)
            views.append(v)
        return flat, views

    packed = [to_device_packed(b) for b in pinned]
    resident = [v for _, v in packed]

    def eager_step(rays_o, rays_d, target, idx):
        losses = train_fn(targs, rays_o, rays_d, 100, render_kwargs_train, target, 1000, 0, dataset_extras, idx)
        loss = torch.mean(losses)
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()          # world > 1: the pre-step hook all-reduces the gradients (one flat NCCL call)
        return loss

    # CUDA-graph replay.  1 GPU: the whole iteration (forward, losses, backward, Adam) is one graph.
    # N GPUs: the graph holds this rank's forward + backward only; the collectives (input broadcast, ONE flat
    # gradient all-reduce, loss all-gather) and Adam run eagerly around it -- NCCL calls are not captured.
    graphed = None
    lo, hi = parallel.shard_bounds(n_global, world, rank)
    if not args.no_graph:
        try:
            from nonrigid_nerf_b200.graphs import GraphedStep
            if world == 1:
                graphed = GraphedStep(eager_step, resident[0], warmup=3)
            else:
                parallel.UNIFORM_GRADS = True
                local_module = train_fn.module   # training_wrapper_class: the per-rank step DataParallel used to wrap

                def local_fwd_bwd(rays_o, rays_d, target, idx):
                    losses = local_module(targs, rays_o, rays_d, 100, render_kwargs_train, target, 1000, 0, dataset_extras, idx)
                    optimizer.zero_grad(set_to_none=True)
                    (losses.sum() / n_global).backward()      # the caller's mean over the GLOBAL batch (train.py:1606-1607)
                    return losses

                graphed = GraphedStep(local_fwd_bwd, [t[lo:hi].contiguous() for t in resident[0]], warmup=3)
        except Exception as exc:  # noqa: BLE001 - fall back to the eager loop, and say so in the JSON line
            print(f"[bench] CUDA graph capture failed ({type(exc).__name__}: {exc}); running eagerly", file=sys.stderr)
            graphed = None
            torch.cuda.synchronize()

    lrate, lrate_decay, iteration = 5e-4, 250, [0]

    def step(i, batch, batch_flat=None):
        iteration[0] += 1
        optimizer.set_lr(lrate * (0.1 ** (iteration[0] / (lrate_decay * 1000))))   # per-iteration decay of train.py:1611-1616
        if graphed is None:
            return eager_step(*batch)
        if world == 1:
            return graphed(*batch)
        if batch_flat is None:
            batch_flat = packed[i % pool][0]
        dist.broadcast(batch_flat, src=0)     # every rank uses rank 0's batch (the reference samples with unseeded numpy)
        losses_local = graphed(*[t[lo:hi] for t in batch])
        optimizer.step()                      # pre-step hook: one flat gradient all-reduce, then Adam on every rank
        gathered = torch.empty(world * (hi - lo), dtype=losses_local.dtype, device=dev)
        dist.all_gather_into_tensor(gathered, losses_local.contiguous())   # per-ray losses [N_rand] on every rank
        return gathered.mean()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(loop_steps, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(loop_steps):
            if e2e:
                flat, b = to_device_packed(pinned[i % pool])   # host -> device copy of this step's inputs (pinned memory)
                loss = step(i, b, flat)
                loss.item()                      # device -> host read of the step's result
            else:
                step(i, resident[i % pool])
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        step(i, resident[i % pool])
    _lib.device_error_check()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # headline numbers: the plain step (no instrumentation inside the graph)
    ms_total = timed(args.steps, e2e=False)
    ms_e2e = timed(args.steps, e2e=True)
    # per-kernel breakdown: the same K steps once more with CUDA event records around every launch of this repo's
    # kernels (external event-record nodes inside the re-captured graph; they cost a few us per step themselves)
    _lib.timing_enable(True)
    if graphed is not None:
        if world == 1:
            graphed = GraphedStep(eager_step, resident[0], warmup=1)
        else:
            graphed = GraphedStep(local_fwd_bwd, [t[lo:hi].contiguous() for t in resident[0]], warmup=1)
        for i in range(3):
            step(i, resident[i % pool])
    ms_instrumented = timed(args.steps, e2e=False)
    kinds = _lib.timing_read()
    _lib.timing_enable(False)
    clocks = sampler.stop() if sampler else None
    _lib.device_error_check()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = n_global * args.steps / (ms_total * 1e-3)
    e2e_value = n_global * args.steps / (ms_e2e * 1e-3)
    peak_tf, peak_gbs, peak_src = read_peaks()
    # Per-step time of each kernel kind from the instrumented pass (this rank).  Launches per step: field forward /
    # DGRAD / composite / composite backward 2 (coarse + fine pass), divergence 2 (forward + backward), WGRAD 3 (fine,
    # coarse, divergence; each followed by its split reduction, timed with it).  Eager: the events of every step
    # were recorded; graph replay: the captured event pairs hold the timestamps of the last replay.
    launches = {"wgrad": 3}
    per_step = {k: (kinds[k][0] / (kinds[k][1] / float(launches.get(k, 2))) if kinds[k][1] else 0.0) for k in kinds}
    dom = max(("field_fwd", "field_dgrad", "wgrad"), key=lambda k: per_step[k])
    dom_ms_per_step = per_step[dom]
    tiles = args.n_rand * POINTS_PER_RAY // 128
    if dom == "wgrad":
        # HBM-bound by construction: every stash byte is read once and feeds 256 MACs.
        alg = tiles * (634880 + 618496) + (args.n_rand * N_SAMPLES // 128) * (94208 + 90112)
        roof = {"bound": "hbm", "kernel": dom, "achieved": alg / (dom_ms_per_step * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                "note": "algorithmic bytes per step = tiles x (634,880 + 618,496) B of stash reads (+ 184,320 B per coarse tile for the "
                        "divergence term); the time includes the three split reductions"}
    else:
        alg = args.n_rand * POINTS_PER_RAY * FLOP_PER_POINT
        roof = {"bound": "tensor", "kernel": dom, "achieved": alg / (dom_ms_per_step * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                "note": "algorithmic FLOPs per step of this kernel kind (coarse + fine launch) = N_rand x 192 x 1,016,320"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    # DRAM bytes of that kernel kind per step, from the committed `ncu --set full` capture of this workload
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.n_rand == N_RAND:
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    roof["traffic"] = traffic
    roof["peak_source"] = peak_src
    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 (tensor-core operands; f32 accumulate)", "data": "synthetic",
        "config": {"workload": f"example_sequence training step: N_rand={args.n_rand}/GPU, 64c+128f, 8x256 MLP, ray bending on, "
                               "perturb=1, raw_noise_std=1, offsets+rigidity+divergence regularisers, backward, Adam",
                   "parallelism": f"ray-sharded x{world}, one flat NCCL grad all-reduce per step",
                   "l2": "per-step working set (activation + gradient stash ~1.9 GB) exceeds the 126 MB L2; 8 rotating input batches"},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(sum(a.nbytes for a in host[0])),
                "d2h_bytes_per_step": 4},
        "gpu_launches": 30 * args.steps,   # this repo's kernels per step: 3 weight packs, 1 coarse sampler, 2 field forwards, 2 composites,
        # 1 ray loss + 4 scalings, 3 divergence, 2 composite backwards, 2 absmax, 2 field DGRADs, 3 WGRADs + 3 reductions, 2 optimizer
        "kernel_ms_per_step": per_step, "ms_per_step_instrumented": ms_instrumented / args.steps,
        "cuda_graph": graphed is not None,
        "roofline": roof,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        threads = min(os.cpu_count() or 1, 32)   # PyTorch's CPU ops stop scaling (and regress) beyond a few dozen threads here
        rate, sec = cpu_training_rate(256, 3, 1, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": "256-ray slice of the same training step, 1 warm-up + median of 3 steps"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""bench.py --workload render | sweep: the BASELINE.json configurations besides the training step.

render (configs[2]): free_viewpoint_rendering.py's full-frame forward -- 504 x 378 = 190,512 rays per frame, fixed camera
  (`--camera_path fixed`: the same rays every frame, a different latent per frame), perturb = 0, raw_noise_std = 0,
  deterministic sample_pdf, 64c + 128f, through render() exactly as render_path calls it (train.py:473-480: one latent row
  expanded to all rays, chunked by `chunk`).  A "step" is one frame.  Reported: rays/s, ms/frame, and the forward kernels'
  tensor-core fraction.  detailed_output=True (the reference's default in render_path) is reported beside it.
sweep (configs[4]): N in {1k, 4k, 16k, 64k, 256k, 1M} rays x S in {64, 128, 256} samples, ONE pass of the field (no
  hierarchical resampling), forward and forward+backward, synthetic unit-cube rays.
Multi-GPU: frames / sweep batches are independent -> every rank renders its own frames (no collective), value = sum.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(args, rank, local_rank, world):
    import torch.distributed as dist
    import bench
    from nonrigid_nerf_b200 import _lib, train as T, run_nerf_helpers as H
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    coarse, fine, bender = bench.build_models(dev, H)
    peaks = bench.read_peaks()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "render":
        Hh, Ww, focal = 378, 504, 252.6
        n = Hh * Ww
        j, i = np.meshgrid(np.arange(Hh, dtype=np.float32), np.arange(Ww, dtype=np.float32), indexing="ij")
        dirs = np.stack([(i - Ww * 0.5) / focal, -(j - Hh * 0.5) / focal, -np.ones_like(i)], -1).reshape(-1, 3).astype(np.float32)
        rays_d_h = torch.from_numpy(dirs).pin_memory()
        rays_o_h = torch.zeros(n, 3).pin_memory()
        rs = np.random.RandomState(7)
        latents_h = torch.from_numpy((rs.randn(256, 32) * 0.1).astype(np.float32)).pin_memory()
        kw = {"network_query_fn": None, "perturb": 0.0, "N_importance": 64, "network_fine": fine, "N_samples": 64, "network_fn": coarse,
              "ray_bender": bender, "use_viewdirs": False, "white_bkgd": False, "raw_noise_std": 0.0, "ndc": False, "lindisp": False,
              "near": 0.0022, "far": 1.0024}
        rays_o, rays_d = rays_o_h.to(dev), rays_d_h.to(dev)
        lat_dev = latents_h.to(dev)
        results = {}
        d2h = {}
        # three ways to get a frame: plain (rgb, disp), "surface" (+ the fused free-viewpoint post-processing: median-visibility
        # sample's canonical point and rigidity, 4 floats + an index per ray), "detailed" (the reference's default in render_path:
        # every per-sample tensor, which render_path then copies to the host, train.py:484-492)
        for mode in ("plain", "surface", "detailed"):
            def frame(k, e2e):
                with torch.no_grad():
                    if e2e:
                        ro, rd = rays_o_h.to(dev, non_blocking=True), rays_d_h.to(dev, non_blocking=True)
                        lat = latents_h[k % 256].to(dev, non_blocking=True)
                    else:
                        ro, rd, lat = rays_o, rays_d, lat_dev[k % 256]
                    rgb, disp, acc, extras = T.render(ro, rd, chunk=65536, additional_pixel_information={"ray_bending_latents": lat[None].expand(n, 32)},
                                                      detailed_output=(mode == "detailed"), surface_output=(mode == "surface"), **kw)
                    if e2e:
                        host = [rgb.cpu(), disp.cpu()]                     # what render_path keeps per frame (train.py:481-483)
                        if mode != "plain":
                            host += [v.cpu() for v in extras.values()]      # train.py:484-492 / the 4 floats per ray
                        d2h[mode] = int(sum(t.numel() * t.element_size() for t in host))
                        return host
                    return rgb
            for k in range(args.warmup):
                frame(k, False)
            out = {}
            n_e2e = args.steps if mode != "detailed" else max(2, args.steps // 5)   # 2.3 GB per frame over PCIe: a few frames suffice
            for e2e in (False, True):
                reps = n_e2e if e2e else args.steps
                sync_all()
                e0, e1 = _events()
                e0.record()
                for k in range(reps):
                    frame(k, e2e)
                e1.record()
                sync_all()
                out[e2e] = max_over_ranks(e0.elapsed_time(e1)) / reps
            results[mode] = out
        _lib.device_error_check()
        if world > 1:
            dist.destroy_process_group()
        if rank != 0:
            return
        ms = results["plain"][False]
        tf = n * bench.POINTS_PER_RAY * bench.FLOP_PER_POINT / (ms * 1e-3) / 1e12
        line = {"metric": "rays/sec (64c+128f samples, 8x256 MLP), full-frame test-time render", "value": world * n / (ms * 1e-3), "unit": "rays/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16 (tensor-core operands; f32 accumulate)", "data": "synthetic",
                "config": {"workload": "free_viewpoint_rendering full-frame forward 504x378, fixed pose, one latent per frame, 64c+128f, det sampling, chunk=65536",
                           "parallelism": f"frames partitioned over {world} rank(s), no collective", "l2": "each frame streams 36.6 M point evaluations; inputs larger than L2"},
                "ms_per_frame": {m: results[m][False] for m in results},
                "ms_per_frame_e2e": {m: results[m][True] for m in results}, "d2h_bytes_per_frame": d2h,
                "e2e": {"value": world * n / (results["plain"][True] * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": n * 24 + 128,
                        "d2h_bytes_per_step": d2h.get("plain")},
                "gpu_launches": 5 * 3 * args.steps,
                "roofline": {"bound": "tensor", "kernel": "field_fwd (whole frame incl. compositing)", "achieved": tf, "peak": peaks["tf_sustained"],
                             "unit": "TFLOP/s", "frac": tf / peaks["tf_sustained"], "frac_burst": tf / peaks["tf_burst"], "traffic": None,
                             "peak_source": peaks["source"]}}
        if not args.no_cpu_baseline:
            from oracle import reference_arm as RA
            host = os.cpu_count() or 1
            cands = sorted({c for c in (host, 64, 32, 16, 8) if c <= host}, reverse=True)
            calib = {c: RA.render_rate(512, 1, c, False)[0] for c in cands}      # same thread calibration as the training arm
            threads = max(calib, key=calib.get)
            rate, sec, kind = RA.render_rate(4096, 2, threads, False)
            line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": threads, "kind": kind, "host_cores": host,
                                    "calibration_rays_per_s_by_threads": {str(k): v for k, v in calib.items()},
                                    "sample": "4096 rays of the same test-time render (64c+128f), 1 warm-up + median of 2"}
        print(json.dumps(line), flush=True)
        return

    # ---- sweep -------------------------------------------------------------------------------------------------
    from nonrigid_nerf_b200 import autograd as ag
    rows = []
    rs = np.random.RandomState(rank)
    for S in (64, 128, 256):
        for n in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20):
            chunk = max(1024, (65536 * 128) // S)         # rays per launch: bounds the training stash to ~80 GB
            o = torch.from_numpy(rs.uniform(-0.2, 0.2, size=(min(n, chunk), 3)).astype(np.float32)).to(dev)
            d = torch.from_numpy(rs.uniform(-1, 1, size=(min(n, chunk), 3)).astype(np.float32)).to(dev)
            rays = torch.cat([o, d, torch.full((o.shape[0], 1), 0.05, device=dev), torch.full((o.shape[0], 1), 1.0, device=dev)], -1)
            z = torch.linspace(0.05, 1.0, S, device=dev)[None].expand(rays.shape[0], S).contiguous()
            lat = (torch.randn(rays.shape[0], 32, device=dev) * 0.1)
            n_launch = (n + chunk - 1) // chunk
            reps = 3 if n >= (1 << 18) else 10

            def fwd():
                with torch.no_grad():
                    for _ in range(n_launch):
                        ag.field(fine, rays, z, lat, False)

            def fwd_bwd():
                for _ in range(n_launch):
                    lat_g = lat.detach().requires_grad_(True)
                    raw, _ = ag.field(fine, rays, z, lat_g, False)
                    raw.backward(torch.ones_like(raw) * 1e-3)

            res = {}
            for name, fn in (("fwd", fwd), ("fwd_bwd", fwd_bwd)):
                fn()
                sync_all()
                e0, e1 = _events()
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                sync_all()
                res[name] = max_over_ranks(e0.elapsed_time(e1)) / reps
            n_eff = n_launch * rays.shape[0]
            pts = n_eff * S
            rows.append({"rays": n_eff, "samples": S, "ms_fwd": res["fwd"], "ms_fwd_bwd": res["fwd_bwd"],
                         "rays_per_s_fwd": world * n_eff / (res["fwd"] * 1e-3), "rays_per_s_fwd_bwd": world * n_eff / (res["fwd_bwd"] * 1e-3),
                         "frac_fwd": pts * bench.FLOP_PER_POINT / (res["fwd"] * 1e-3) / 1e12 / peaks["tf_sustained"],
                         "frac_fwd_bwd": 3 * pts * bench.FLOP_PER_POINT / (res["fwd_bwd"] * 1e-3) / 1e12 / peaks["tf_sustained"]})
            for p in fine.parameters():
                p.grad = None
            for p in bender.parameters():
                p.grad = None
    _lib.device_error_check()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    best = max(rows, key=lambda r: r["rays_per_s_fwd_bwd"] * r["samples"])
    print(json.dumps({"metric": "rays/sec, single-pass field sweep (fwd and fwd+bwd)", "value": best["rays_per_s_fwd_bwd"], "unit": "rays/s",
                      "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": best["ms_fwd_bwd"], "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f16 (tensor-core operands; f32 accumulate)", "data": "synthetic",
                      "config": {"workload": "sweep 1k-1M rays x {64,128,256} samples, one field pass, ray bending on",
                                 "headline": f"{best['rays']} rays x {best['samples']} samples fwd+bwd"},
                      "sweep": rows, "roofline": {"bound": "tensor", "kernel": "field fwd+dgrad+wgrad", "achieved": best["frac_fwd_bwd"] * peaks["tf_sustained"],
                                                  "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": best["frac_fwd_bwd"], "traffic": None}}), flush=True)

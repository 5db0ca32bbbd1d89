"""CPU arm of bench.py: the reference's own training step / render on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE (like everything under oracle/): imported only by bench.py's `--impl reference` leg and
its `cpu_baseline` leg.  When oracle/_ref/ holds the unmodified reference sources (oracle/make_ref.py), the numbers come
from that code (kind = "reference"): training_wrapper_class.forward (train.py:152-287) + backward + torch.optim.Adam, exactly
the step DataParallel wraps; otherwise from the oracle port (kind = "port").  Shims, none touching arithmetic:
empty stand-ins for imageio / matplotlib / load_llff (imported at the top of train.py, unused on the path), and Tensor.get_device returning the device object on CPU.
"""
from __future__ import annotations

import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def load_reference():
    """(train, run_nerf_helpers) modules of the unmodified reference, or None when oracle/_ref is absent."""
    if not (os.path.exists(os.path.join(REF_DIR, "train.py")) and os.path.exists(os.path.join(REF_DIR, "run_nerf_helpers.py"))):
        return None
    for name in ("imageio", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "load_llff" not in sys.modules:      # the dataset loader (train.py:20): a caller of the path, not part of it
        stub = types.ModuleType("load_llff")
        stub.load_llff_data = None
        sys.modules["load_llff"] = stub
    if not getattr(torch.Tensor.get_device, "_nrn_shim", False):
        orig = torch.Tensor.get_device

        def get_device(t):
            return t.device if not t.is_cuda else orig(t)

        get_device._nrn_shim = True
        torch.Tensor.get_device = get_device
    import importlib
    # top-level names `train` / `run_nerf_helpers` (train.py does `from run_nerf_helpers import *`); this repository's own
    # modules are nonrigid_nerf_b200.train / .run_nerf_helpers, so the names do not collide
    sys.path.insert(0, REF_DIR)
    try:
        mods = {"run_nerf_helpers": importlib.import_module("run_nerf_helpers"), "train": importlib.import_module("train")}
    finally:
        sys.path.remove(REF_DIR)
    for m in mods.values():
        if os.path.dirname(os.path.abspath(m.__file__)) != REF_DIR:
            raise RuntimeError(f"oracle.reference_arm: {m.__name__} resolved to {m.__file__}, not to oracle/_ref")
    mods["train"].DEBUG = False
    mods["train"].device = torch.device("cpu")
    return mods["train"], mods["run_nerf_helpers"]


def _reference_models(rt, rh, seed):
    torch.manual_seed(seed)
    embed_fn, input_ch = rh.get_embedder(10, 0)
    bender = rh.ray_bending(input_ch, 32, "simple_neural", embed_fn)
    with torch.no_grad():      # the reference zero-initialises the last layers (identity bending): draw them like bench.py does
        bender.network[-1].weight.normal_(0, 0.01)
        bender.rigidity_network[-1].weight.normal_(0, 0.1)
    kw = dict(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=0, use_viewdirs=False, ray_bender=bender,
              ray_bending_latent_size=32, embeddirs_fn=None, approx_nonrigid_viewdirs=True, time_conditioned_baseline=False)
    coarse, fine = rh.NeRF(num_ray_samples=64, **kw), rh.NeRF(num_ray_samples=128, **kw)

    def network_query_fn(inputs, viewdirs, additional_pixel_information, network_fn, detailed_output=False):
        return rt.run_network(inputs, viewdirs, additional_pixel_information, network_fn, embed_fn=embed_fn, embeddirs_fn=None,
                              netchunk=65536, detailed_output=detailed_output)

    kwargs = {"network_query_fn": network_query_fn, "perturb": 1.0, "N_importance": 64, "network_fine": fine, "N_samples": 64,
              "network_fn": coarse, "ray_bender": bender, "use_viewdirs": False, "white_bkgd": False, "raw_noise_std": 1.0,
              "ndc": False, "lindisp": False, "near": 0.0022, "far": 1.0024}
    return coarse, fine, bender, kwargs


def training_rate(synth_batch, n_rays: int, steps: int, warmup: int, threads: int, targs):
    """rays/s of one example_sequence training step on the host cores.  Returns (rate, seconds per step, kind)."""
    torch.set_num_threads(threads)
    ref = load_reference()
    n_images = 86
    rs = np.random.RandomState(4321)
    times = []
    if ref is not None:
        rt, rh = ref
        coarse, fine, bender, kwargs = _reference_models(rt, rh, 0)
        latents = [torch.zeros(32).normal_(0, 0.1).requires_grad_(True) for _ in range(n_images)]
        params = latents + list(bender.parameters()) + list(coarse.parameters()) + list(fine.parameters())
        opt = torch.optim.Adam(params, lr=5e-4, betas=(0.9, 0.999))
        wrapper = rt.training_wrapper_class(coarse, latents, fine_model=fine, ray_bender=bender)
        extras = {"imageid_to_timestepid": list(range(n_images))}
        for it in range(warmup + steps):
            ro, rd, tgt, idx = (torch.from_numpy(a) for a in synth_batch(rs, n_rays, n_images))
            t0 = time.perf_counter()
            opt.zero_grad()
            losses = wrapper(targs, ro, rd, 100, kwargs, tgt, 1000 + it, 0, extras, idx)
            losses.mean().backward()
            opt.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "reference"
    else:
        import oracle.nrnerf_oracle as O
        cp, fp, bp = (O.clone_params(O.make_nerf_params(1, 5, 30.0), True), O.clone_params(O.make_nerf_params(2, 5, 30.0), True),
                      O.clone_params(O.make_bender_params(3), True))
        params = O.flat_param_list(cp) + O.flat_param_list(fp) + O.flat_param_list(bp)
        opt = torch.optim.Adam(params, lr=5e-4, betas=(0.9, 0.999))
        for it in range(warmup + steps):
            r = O.make_rays(100 + it, n_rays)
            rnd = O.make_randomness(100 + it, n_rays, 64, 64)
            lat = r["latents"].clone().requires_grad_(True)
            t0 = time.perf_counter()
            ret = O.render_rays(cp, fp, bp, r["rays_o"], r["rays_d"], r["near"], r["far"], lat, 64, 64, perturb=True, raw_noise_std=1.0, rnd=rnd)
            loss = O.training_loss(ret, r["target"], 60.0, 0.0005, 0.01)
            loss = loss + 3.0 * 0.01 * O.divergence_loss(bp, ret, lat, n_rays, 64)
            opt.zero_grad()
            loss.mean().backward()
            opt.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "port"
    sec = float(np.median(times))
    return n_rays / sec, sec, kind


def render_rate(n_rays: int, reps: int, threads: int, detailed: bool):
    """rays/s of the test-time render (64c + 128f, deterministic) on the host cores.  (rate, seconds, kind)."""
    torch.set_num_threads(threads)
    ref = load_reference()
    rs = np.random.RandomState(99)
    d = rs.randn(n_rays, 3).astype(np.float32) * 0.3
    d[:, 2] = -1.0
    rays_d = torch.from_numpy(d)
    rays_o = torch.zeros(n_rays, 3)
    lat = (torch.randn(1, 32) * 0.1).expand(n_rays, 32)
    times = []
    if ref is not None:
        rt, rh = ref
        coarse, fine, bender, kwargs = _reference_models(rt, rh, 0)
        kwargs = dict(kwargs, perturb=0.0, raw_noise_std=0.0)
        with torch.no_grad():
            for it in range(reps + 1):
                t0 = time.perf_counter()
                rt.render(rays_o, rays_d, chunk=32768, additional_pixel_information={"ray_bending_latents": lat},
                          detailed_output=detailed, **kwargs)
                if it >= 1:
                    times.append(time.perf_counter() - t0)
        kind = "reference"
    else:
        import oracle.nrnerf_oracle as O
        cp, fp, bp = O.make_nerf_params(1, 5, 30.0), O.make_nerf_params(2, 5, 30.0), O.make_bender_params(3)
        with torch.no_grad():
            for it in range(reps + 1):
                t0 = time.perf_counter()
                O.render_rays(cp, fp, bp, rays_o, rays_d, 0.0022, 1.0024, lat, 64, 64)
                if it >= 1:
                    times.append(time.perf_counter() - t0)
        kind = "port"
    sec = float(np.median(times))
    return n_rays / sec, sec, kind

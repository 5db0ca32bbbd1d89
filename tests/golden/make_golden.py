#!/usr/bin/env python3
"""Generate golden vectors by EXECUTING THE UNMODIFIED REFERENCE (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Inputs (weights, rays, latents, randomness) are NOT stored: they are
re-created bit-identically from seeds by oracle/nrnerf_oracle.py (numpy RandomState / torch
Generator), so a fixture holds only what the reference computed.

Shims, none of which touch the arithmetic:
  * empty stand-in modules for imageio / matplotlib (imported at train.py:11,16, unused on the path)
  * torch.Tensor.get_device returns the tensor's device on CPU (the reference passes the -1 it
    normally returns as a `device=` argument, e.g. run_nerf_helpers.py:652, train.py:738)
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    for name in ("imageio", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    _orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda t: t.device if not t.is_cuda else _orig(t)
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import train as ref_train  # noqa
        import run_nerf_helpers as ref_helpers  # noqa
    finally:
        os.chdir(cwd)
    ref_train.DEBUG = False
    ref_train.device = torch.device("cpu")
    return ref_train, ref_helpers


def load_nerf(module, p):
    with torch.no_grad():
        for i in range(8):
            module.pts_linears[i].weight.copy_(p["pts_w"][i])
            module.pts_linears[i].bias.copy_(p["pts_b"][i])
        module.output_linear.weight.copy_(p["out_w"])
        module.output_linear.bias.copy_(p["out_b"])


def load_bender(module, p):
    with torch.no_grad():
        for i in range(5):
            module.network[i].weight.copy_(p["net_w"][i])
            if i < 4:
                module.network[i].bias.copy_(p["net_b"][i])
        for i in range(3):
            module.rigidity_network[i].weight.copy_(p["rig_w"][i])
            module.rigidity_network[i].bias.copy_(p["rig_b"][i])


def build_reference_models(rt, rh, O, seed, with_bender=True, density_boost=30.0):
    embed_fn, input_ch = rh.get_embedder(10, 0)
    bp = O.make_bender_params(seed + 2) if with_bender else None
    bender = None
    if with_bender:
        bender = rh.ray_bending(input_ch, 32, "simple_neural", embed_fn)
        load_bender(bender, bp)
    cp = O.make_nerf_params(seed, 5, density_boost)
    fp = O.make_nerf_params(seed + 1, 5, density_boost)
    kw = dict(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=0, use_viewdirs=False,
              ray_bender=bender, ray_bending_latent_size=32, embeddirs_fn=None, approx_nonrigid_viewdirs=True,
              time_conditioned_baseline=False)
    coarse = rh.NeRF(num_ray_samples=64, **kw)
    fine = rh.NeRF(num_ray_samples=128, **kw)
    load_nerf(coarse, cp)
    load_nerf(fine, fp)

    def network_query_fn(inputs, viewdirs, additional_pixel_information, network_fn, detailed_output=False):
        return rt.run_network(inputs, viewdirs, additional_pixel_information, network_fn, embed_fn=embed_fn,
                              embeddirs_fn=None, netchunk=65536, detailed_output=detailed_output)

    kwargs = {"network_query_fn": network_query_fn, "perturb": 0.0, "N_importance": 64, "network_fine": fine,
              "N_samples": 64, "network_fn": coarse, "ray_bender": bender, "use_viewdirs": False,
              "white_bkgd": False, "raw_noise_std": 0.0, "ndc": False, "lindisp": False}
    return coarse, fine, bender, kwargs, (cp, fp, bp)


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def grad_summary(tensors, seed=7):
    rs = np.random.RandomState(seed)
    out = {}
    for name, t in tensors:
        g = t.grad
        if g is None:
            continue
        g = g.detach().reshape(-1)
        idx = rs.randint(0, g.numel(), size=min(64, g.numel()))
        out[name + ".norm"] = np.array([float(g.norm())], dtype=np.float64)
        out[name + ".sum"] = np.array([float(g.double().sum())], dtype=np.float64)
        out[name + ".idx"] = idx.astype(np.int64)
        out[name + ".val"] = g[torch.from_numpy(idx)].numpy().astype(np.float32)
    return out


def main():
    import oracle.nrnerf_oracle as O
    rt, rh = import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    outdir = os.path.dirname(os.path.abspath(__file__))

    # ---------------- case A: cfg1 coarse only, bending on, deterministic -----------------------
    seed, n = 100, 256
    coarse, fine, bender, kw, _ = build_reference_models(rt, rh, O, seed)
    rays = O.make_rays(seed, n)
    kwa = dict(kw); kwa["N_importance"] = 0; kwa["network_fine"] = None
    # NOTE: the reference raises UnboundLocalError for N_importance == 0 with detailed_output=True
    # (train.py:969 reads visibility_weights_0, only bound when N_importance > 0), so cfg1 runs plain.
    with torch.no_grad():
        rgb, disp, acc, extras = rt.render(rays["rays_o"], rays["rays_d"], chunk=32768, near=rays["near"], far=rays["far"],
                                           additional_pixel_information={"ray_bending_latents": rays["latents"]},
                                           detailed_output=False, retraw=True, **kwa)
    np.savez_compressed(os.path.join(outdir, "caseA_coarse_only.npz"), seed=seed, n=n, rgb_map=np32(rgb), disp_map=np32(disp),
                        acc_map=np32(acc), raw=np32(extras["raw"][:32]), keys=np.array(sorted(extras.keys())))

    # ---------------- case B: coarse + fine, deterministic (test-time render) -------------------
    seed, n = 200, 256
    coarse, fine, bender, kw, _ = build_reference_models(rt, rh, O, seed)
    rays = O.make_rays(seed, n)
    with torch.no_grad():
        rgb, disp, acc, extras = rt.render(rays["rays_o"], rays["rays_d"], chunk=100, near=rays["near"], far=rays["far"],
                                           additional_pixel_information={"ray_bending_latents": rays["latents"]},
                                           detailed_output=True, retraw=True, **kw)
    save = dict(seed=seed, n=n, rgb_map=np32(rgb), disp_map=np32(disp), acc_map=np32(acc), rgb0=np32(extras["rgb0"]),
                disp0=np32(extras["disp0"]), acc0=np32(extras["acc0"]), z_std=np32(extras["z_std"]), raw=np32(extras["raw"][:16]))
    for k in ("fine_visibility_weights", "fine_opacity_alpha", "visibility_weights", "opacity_alpha", "fine_input_pts",
              "fine_unmasked_offsets", "fine_masked_offsets", "fine_rigidity_mask", "fine_initial_input_pts",
              "input_pts", "unmasked_offsets", "masked_offsets", "rigidity_mask", "initial_input_pts"):
        save[k] = np32(extras[k][:16])
    save["keys"] = np.array(sorted(extras.keys()))
    np.savez_compressed(os.path.join(outdir, "caseB_coarse_fine_det.npz"), **save)

    # ---------------- case C: training mode (perturb, sigma noise) + loss + gradients -----------
    seed, n = 300, 128
    coarse, fine, bender, kw, _ = build_reference_models(rt, rh, O, seed)
    rays = O.make_rays(seed, n)
    rnd = O.make_randomness(seed, n, 64, 64)
    kwc = dict(kw); kwc["perturb"] = 1.0; kwc["raw_noise_std"] = 1.0
    latents = rays["latents"].clone().requires_grad_(True)
    torch.manual_seed(seed)
    rgb, disp, acc, extras = rt.render(rays["rays_o"], rays["rays_d"], chunk=32768, near=rays["near"], far=rays["far"],
                                       additional_pixel_information={"ray_bending_latents": latents},
                                       detailed_output=True, retraw=True, **kwc)
    # the oracle's Generator(seed) must reproduce the global-RNG stream the reference consumed
    torch.manual_seed(seed)
    chk = [torch.rand(n, 64), torch.randn(n, 64), torch.rand(n, 64), torch.randn(n, 128)]
    for a, b in zip(chk, (rnd["t_rand"], rnd["noise_c"], rnd["u"], rnd["noise_f"])):
        assert torch.equal(a, b), "Generator stream mismatch"
    target = rays["target"]
    offsets_w, rigidity_w, sched = 60.0, 5e-4, 0.05
    loss = rh.img2mse(rgb, target, n) + rh.img2mse(extras["rgb0"], target, n)
    wts = extras["visibility_weights"].detach().view(-1)
    ol = torch.mean((wts * torch.pow(torch.norm(extras["unmasked_offsets"].view(-1, 3), dim=-1),
                                     2.0 - extras["rigidity_mask"].view(-1))).view(n, -1), dim=-1)
    ol = ol + rigidity_w * torch.mean((wts * extras["rigidity_mask"].view(-1)).view(n, -1), dim=-1)
    loss = loss + offsets_w * sched * ol
    loss.mean().backward()
    named = [("coarse." + k, v) for k, v in coarse.named_parameters()] + \
            [("fine." + k, v) for k, v in fine.named_parameters()] + \
            [("bender." + k, v) for k, v in bender.named_parameters()] + [("latents", latents)]
    save = dict(seed=seed, n=n, rgb_map=np32(rgb), disp_map=np32(disp), acc_map=np32(acc), rgb0=np32(extras["rgb0"]),
                z_std=np32(extras["z_std"]), loss=np32(loss), raw=np32(extras["raw"][:8]),
                offsets_w=offsets_w, rigidity_w=rigidity_w, sched=sched, latents_grad=np32(latents.grad))
    save.update(grad_summary(named))
    np.savez_compressed(os.path.join(outdir, "caseC_train.npz"), **save)

    # ---------------- case D: test-time editing knobs ------------------------------------------
    seed, n = 400, 64
    coarse, fine, bender, kw, _ = build_reference_models(rt, rh, O, seed)
    rays = O.make_rays(seed, n)
    bender.rigidity_test_time_cutoff = 0.5
    bender.test_time_scaling = 1.7
    coarse.test_time_nonrigid_object_removal_threshold = 0.52
    fine.test_time_nonrigid_object_removal_threshold = 0.52
    with torch.no_grad():
        rgb, disp, acc, extras = rt.render(rays["rays_o"], rays["rays_d"], chunk=32768, near=rays["near"], far=rays["far"],
                                           additional_pixel_information={"ray_bending_latents": rays["latents"]},
                                           detailed_output=True, retraw=True, **kw)
    np.savez_compressed(os.path.join(outdir, "caseD_knobs.npz"), seed=seed, n=n, cutoff=0.5, scaling=1.7, removal=0.52,
                        rgb_map=np32(rgb), disp_map=np32(disp), acc_map=np32(acc), rgb0=np32(extras["rgb0"]),
                        rigidity_mask=np32(extras["rigidity_mask"][:16]), masked_offsets=np32(extras["masked_offsets"][:16]),
                        raw=np32(extras["raw"][:8]))

    # ---------------- case E: op-level (sample_pdf, raw2outputs incl. edge cases) ---------------
    rs = np.random.RandomState(5)
    nb = 48
    zc = np.sort(rs.uniform(0.1, 2.0, size=(nb, 64)).astype(np.float32), axis=-1)
    bins = torch.from_numpy(0.5 * (zc[:, 1:] + zc[:, :-1]))
    w = rs.uniform(0, 1, size=(nb, 62)).astype(np.float32) ** 4
    w[0] = 0.0                     # all-zero weights: uniform pdf
    w[1, :] = 0.0; w[1, 17] = 1.0  # delta: many ties / denom < 1e-5
    w[2, :30] = 0.0
    w = torch.from_numpy(w)
    u_rand = torch.from_numpy(rs.uniform(0, 1, size=(nb, 64)).astype(np.float32))
    samples_det = rh.sample_pdf(bins, w, 64, det=True)
    _orig_rand = torch.rand
    torch.rand = lambda *a, **k: u_rand.clone()   # inject u without touching the reference
    try:
        samples_rand = rh.sample_pdf(bins, w, 64, det=False)
    finally:
        torch.rand = _orig_rand
    raw = torch.from_numpy(rs.randn(nb, 64, 5).astype(np.float32) * 3.0)
    raw[3, :, 3] = -5.0            # zero density everywhere -> acc 0, disp NaN (0/0)
    raw[4, :, 3] = 50.0            # opaque at the first sample
    rd = torch.from_numpy(rs.randn(nb, 3).astype(np.float32))
    z = torch.from_numpy(zc)
    o = rt.raw2outputs(raw, z, rd, 0.0, False)
    ow = rt.raw2outputs(raw, z, rd, 0.0, True)
    np.savez_compressed(os.path.join(outdir, "caseE_ops.npz"), bins=np32(bins), weights=np32(w), u_rand=np32(u_rand),
                        samples_det=np32(samples_det), samples_rand=np32(samples_rand), raw=np32(raw), z=np32(z), rays_d=np32(rd),
                        rgb_map=np32(o[0]), disp_map=np32(o[1]), acc_map=np32(o[2]), alpha=np32(o[3]), weights_out=np32(o[4]),
                        depth_map=np32(o[5]), rgb_map_white=np32(ow[0]))

    # ---------------- case F: canonical rendering (ray_bender = None) --------------------------
    seed, n = 500, 64
    coarse, fine, bender, kw, _ = build_reference_models(rt, rh, O, seed, with_bender=False)
    rays = O.make_rays(seed, n)
    with torch.no_grad():
        rgb, disp, acc, extras = rt.render(rays["rays_o"], rays["rays_d"], chunk=32768, near=rays["near"], far=rays["far"],
                                           additional_pixel_information={"ray_bending_latents": rays["latents"]},
                                           detailed_output=True, retraw=True, **kw)
    np.savez_compressed(os.path.join(outdir, "caseF_canonical.npz"), seed=seed, n=n, rgb_map=np32(rgb), disp_map=np32(disp),
                        acc_map=np32(acc), rgb0=np32(extras["rgb0"]), keys=np.array(sorted(extras.keys())))

    # ---------------- case G + H: the training wrapper DataParallel wraps (train.py:140-287) -----------
    # H: per-ray loss [N] of the unmodified training_wrapper_class.forward with all three regularisers, its gradients,
    # and (G) the divergence term alone (compute_divergence_loss, run_nerf_helpers.py:22-116) with the Hutchinson
    # probes `e` recorded from the reference's own torch.randn_like call (recorded, not replaced: the arithmetic is the
    # reference's; the oracle and the kernels get the same e injected).
    seed, n, n_img = 700, 96, 7
    coarse, fine, bender, kw, _ = build_reference_models(rt, rh, O, seed)
    rays = O.make_rays(seed, n)
    lat_rs = np.random.RandomState(seed + 11)
    latent_list = [torch.from_numpy((lat_rs.randn(32) * 0.1).astype(np.float32)).requires_grad_(True) for _ in range(n_img)]
    pix = np.stack([lat_rs.randint(0, n_img, size=n), lat_rs.randint(0, 384, size=n), lat_rs.randint(0, 512, size=n)], -1)
    pix = torch.from_numpy(pix.astype(np.int64))
    i2t = [int(v) for v in lat_rs.permutation(n_img)]       # image id -> time-step id (a permutation: exercises the lookup)
    targs = types.SimpleNamespace(chunk=32768, N_samples=64, N_importance=64, N_iters=200000, offsets_loss_weight=60.0,
                                  divergence_loss_weight=3.0, rigidity_loss_weight=0.0005, ray_bending_latent_size=32)
    global_step = 50000
    kwh = dict(kw); kwh["perturb"] = 1.0; kwh["raw_noise_std"] = 1.0
    kwh["near"], kwh["far"] = rays["near"], rays["far"]
    wrapper = rt.training_wrapper_class(coarse, latent_list, fine_model=fine, ray_bender=bender)
    recorded = []
    _orig_randn_like = torch.randn_like

    def recording_randn_like(t, *a, **k):
        out = _orig_randn_like(t, *a, **k)
        recorded.append(out.detach().clone())
        return out

    torch.manual_seed(seed)
    torch.randn_like = recording_randn_like
    try:
        loss = wrapper(targs, rays["rays_o"], rays["rays_d"], 100, kwh, rays["target"], global_step, 0,
                       {"imageid_to_timestepid": i2t}, pix)
    finally:
        torch.randn_like = _orig_randn_like
    assert len(recorded) == 1 and recorded[0].shape == (n * 64, 3)
    e = recorded[0]
    loss.mean().backward()
    named = [("coarse." + k, v) for k, v in coarse.named_parameters()] + \
            [("fine." + k, v) for k, v in fine.named_parameters()] + \
            [("bender." + k, v) for k, v in bender.named_parameters()]
    save = dict(seed=seed, n=n, n_img=n_img, pix=pix.numpy(), i2t=np.asarray(i2t, dtype=np.int64), global_step=global_step,
                N_iters=targs.N_iters, offsets_w=targs.offsets_loss_weight, divergence_w=targs.divergence_loss_weight,
                rigidity_w=targs.rigidity_loss_weight, loss=np32(loss), e=np32(e),
                latent_table=np.stack([np32(l) for l in latent_list]),
                latent_grads=np.stack([np32(l.grad) if l.grad is not None else np.zeros(32, np.float32) for l in latent_list]))
    save.update(grad_summary(named))
    np.savez_compressed(os.path.join(outdir, "caseH_training_wrapper.npz"), **save)

    # G: the divergence term on its own, same models / rays / probes, deterministic render (perturb = 0, no noise)
    for m in (coarse, fine, bender):
        m.zero_grad()
    lat_rows = torch.stack([l.detach() for l in latent_list], 0)[torch.tensor(i2t)[pix[:, 0]]].clone().requires_grad_(True)
    kwg = dict(kw); kwg["near"], kwg["far"] = rays["near"], rays["far"]
    rgb, disp, acc, extras = rt.render(rays["rays_o"], rays["rays_d"], chunk=32768, retraw=True,
                                       additional_pixel_information={"ray_bending_latents": lat_rows}, detailed_output=True, **kwg)
    wdiv = 1.0 - torch.exp(-torch.nn.functional.relu(extras["opacity_alpha"].view(-1)))
    dlat = lat_rows.view(n, 1, -1).expand(n, 64, 32).reshape(-1, 32)
    torch.randn_like = lambda t, *a, **k: e.clone()
    try:
        div = rh.compute_divergence_loss(extras["masked_offsets"].view(-1, 3), extras["initial_input_pts"].view(-1, 3), dlat, bender,
                                         exact=False, chunk=32768, N_rays=n, weights=wdiv, backprop_into_weights=False)
    finally:
        torch.randn_like = _orig_randn_like
    div.mean().backward()
    named = [("bender." + k, v) for k, v in bender.named_parameters()]
    save = dict(seed=seed, n=n, div=np32(div), e=np32(e), latents=np32(lat_rows), latents_grad=np32(lat_rows.grad))
    save.update(grad_summary(named))
    np.savez_compressed(os.path.join(outdir, "caseG_divergence.npz"), **save)

    # ---------------- case I: ray generation (run_nerf_helpers.py:588-622) ------------------------------
    rs = np.random.RandomState(31)
    intrin = {"height": 24, "width": 40, "focal_x": 31.7, "focal_y": 30.9, "center_x": 19.3, "center_y": 12.4}
    q, _ = np.linalg.qr(rs.randn(3, 3))
    c2w = np.concatenate([q, rs.randn(3, 1) * 0.3], 1).astype(np.float32)
    ro_np, rd_np = rh.get_rays_np(c2w, intrin)
    ro_t, rd_t = rh.get_rays(torch.from_numpy(c2w), intrin)
    assert np.array_equal(rd_np.astype(np.float32), rd_t.numpy()), "get_rays and get_rays_np disagree"
    np.savez_compressed(os.path.join(outdir, "caseI_get_rays.npz"), c2w=c2w, rays_o=np32(ro_t), rays_d=np32(rd_t),
                        **{k: np.float64(v) for k, v in intrin.items()})

    # ---------------- case J: free-viewpoint surface selection (free_viewpoint_rendering.py:617-658) -----
    # the reference does this inline in a script function; the lines are executed here on a reference render
    seed, n = 800, 192
    coarse, fine, bender, kw, _ = build_reference_models(rt, rh, O, seed)
    rays = O.make_rays(seed, n)
    lat_one = rays["latents"][:1].expand(n, 32)          # render_path: one latent for the whole frame (train.py:465)
    with torch.no_grad():
        rgb, disp, acc, det = rt.render(rays["rays_o"], rays["rays_d"], chunk=32768, near=rays["near"], far=rays["far"],
                                        additional_pixel_information={"ray_bending_latents": lat_one}, detailed_output=True, retraw=True, **kw)
    accumulated_visibility = torch.cumsum(det["fine_visibility_weights"], dim=-1)
    median_indices = torch.min(torch.abs(accumulated_visibility - 0.5), dim=-1)[1]
    surface = det["fine_input_pts"].numpy().reshape(n, -1, 3)[np.arange(n), median_indices.numpy(), :]
    rigidity = det["fine_rigidity_mask"].numpy().reshape(n, -1)[np.arange(n), median_indices.numpy()]
    np.savez_compressed(os.path.join(outdir, "caseJ_surface.npz"), seed=seed, n=n, rgb_map=np32(rgb), median_indices=median_indices.numpy(),
                        surface_pts=surface.astype(np.float32), surface_rigidity=rigidity.astype(np.float32),
                        fine_visibility_weights=np32(det["fine_visibility_weights"]))

    # ---------------- case K: view-dependent head + approximate non-rigid view directions (row f1) -----------
    # use_viewdirs=True, approx_nonrigid_viewdirs=True (run_nerf_helpers.py:233-236, 284-304, 316-356; train.py:364-381).
    # Not implemented by the CUDA path yet (it raises): this case pins the ORACLE's restatement for the round that builds it.
    seed, n = 900, 64
    embed_fn, input_ch = rh.get_embedder(10, 0)
    embeddirs_fn, input_ch_views = rh.get_embedder(4, 0)
    bp = O.make_bender_params(seed + 2)
    bender = rh.ray_bending(input_ch, 32, "simple_neural", embed_fn)
    load_bender(bender, bp)
    mods = []
    for k, ns in ((0, 64), (1, 128)):
        cp_k, vp_k = O.make_nerf_params(seed + k, 5, 30.0), O.make_view_params(seed + 10 + k, 30.0)
        m = rh.NeRF(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=input_ch_views, use_viewdirs=True,
                    ray_bender=bender, ray_bending_latent_size=32, embeddirs_fn=embeddirs_fn, num_ray_samples=ns,
                    approx_nonrigid_viewdirs=True, time_conditioned_baseline=False)
        with torch.no_grad():
            for i in range(8):
                m.pts_linears[i].weight.copy_(cp_k["pts_w"][i]); m.pts_linears[i].bias.copy_(cp_k["pts_b"][i])
            m.alpha_linear.weight.copy_(vp_k["alpha_w"]); m.alpha_linear.bias.copy_(vp_k["alpha_b"])
            m.feature_linear.weight.copy_(vp_k["feature_w"]); m.feature_linear.bias.copy_(vp_k["feature_b"])
            m.views_linears[0].weight.copy_(vp_k["views_w"]); m.views_linears[0].bias.copy_(vp_k["views_b"])
            m.rgb_linear.weight.copy_(vp_k["rgb_w"]); m.rgb_linear.bias.copy_(vp_k["rgb_b"])
        mods.append(m)
    coarse, fine = mods

    def query_fn_views(inputs, viewdirs, additional_pixel_information, network_fn, detailed_output=False):
        return rt.run_network(inputs, viewdirs, additional_pixel_information, network_fn, embed_fn=embed_fn,
                              embeddirs_fn=embeddirs_fn, netchunk=64 * 128, detailed_output=detailed_output)

    kwk = {"network_query_fn": query_fn_views, "perturb": 0.0, "N_importance": 64, "network_fine": fine, "N_samples": 64,
           "network_fn": coarse, "ray_bender": bender, "white_bkgd": False, "raw_noise_std": 0.0, "ndc": False, "lindisp": False}
    rays = O.make_rays(seed, n)
    latents = rays["latents"].clone().requires_grad_(True)
    rgb, disp, acc, extras = rt.render(rays["rays_o"], rays["rays_d"], chunk=32768, near=rays["near"], far=rays["far"], use_viewdirs=True,
                                       additional_pixel_information={"ray_bending_latents": latents}, detailed_output=True,
                                       retraw=True, **kwk)
    loss = rh.img2mse(rgb, rays["target"], n) + rh.img2mse(extras["rgb0"], rays["target"], n)
    loss.mean().backward()
    named = [("coarse." + k, v) for k, v in coarse.named_parameters()] + [("fine." + k, v) for k, v in fine.named_parameters()] + \
            [("bender." + k, v) for k, v in bender.named_parameters()]
    save = dict(seed=seed, n=n, rgb_map=np32(rgb), rgb0=np32(extras["rgb0"]), acc_map=np32(acc), raw=np32(extras["raw"][:8]),
                loss=np32(loss), latents_grad=np32(latents.grad))
    save.update(grad_summary(named))
    # static scene (no bender): the view direction is the normalised ray direction itself (train.py:364-381, 80-84)
    for m in mods:
        m.ray_bender = (None,)
    kws = dict(kwk, ray_bender=None)
    with torch.no_grad():
        rgb_s, _, acc_s, extras_s = rt.render(rays["rays_o"], rays["rays_d"], chunk=32768, near=rays["near"], far=rays["far"],
                                              use_viewdirs=True, additional_pixel_information={"ray_bending_latents": rays["latents"]},
                                              retraw=True, **kws)
    save.update(static_rgb_map=np32(rgb_s), static_rgb0=np32(extras_s["rgb0"]), static_raw=np32(extras_s["raw"][:8]))
    np.savez_compressed(os.path.join(outdir, "caseK_viewdirs.npz"), **save)
    print("golden vectors written to", outdir)


if __name__ == "__main__":
    main()

"""torch.autograd glue between the PyTorch-owned parameters and the fused kernels.

PyTorch owns every tensor (weights stay nn.Parameters with the reference's names); the extension
keeps a derived packed fp16 copy (ops.pack_*), rebuilt when a parameter's version counter changes.
Gradients come back from the DGRAD/WGRAD kernels as flat fp32 buffers laid out in the reference's
parameter shapes and are handed to autograd as views, so the reference's Adam, its two-pass
test-latent backward (train.py:1595-1608) and the PyTorch-side regularisers keep working unchanged.

Gradient flow implemented: rgb_map / acc_map -> raw -> both MLPs -> positional
encoding -> bent point -> bender weights and per-ray latents; plus upstream gradients on the coarse
`unmasked_offsets` / `rigidity_mask` details (offsets / rigidity regularisers, train.py:219-242).
z_vals, ray origins/directions and the importance samples carry no gradient (train.py:918).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib, ops


def _knobs(net):
    bender = net.ray_bender[0]
    cutoff = getattr(bender, "rigidity_test_time_cutoff", None) if bender is not None else None
    scaling = getattr(bender, "test_time_scaling", None) if bender is not None else None
    removal = getattr(net, "test_time_nonrigid_object_removal_threshold", None)
    return cutoff, scaling, removal


def _flat_params(net, bender):
    """Parameters in the flat order of the WGRAD output buffers (csrc/wgrad.cu)."""
    ws, bs = ops.nerf_param_list(net)
    nerf = []
    for w, b in zip(ws, bs):
        nerf += [w, b]
    bend = []
    if bender is not None:
        net_w, net_b, rig_w, rig_b = ops.bender_param_list(bender)
        for i in range(4):
            bend += [net_w[i], net_b[i]]
        bend.append(net_w[4])
        for i in range(3):
            bend += [rig_w[i], rig_b[i]]
    return nerf, bend


def _split_flat(flat: torch.Tensor, like):
    out, o = [], 0
    for p in like:
        n = p.numel()
        out.append(flat[o:o + n].view_as(p))
        o += n
    assert o == flat.numel(), (o, flat.numel())
    return out


class _FieldTrainFn(torch.autograd.Function):
    """raw, unmasked_offsets, rigidity_mask (differentiable) + point details (not differentiable).
    params = the NeRF's parameters in WGRAD order, followed (with a bender) by the bender's parameters in WGRAD order.
    (No autograd object outlives an iteration: a cached graph fragment would pin AccumulateGrad nodes -- and the CUDA
    stream they were created on -- across iterations, which breaks CUDA-graph capture of the step.)"""

    @staticmethod
    def forward(ctx, net, rays, z_vals, latents, n_nerf, *params):
        bender = net.ray_bender[0]
        cutoff, scaling, removal = _knobs(net)
        if removal is not None:
            raise RuntimeError("nonrigid_nerf_b200: test_time_nonrigid_object_removal_threshold is a test-time knob; "
                               "it is not differentiable")
        nerf_pack = ops.pack_nerf(net)
        bender_pack = ops.pack_bender(bender) if bender is not None else None
        out_ch = net.output_linear.weight.shape[0]
        n, s = z_vals.shape
        lib = _lib.load()
        stash = torch.empty(lib.nrn_stash_bytes(n, s), dtype=torch.uint8, device=z_vals.device)
        raw, det = ops.field_forward(rays, z_vals, latents, nerf_pack, bender_pack, out_ch, cutoff, scaling, None, True, stash)
        ctx.net, ctx.n_nerf = net, n_nerf
        ctx.shape = (n, s, out_ch)
        ctx.knobs = (cutoff, scaling)
        ctx.packs = (nerf_pack, bender_pack)
        ctx.stash = stash
        ctx.params = params
        ctx.set_materialize_grads(False)
        if bender is not None:
            _register_stash(det["unmasked_offsets"], stash)
            ctx.save_for_backward(det["unmasked_offsets"], det["rigidity_mask"])
            outs = (raw, det["unmasked_offsets"], det["rigidity_mask"], det["initial_input_pts"], det["input_pts"],
                    det["masked_offsets"])
            ctx.mark_non_differentiable(*outs[3:])
        else:
            outs = (raw, det["initial_input_pts"], det["input_pts"])
            ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, d_raw, *rest):
        net = ctx.net
        bender = net.ray_bender[0]
        n, s, out_ch = ctx.shape
        nerf_pack, bender_pack = ctx.packs
        cutoff, scaling = ctx.knobs
        dev = ctx.stash.device
        lib = _lib.load()
        d_un = d_rig = None
        if bender is not None:
            d_un, d_rig = rest[0], rest[1]
        if d_raw is None:
            d_raw = torch.zeros(n, s, out_ch, dtype=torch.float32, device=dev)
        a = _lib.NrnFieldBwdArgs()
        a.n_rays, a.n_samples, a.out_ch = n, s, out_ch
        d_raw = d_raw.contiguous().float()
        a.d_raw = d_raw.data_ptr()
        a.stash = ctx.stash.data_ptr()
        gstash = torch.empty(lib.nrn_grad_stash_bytes(n, s), dtype=torch.uint8, device=dev)
        scratch = torch.empty(lib.nrn_wgrad_scratch_bytes(), dtype=torch.uint8, device=dev)
        a.grad_stash, a.wgrad_scratch = gstash.data_ptr(), scratch.data_ptr()
        a.nerf_packed = nerf_pack.data_ptr()
        # Where the weight gradients go.  If the .grad tensors of this module's parameters lie back to back in one buffer
        # (optim.Adam's arena) the WGRAD reduction ADDS into them in place and autograd gets nothing to accumulate;
        # otherwise fresh flat buffers are handed to autograd as per-parameter views (torch.optim.Adam, or after the
        # caller re-bound gradients -- the reference sets weights.grad = None between its two backward passes,
        # train.py:1598-1604).
        nerf_p = list(ctx.params[:ctx.n_nerf])
        pts_dst = _arena_destination(nerf_p[:-2]) if all(p.requires_grad for p in nerf_p) else None
        head_dst = _arena_destination(nerf_p[-2:]) if pts_dst is not None else None
        nerf_grad = None
        if head_dst is not None:
            a.nerf_grad, a.nerf_grad_head, a.accumulate_nerf = pts_dst, head_dst, 1
        else:
            nerf_grad = torch.empty(lib.nrn_nerf_grad_floats(out_ch), dtype=torch.float32, device=dev)
            a.nerf_grad = nerf_grad.data_ptr()
        bend_grad = d_lat = None
        bend_in_place = False
        keep = [d_raw]
        if bender is not None:
            un, rig = ctx.saved_tensors
            a.bender_packed = bender_pack.data_ptr()
            a.unmasked_offsets, a.rigidity_mask = un.data_ptr(), rig.data_ptr()
            if d_un is not None:
                d_un = d_un.contiguous().float()
                a.d_unmasked_offsets = d_un.data_ptr()
                keep.append(d_un)
            if d_rig is not None:
                d_rig = d_rig.contiguous().float()
                a.d_rigidity_mask = d_rig.data_ptr()
                keep.append(d_rig)
            if cutoff is not None:
                a.use_cutoff, a.rigidity_cutoff = 1, float(cutoff)
            if scaling is not None:
                a.use_scaling, a.scaling = 1, float(scaling)
            bend_dst = _bender_arena(bender)
            if bend_dst is not None:
                a.bender_grad, a.accumulate_bender, bend_in_place = bend_dst, 1, True
            else:
                bend_grad = torch.empty(lib.nrn_bender_grad_floats(), dtype=torch.float32, device=dev)
                a.bender_grad = bend_grad.data_ptr()
            d_lat = torch.empty(n, ops.LATENT, dtype=torch.float32, device=dev)
            a.d_latents = d_lat.data_ptr()
        a.stream = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib.nrn_field_backward(C.byref(a)), "field_backward")
        # the stash lives as long as the autograd node: backward(retain_graph=True) followed by a second backward()
        # over the same graph (test-latent pass of the reference loop, train.py:1595-1606) reads it again
        if nerf_grad is None:
            grads = [None] * len(nerf_p)
        else:
            grads = [g if p.requires_grad else None for g, p in zip(_split_flat(nerf_grad, nerf_p), nerf_p)]
        if bender is not None:
            bend_p = list(ctx.params[ctx.n_nerf:])
            if bend_in_place:
                grads += [None] * len(bend_p)
            else:
                grads += [g if p.requires_grad else None for g, p in zip(_split_flat(bend_grad, bend_p), bend_p)]
        return (None, None, None, d_lat, None, *grads)


# ---------------------------------------------------------------------------------------------
# gradient arena look-ups (optim.Adam seats every .grad as a view of one flat buffer)
# ---------------------------------------------------------------------------------------------
def _arena_destination(params):
    from .optim import arena_destination
    return arena_destination(list(params))


def _bender_arena(bender):
    _, bend_p = _flat_params_bender(bender)
    if not all(p.requires_grad for p in bend_p):
        return None
    return _arena_destination(bend_p)


class _LatentGatherFn(torch.autograd.Function):
    """latents[timestep[i]] for every ray i (train.py:173-189: stack the per-frame latents, index by the ray's time step).
    The per-frame latents are views of the optimizer's flat buffer, so the table is read in place; the backward adds the
    per-ray gradients [N, Z] into the latents' .grad arena with one index_add_ (fallback: per-latent gradient views)."""

    @staticmethod
    def forward(ctx, timestep, *latents):
        z = latents[0].numel()
        base, contiguous = latents[0].data_ptr(), True
        for i, l in enumerate(latents):
            if l.data_ptr() != base + 4 * z * i or l.dtype != torch.float32 or not l.is_contiguous():
                contiguous = False
                break
        if contiguous and latents[0].untyped_storage().nbytes() >= latents[0].storage_offset() * 4 + 4 * z * len(latents):
            table = latents[0].detach().as_strided((len(latents), z), (z, 1))
        else:
            table = torch.stack([l.detach() for l in latents], 0)
        ctx.save_for_backward(timestep)
        ctx.latents = latents
        ctx.z = z
        return torch.index_select(table, 0, timestep)

    @staticmethod
    def backward(ctx, d_sel):
        (timestep,) = ctx.saved_tensors
        latents, z = ctx.latents, ctx.z
        t = len(latents)
        dst = _arena_destination(latents) if all(l.requires_grad for l in latents) else None
        if dst is not None:
            g0 = latents[0].grad
            g0.as_strided((t, z), (z, 1)).index_add_(0, timestep, d_sel.contiguous().float())
            return (None,) * (t + 1)
        g = torch.zeros(t, z, dtype=torch.float32, device=d_sel.device).index_add_(0, timestep, d_sel.float())
        return (None, *[g[i] if l.requires_grad else None for i, l in enumerate(latents)])


def gather_latents(latents, timestep: torch.Tensor) -> torch.Tensor:
    """[N, Z] per-ray latents from the list of per-frame leaf tensors and the rays' time-step ids."""
    return _LatentGatherFn.apply(timestep, *latents)


# ---------------------------------------------------------------------------------------------
# divergence regulariser (fused)
# ---------------------------------------------------------------------------------------------
_STASH_BY_PTR = {}   # data_ptr of a coarse pass's unmasked_offsets -> weakref to that pass's activation stash


def _register_stash(unmasked: torch.Tensor, stash: torch.Tensor) -> None:
    import weakref
    for k in [k for k, v in _STASH_BY_PTR.items() if v() is None]:
        del _STASH_BY_PTR[k]
    _STASH_BY_PTR[unmasked.data_ptr()] = weakref.ref(stash)


def lookup_stash(unmasked: torch.Tensor) -> Optional[torch.Tensor]:
    """The activation stash of the coarse pass that produced `unmasked`: found by walking the tensor's autograd history
    (through the reshapes of render()) to the _FieldTrainFn node, which owns the stash; the address table is only the
    fallback for detached tensors."""
    fn = unmasked.grad_fn
    for _ in range(8):
        if fn is None:
            break
        stash = getattr(fn, "stash", None)
        if isinstance(stash, torch.Tensor):
            return stash
        nxt = [f for f, _ in fn.next_functions if f is not None]
        fn = nxt[0] if len(nxt) == 1 else None
    ref = _STASH_BY_PTR.get(unmasked.data_ptr())
    return ref() if ref is not None else None


class _DivergenceFn(torch.autograd.Function):
    """per-ray mean_s(w * (e^T J e)^2) of the offset field, closed-form forward and backward (csrc/div.cu)."""

    @staticmethod
    def forward(ctx, unmasked, rigidity, weights, e, stash, bender, w_is_alpha, *bend_p):
        n, s = unmasked.shape[0], unmasked.shape[1]
        dev = unmasked.device
        lib = _lib.load()
        a = _lib.NrnDivArgs()
        a.n_rays, a.n_samples = n, s
        un = unmasked.detach().contiguous().float()
        rg = rigidity.detach().contiguous().float()
        w = weights.detach().contiguous().float()
        e = e.contiguous().float()
        net_w, _, rig_w, _ = ops.bender_param_list(bender)
        net_arr = ops._ptr_array([t.detach() for t in net_w])
        rig_arr = ops._ptr_array([t.detach() for t in rig_w])
        tan = torch.empty(lib.nrn_div_stash_bytes(n, s), dtype=torch.uint8, device=dev)
        scal = torch.empty(4, n * s, dtype=torch.float32, device=dev)
        loss = torch.empty(n, dtype=torch.float32, device=dev)
        a.stash, a.e, a.unmasked_offsets, a.rigidity_mask, a.weights = stash.data_ptr(), e.data_ptr(), un.data_ptr(), rg.data_ptr(), w.data_ptr()
        a.weights_are_opacity_alpha = 1 if w_is_alpha else 0
        a.net_w, a.rig_w = net_arr, rig_arr
        a.tangent_stash = tan.data_ptr()
        a.d, a.alpha, a.beta, a.tau_c = (scal[i].data_ptr() for i in range(4))
        a.loss = loss.data_ptr()
        a.stream = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib.nrn_divergence_forward(C.byref(a)), "divergence_forward")
        ctx.keep = (un, rg, w, e, stash, tan, scal, bender)
        ctx.bend_p = bend_p
        ctx.w_is_alpha = bool(w_is_alpha)
        ctx.shape = (n, s)
        ctx.in_shapes = (unmasked.shape, rigidity.shape)
        return loss

    @staticmethod
    def backward(ctx, g):
        un, rg, w, e, stash, tan, scal, bender = ctx.keep
        n, s = ctx.shape
        dev = un.device
        lib = _lib.load()
        # G = dL/dd per point = g_ray * 2 * w * d / S, computed (with its max, the loss-scale source) by the library
        g = g.reshape(n).contiguous().float()
        G = torch.empty(n * s, dtype=torch.float32, device=dev)
        a = _lib.NrnDivArgs()
        a.n_rays, a.n_samples = n, s
        net_w, _, rig_w, _ = ops.bender_param_list(bender)
        net_arr = ops._ptr_array([t.detach() for t in net_w])
        rig_arr = ops._ptr_array([t.detach() for t in rig_w])
        a.stash, a.e, a.unmasked_offsets, a.rigidity_mask, a.weights = stash.data_ptr(), e.data_ptr(), un.data_ptr(), rg.data_ptr(), w.data_ptr()
        a.weights_are_opacity_alpha = 1 if ctx.w_is_alpha else 0
        a.net_w, a.rig_w = net_arr, rig_arr
        a.tangent_stash = tan.data_ptr()
        a.d, a.alpha, a.beta, a.tau_c = (scal[i].data_ptr() for i in range(4))
        a.g_ray, a.G_workspace = g.data_ptr(), G.data_ptr()
        adj = torch.empty(lib.nrn_div_grad_stash_bytes(n, s), dtype=torch.uint8, device=dev)
        scratch = torch.empty(lib.nrn_wgrad_scratch_bytes(), dtype=torch.uint8, device=dev)
        d_un = torch.empty(n * s, 3, dtype=torch.float32, device=dev)
        d_rg = torch.empty(n * s, dtype=torch.float32, device=dev)
        bend_dst = _bender_arena(bender)
        bend_grad = None
        if bend_dst is not None:
            a.bender_grad, a.accumulate_bender = bend_dst, 1
        else:
            bend_grad = torch.empty(lib.nrn_bender_grad_floats(), dtype=torch.float32, device=dev)
            a.bender_grad = bend_grad.data_ptr()
        a.adjoint_stash, a.wgrad_scratch = adj.data_ptr(), scratch.data_ptr()
        a.d_unmasked_offsets, a.d_rigidity_mask = d_un.data_ptr(), d_rg.data_ptr()
        a.stream = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib.nrn_divergence_backward(C.byref(a)), "divergence_backward")
        bend_p = list(ctx.bend_p)
        if bend_grad is None:
            pgrads = [None] * len(bend_p)
        else:
            pgrads = [g if p.requires_grad else None for g, p in zip(_split_flat(bend_grad, bend_p), bend_p)]
        return (d_un.view(ctx.in_shapes[0]), d_rg.view(ctx.in_shapes[1]), None, None, None, None, None, *pgrads)


def divergence_loss(unmasked: torch.Tensor, rigidity: torch.Tensor, weights: Optional[torch.Tensor], bender,
                    e: Optional[torch.Tensor] = None, opacity_alpha: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused divergence regulariser on the coarse samples of the LAST differentiable coarse pass.
    unmasked [N,S,3], rigidity [N,S,1] must be that pass's outputs (they locate its activation stash and
    carry the gradient w.r.t. the primal bender evaluation); weights [N,S] are used detached; `e` [N*S,3]
    are the Hutchinson probes (drawn with torch.randn like run_nerf_helpers.py:110 when None).
    Instead of `weights`, `opacity_alpha` [N,S] may be given: the kernels then apply the reference's
    1 - exp(-relu(opacity_alpha)) (train.py:267) themselves."""
    stash = lookup_stash(unmasked)
    if stash is None:
        raise RuntimeError("nonrigid_nerf_b200: no activation stash for these offsets -- the fused divergence term needs the "
                           "un-chunked coarse pass of the current differentiable render() call (N_rand <= chunk)")
    n, s = unmasked.shape[0], unmasked.shape[1]
    if e is None:
        e = torch.randn(n * s, 3, device=unmasked.device)
    _, bend_p = _flat_params_bender(bender)
    if opacity_alpha is not None:
        return _DivergenceFn.apply(unmasked, rigidity, opacity_alpha, e, stash, bender, True, *bend_p)
    return _DivergenceFn.apply(unmasked, rigidity, weights, e, stash, bender, False, *bend_p)


def _flat_params_bender(bender):
    net_w, net_b, rig_w, rig_b = ops.bender_param_list(bender)
    bend = []
    for i in range(4):
        bend += [net_w[i], net_b[i]]
    bend.append(net_w[4])
    for i in range(3):
        bend += [rig_w[i], rig_b[i]]
    return None, bend


class _RayLossFn(torch.autograd.Function):
    """Per-ray loss of training_wrapper_class.forward (train.py:208-287) in one kernel (csrc/loss.cu): data terms, offsets /
    rigidity regulariser, and the (already reduced) divergence regulariser with its weight and the regularisers' schedule."""

    @staticmethod
    def forward(ctx, rgb, rgb0, target, weights, unmasked, rigidity, lam_o, lam_r, sched_step, sched_n_iters, div, lam_div):
        n = rgb.shape[0]
        dev = rgb.device
        lib = _lib.load()
        a = _lib.NrnRayLossArgs()
        keep = [rgb.detach().contiguous().float(), target.contiguous().float()]
        a.rgb, a.target = keep[0].data_ptr(), keep[1].data_ptr()
        u_rgb = torch.empty(n, 3, dtype=torch.float32, device=dev)
        a.u_rgb = u_rgb.data_ptr()
        u_rgb0 = u_off = u_rig = u_div = None
        if rgb0 is not None:
            keep.append(rgb0.detach().contiguous().float())
            a.rgb0 = keep[-1].data_ptr()
            u_rgb0 = torch.empty(n, 3, dtype=torch.float32, device=dev)
            a.u_rgb0 = u_rgb0.data_ptr()
        s = 1
        if unmasked is not None:
            s = unmasked.shape[1]
            keep += [weights.detach().contiguous().float(), unmasked.detach().contiguous().float(), rigidity.detach().contiguous().float()]
            a.weights, a.unmasked_offsets, a.rigidity_mask = (t.data_ptr() for t in keep[-3:])
            u_off = torch.empty(n, s, 3, dtype=torch.float32, device=dev)
            u_rig = torch.empty(rigidity.shape, dtype=torch.float32, device=dev)
            a.u_unmasked_offsets, a.u_rigidity_mask = u_off.data_ptr(), u_rig.data_ptr()
        a.n_rays, a.n_samples = n, s
        a.lam_offsets, a.lam_rigidity = float(lam_o), float(lam_r)
        if sched_step is not None:
            keep.append(sched_step.detach().float().contiguous())
            a.sched_step, a.sched_n_iters = keep[-1].data_ptr(), float(sched_n_iters)
        if div is not None:
            keep.append(div.detach().contiguous().float())
            u_div = torch.empty(n, dtype=torch.float32, device=dev)
            a.divergence, a.lam_divergence, a.u_divergence = keep[-1].data_ptr(), float(lam_div), u_div.data_ptr()
        loss = torch.empty(n, dtype=torch.float32, device=dev)
        a.loss = loss.data_ptr()
        a.stream = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib.nrn_ray_loss(C.byref(a)), "ray_loss")
        ctx.units = (u_rgb, u_rgb0, u_off, u_rig, u_div)
        ctx.ns = (n, s)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        g = g.contiguous().float()
        a = _lib.NrnRayLossBwdArgs()
        a.n_rays, a.n_samples = ctx.ns
        a.g = g.data_ptr()
        outs = [None if u is None else torch.empty_like(u) for u in ctx.units]
        names = ("rgb", "rgb0", "unmasked_offsets", "rigidity_mask", "divergence")
        for nm, u, o in zip(names, ctx.units, outs):
            if u is not None:
                setattr(a, "u_" + nm, u.data_ptr())
                setattr(a, "d_" + nm, o.data_ptr())
        a.stream = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(g.device):
            _lib.check(lib.nrn_ray_loss_backward(C.byref(a)), "ray_loss_backward")
        d_rgb, d_rgb0, d_off, d_rig, d_div = outs
        return d_rgb, d_rgb0, None, None, d_off, d_rig, None, None, None, None, d_div, None


def ray_loss(rgb, rgb0, target, weights=None, unmasked=None, rigidity=None, lam_offsets=0.0, lam_rigidity=0.0,
             sched_step: Optional[torch.Tensor] = None, sched_n_iters: float = 1.0, divergence: Optional[torch.Tensor] = None,
             lam_divergence: float = 0.0):
    """loss[N] = img2mse(rgb) + img2mse(rgb0) + sched * lam_offsets * (offsets + lam_rigidity * rigidity regulariser)
                 + sched * lam_divergence * divergence,   sched = (1/100)^(1 - sched_step / sched_n_iters) evaluated on the device
    from the 0-dim CUDA tensor `sched_step` (CUDA-graph-safe), or 1 when sched_step is None (the caller folds the schedule
    into the weights)."""
    return _RayLossFn.apply(rgb, rgb0, target, weights, unmasked, rigidity, lam_offsets, lam_rigidity, sched_step, sched_n_iters,
                            divergence, lam_divergence)


class _CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, noise, white_bkgd, n_importance, u):
        o = ops.composite(raw, z_vals, rays_d, noise, white_bkgd, n_importance, u, True)
        ctx.save_for_backward(raw, z_vals, rays_d, noise if noise is not None else raw.new_empty(0))
        ctx.has_noise = noise is not None
        ctx.white = bool(white_bkgd)
        ctx.set_materialize_grads(False)
        outs = [o["rgb_map"], o["acc_map"], o["disp_map"], o["depth_map"], o["weights"], o["alpha"]]
        if n_importance > 0:
            outs += [o["z_vals_out"], o["z_std"]]
        ctx.mark_non_differentiable(*outs[2:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_rgb, d_acc, *unsupported):
        raw, z_vals, rays_d, noise = ctx.saved_tensors
        if d_rgb is None and d_acc is None:
            return None, None, None, None, None, None, None
        if d_rgb is None:
            d_rgb = torch.zeros(raw.shape[0], 3, dtype=torch.float32, device=raw.device)
        d_raw = ops.composite_backward(raw, z_vals, rays_d, noise if ctx.has_noise else None, ctx.white, d_rgb, d_acc)
        return d_raw, None, None, None, None, None, None


# ---------------------------------------------------------------------------------------------
# entry points used by train.py / run_nerf_helpers.py
# ---------------------------------------------------------------------------------------------
def _needs_grad(net, latents) -> bool:
    if not torch.is_grad_enabled():
        return False
    bender = net.ray_bender[0]
    if latents is not None and latents.requires_grad:
        return True
    if any(p.requires_grad for p in net.parameters()):
        return True
    return bender is not None and any(p.requires_grad for p in bender.parameters())


def field(net, rays: torch.Tensor, z_vals: torch.Tensor, latents: Optional[torch.Tensor],
          want_details: bool) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Fused field evaluation for rays x samples; differentiable when autograd is recording."""
    bender = net.ray_bender[0]
    if not _needs_grad(net, latents):
        return field_rays(net, rays, z_vals, latents, want_details)
    nerf_p, bend_p = _flat_params(net, bender)
    outs = _FieldTrainFn.apply(net, rays, z_vals, latents, len(nerf_p), *nerf_p, *bend_p)
    if bender is not None:
        raw, un, rig, init, bent, masked = outs
        details = {"initial_input_pts": init, "unmasked_offsets": un, "rigidity_mask": rig, "masked_offsets": masked,
                   "input_pts": bent}
    else:
        raw, init, bent = outs
        details = {"initial_input_pts": init, "input_pts": bent}
    return raw, (details if want_details else {})


def field_rays(net, rays, z_vals, latents, want_details):
    """Inference path (no stash)."""
    bender = net.ray_bender[0]
    cutoff, scaling, removal = _knobs(net)
    nerf_pack = ops.pack_nerf(net)
    bender_pack = ops.pack_bender(bender) if bender is not None else None
    out_ch = net.output_linear.weight.shape[0]
    return ops.field_forward(rays, z_vals, latents, nerf_pack, bender_pack, out_ch, cutoff, scaling, removal, want_details)


def field_points(net, pts, latents, want_details):
    """NeRF.forward(x) semantics: one xyz (+ latent) per row.  Inference only."""
    if _needs_grad(net, latents):
        raise RuntimeError("nonrigid_nerf_b200: the point-wise NeRF.forward / run_network entry is inference-only; "
                           "differentiable rendering goes through render() / render_rays()")
    bender = net.ray_bender[0]
    cutoff, scaling, removal = _knobs(net)
    nerf_pack = ops.pack_nerf(net)
    bender_pack = ops.pack_bender(bender) if bender is not None else None
    out_ch = net.output_linear.weight.shape[0]
    return ops.field_forward_points(pts, latents, nerf_pack, bender_pack, out_ch, cutoff, scaling, removal, want_details)


def composite(raw, z_vals, rays_d, noise=None, white_bkgd=False, n_importance=0, u=None) -> Dict[str, torch.Tensor]:
    """raw2outputs (+ hierarchical resampling); differentiable w.r.t. raw through rgb_map / acc_map."""
    if torch.is_grad_enabled() and raw.requires_grad:
        outs = _CompositeFn.apply(raw, z_vals, rays_d, noise, white_bkgd, n_importance, u)
        keys = ["rgb_map", "acc_map", "disp_map", "depth_map", "weights", "alpha"] + (["z_vals_out", "z_std"] if n_importance > 0 else [])
        return dict(zip(keys, outs))
    return ops.composite(raw, z_vals, rays_d, noise, white_bkgd, n_importance, u, True)

"""torch.autograd glue between the PyTorch-owned parameters and the fused kernels.

PyTorch owns every tensor (weights stay nn.Parameters with the reference's names); the extension
keeps a derived packed fp16 copy (ops.pack_*), rebuilt when a parameter's version counter changes.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops


def _knobs(net):
    bender = net.ray_bender[0]
    cutoff = getattr(bender, "rigidity_test_time_cutoff", None) if bender is not None else None
    scaling = getattr(bender, "test_time_scaling", None) if bender is not None else None
    removal = getattr(net, "test_time_nonrigid_object_removal_threshold", None)
    return cutoff, scaling, removal


def field_rays(net, rays: torch.Tensor, z_vals: torch.Tensor, latents: Optional[torch.Tensor],
               want_details: bool) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Fused field evaluation for rays x samples (inference path; the differentiable path is
    autograd_train.FieldTrainFn)."""
    bender = net.ray_bender[0]
    cutoff, scaling, removal = _knobs(net)
    nerf_pack = ops.pack_nerf(net)
    bender_pack = ops.pack_bender(bender) if bender is not None else None
    out_ch = net.output_linear.weight.shape[0]
    return ops.field_forward(rays, z_vals, latents, nerf_pack, bender_pack, out_ch, cutoff, scaling, removal, want_details)


def field_points(net, pts: torch.Tensor, latents: Optional[torch.Tensor], want_details: bool):
    """NeRF.forward(x) semantics: one xyz (+ latent) per row."""
    bender = net.ray_bender[0]
    cutoff, scaling, removal = _knobs(net)
    nerf_pack = ops.pack_nerf(net)
    bender_pack = ops.pack_bender(bender) if bender is not None else None
    out_ch = net.output_linear.weight.shape[0]
    return ops.field_forward_points(pts, latents, nerf_pack, bender_pack, out_ch, cutoff, scaling, removal, want_details)

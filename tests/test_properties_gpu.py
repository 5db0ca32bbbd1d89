"""Size-independent properties of the CUDA path at BASELINE's full batch size (N_rand = 1024, 64c + 128f),
plus the less common switches (lindisp, white_bkgd, N_importance != N_samples, odd sample counts)."""
import numpy as np
import pytest
import torch

import oracle.nrnerf_oracle as O
from tests import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _kw(coarse, fine, bender, **over):
    kw = dict(network_query_fn=None, perturb=0.0, N_importance=64, network_fine=fine, N_samples=64, network_fn=coarse,
              ray_bender=bender, use_viewdirs=False, white_bkgd=False, raw_noise_std=0.0, ndc=False, lindisp=False)
    kw.update(over)
    return kw


def _render(r, kw, chunk=32768, detailed=True, latents=None):
    from nonrigid_nerf_b200 import train as T
    lat = r["latents"].to(DEV) if latents is None else latents
    return T.render(r["rays_o"].to(DEV), r["rays_d"].to(DEV), chunk=chunk, near=r["near"], far=r["far"],
                    additional_pixel_information={"ray_bending_latents": lat}, detailed_output=detailed, retraw=True, **kw)


def test_full_size_invariants_and_chunk_independence():
    from nonrigid_nerf_b200 import _lib
    seed, n = 31, 1024
    coarse, fine, bender, _ = helpers.build_models(O, seed, DEV)
    r = O.make_rays(seed, n)
    kw = _kw(coarse, fine, bender)
    with torch.no_grad():
        rgb, disp, acc, ex = _render(r, kw)
        rgb2, _, acc2, ex2 = _render(r, kw, chunk=100)          # ragged chunks: 10 x 100 + 24 rays
        rgb3, _, _, _ = _render(r, kw)                           # same call again
    _lib.device_error_check()
    assert torch.equal(rgb, rgb3), "forward must be deterministic"
    assert torch.equal(rgb, rgb2) and torch.equal(acc, acc2), "chunk must not change results (train.py:344-345)"
    assert torch.equal(ex["raw"], ex2["raw"])
    w, wf = ex["visibility_weights"], ex["fine_visibility_weights"]
    assert w.shape == (n, 64) and wf.shape == (n, 128)
    np.testing.assert_allclose(wf.sum(-1).cpu().numpy(), acc.cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert bool((acc >= 0).all()) and bool((acc <= 1 + 1e-5).all())
    assert bool((rgb >= 0).all()) and bool((rgb <= 1 + 1e-5).all())
    assert bool((ex["fine_opacity_alpha"] >= 0).all()) and bool((ex["fine_opacity_alpha"] <= 1).all())
    # the last interval is 1e10 long: the last sample's alpha is 0 or 1 (train.py:743-746)
    last = ex["fine_opacity_alpha"][:, -1]
    assert bool(((last == 0) | (last == 1)).all())
    # fine depths: sorted union containing every coarse depth
    z_f = (ex["fine_initial_input_pts"][..., 2] - r["rays_o"].to(DEV)[:, None, 2]) / r["rays_d"].to(DEV)[:, None, 2]
    assert bool((z_f[:, 1:] >= z_f[:, :-1] - 1e-6).all())
    # masked = rigidity * unmasked; bent = initial + masked (run_nerf_helpers.py:567-570)
    np.testing.assert_allclose(ex["masked_offsets"].cpu().numpy(), (ex["rigidity_mask"] * ex["unmasked_offsets"]).cpu().numpy(),
                               rtol=0, atol=1e-9)
    np.testing.assert_allclose(ex["input_pts"].cpu().numpy(), (ex["initial_input_pts"] + ex["masked_offsets"]).cpu().numpy(),
                               rtol=0, atol=1e-7)


def test_broadcast_latent_equals_materialised_and_canonical_ignores_latents():
    seed, n = 32, 300
    coarse, fine, bender, _ = helpers.build_models(O, seed, DEV)
    r = O.make_rays(seed, n)
    row = r["latents"][:1].to(DEV)
    kw = _kw(coarse, fine, bender)
    with torch.no_grad():
        a = _render(r, kw, latents=row.expand(n, 32))            # stride-0 view, like render_path (train.py:465)
        b = _render(r, kw, latents=row.expand(n, 32).contiguous())
    assert torch.equal(a[0], b[0])
    coarse.ray_bender = (None,); fine.ray_bender = (None,)        # canonical rendering (free_viewpoint_rendering.py:285-289)
    kw = _kw(coarse, fine, None)
    with torch.no_grad():
        c = _render(r, kw)
        d = _render(r, kw, latents=torch.zeros(n, 32, device=DEV))
    assert torch.equal(c[0], d[0])
    assert "unmasked_offsets" not in c[3] and "input_pts" in c[3]


@pytest.mark.parametrize("over", [dict(lindisp=True), dict(white_bkgd=True), dict(N_importance=32), dict(N_samples=48, N_importance=16),
                                  dict(N_samples=200, N_importance=56)])
def test_less_common_switches_match_oracle(over):
    seed, n = 33, 41
    coarse, fine, bender, (cp, fp, bp) = helpers.build_models(O, seed, DEV)
    r = O.make_rays(seed, n)
    kw = _kw(coarse, fine, bender, **over)
    with torch.no_grad():
        rgb, disp, acc, ex = _render(r, kw)
        ret = O.render_rays(cp, fp, bp, r["rays_o"], r["rays_d"], r["near"], r["far"], r["latents"], kw["N_samples"], kw["N_importance"],
                            lindisp=kw["lindisp"], white_bkgd=kw["white_bkgd"])
    d = (rgb.cpu() - ret["rgb_map"]).abs().max().item()
    print(f"{over}: rgb L-inf {d:.3e}")
    assert d <= 5e-4, (over, d)
    assert (acc.cpu() - ret["acc_map"]).abs().max().item() <= 5e-4
    assert ex["raw"].shape == (n, kw["N_samples"] + kw["N_importance"], 5)


def test_weight_gradients_are_bit_reproducible():
    """WGRAD reduces its split-K partials in a fixed order: two identical steps give identical weight gradients."""
    seed, n = 34, 256
    grads = []
    for _ in range(2):
        coarse, fine, bender, _ = helpers.build_models(O, seed, DEV)
        r = O.make_rays(seed, n)
        rnd = O.make_randomness(seed, n, 64, 64)
        kw = _kw(coarse, fine, bender, perturb=1.0, raw_noise_std=1.0, randomness=rnd)
        rgb, _, _, ex = _render(r, kw)
        tgt = r["target"].to(DEV)
        (((rgb - tgt) ** 2).mean() + ((ex["rgb0"] - tgt) ** 2).mean()).backward()
        grads.append([p.grad.clone() for p in list(coarse.parameters()) + list(fine.parameters()) if p.grad is not None])
    assert len(grads[0]) == 36
    for a, b in zip(*grads):
        assert torch.equal(a, b)

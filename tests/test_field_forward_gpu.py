"""GPU parity of the fused field kernel and the per-ray kernels against the oracle (same seeded inputs).
Tolerances (stated, fp16 tensor-core operands with fp32 accumulation vs the fp32 reference):
  raw logits  : |d| <= 2e-2 + 1e-2*|ref|  (typical 2e-3)
  bent points / offsets : |d| <= 1e-4 (typical 3e-5: 64-term fp16 dot products of O(1e-2) terms) ;  rigidity mask : |d| <= 3e-4
"""
import numpy as np
import pytest
import torch

import oracle.nrnerf_oracle as O
from tests import helpers

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("n,s,with_bender", [(64, 64, True), (37, 64, True), (16, 128, True), (50, 64, False), (3, 7, True)])
def test_field_forward_matches_oracle(n, s, with_bender):
    from nonrigid_nerf_b200 import autograd as ag, ops, _lib
    dev = _dev()
    seed = 1000 + n
    coarse, fine, bender, (cp, fp, bp) = helpers.build_models(O, seed, dev, with_bender)
    r = O.make_rays(seed, n)
    rays = helpers.rays8(r, dev)
    near = torch.full((n, 1), float(r["near"])); far = torch.full((n, 1), float(r["far"]))
    z_ref = O.stratified_z(near, far, s, None)
    z = z_ref.to(dev)   # same depths on both sides; sample_coarse has its own test
    raw, det = ag.field_rays(coarse, rays, z, r["latents"].to(dev) if with_bender else None, True)
    _lib.device_error_check()
    pts = r["rays_o"][:, None, :] + r["rays_d"][:, None, :] * z_ref[:, :, None]
    with torch.no_grad():
        raw_ref, det_ref = O.query_field(cp, bp, pts, r["latents"])
    assert torch.equal(det["initial_input_pts"].cpu(), det_ref["initial_input_pts"]), "sample points must be bit-exact"
    if with_bender:
        for k, tol in (("unmasked_offsets", 1e-4), ("masked_offsets", 1e-4), ("rigidity_mask", 3e-4), ("input_pts", 1e-4)):
            d = (det[k].cpu() - det_ref[k]).abs().max().item()
            print(f"  {k}: max abs err {d:.3e}")
            assert d <= tol, (k, d)
    d = (raw.cpu() - raw_ref).abs()
    tol = 2e-2 + 1e-2 * raw_ref.abs()
    assert bool((d <= tol).all()), f"raw mismatch: max abs {d.max().item():.3e}, mean {d.mean().item():.3e}"
    print(f"n={n} s={s} bender={with_bender}: raw max abs err {d.max().item():.3e} mean {d.mean().item():.3e}")


def test_field_forward_large_multi_tile():
    """More tiles than SMs, ragged tail, broadcast (stride-0) latent row."""
    from nonrigid_nerf_b200 import autograd as ag, ops, _lib
    dev = _dev()
    seed, n, s = 77, 1531, 64
    coarse, fine, bender, (cp, fp, bp) = helpers.build_models(O, seed, dev, True)
    r = O.make_rays(seed, n)
    rays = helpers.rays8(r, dev)
    z = ops.sample_coarse(rays, s, None, False)
    lat_row = r["latents"][:1]
    raw, _ = ag.field_rays(coarse, rays, z, lat_row.to(dev).expand(n, 32), False)
    _lib.device_error_check()
    idx = torch.arange(0, n, 97)
    pts = r["rays_o"][idx, None, :] + r["rays_d"][idx, None, :] * z.cpu()[idx][:, :, None]
    with torch.no_grad():
        raw_ref, _ = O.query_field(cp, bp, pts, lat_row.expand(len(idx), 32))
    d = (raw.cpu()[idx] - raw_ref).abs()
    assert bool((d <= 2e-2 + 1e-2 * raw_ref.abs()).all()), d.max().item()


def test_composite_and_resample_match_oracle():
    from nonrigid_nerf_b200 import ops
    dev = _dev()
    g = np.load("tests/golden/caseE_ops.npz")
    raw, z, rd = (torch.from_numpy(g[k]) for k in ("raw", "z", "rays_d"))
    out = ops.composite(raw.to(dev), z.to(dev), rd.to(dev))
    for k_out, k_ref in (("rgb_map", "rgb_map"), ("disp_map", "disp_map"), ("acc_map", "acc_map"), ("alpha", "alpha"),
                         ("weights", "weights_out"), ("depth_map", "depth_map")):
        np.testing.assert_allclose(out[k_out].cpu().numpy(), g[k_ref], atol=2e-6, rtol=2e-5, equal_nan=True, err_msg=k_out)
    outw = ops.composite(raw.to(dev), z.to(dev), rd.to(dev), white_bkgd=True)
    np.testing.assert_allclose(outw["rgb_map"].cpu().numpy(), g["rgb_map_white"], atol=2e-6, rtol=2e-5)
    bins, w = torch.from_numpy(g["bins"]), torch.from_numpy(g["weights"])
    # The inverse CDF is continuous in u except where the CDF is flat (zero-weight bins, and u == 1.0 against
    # a last CDF entry that rounds to 1 -/+ 1 ulp depending on the summation ORDER: torch's CPU cumsum is a
    # sequential double accumulation, torch's CUDA cumsum and this kernel are parallel fp32 scans).  So: exact
    # agreement everywhere except on a small set of such samples, which must still lie inside the bin range.
    def close_but_for_flat_spots(ours, ref, frac):
        bad = np.abs(ours - ref) > (3e-6 + 1e-5 * np.abs(ref))
        assert bad.mean() <= frac, bad.mean()
        assert np.all(ours >= g["bins"].min(-1, keepdims=True) - 1e-6) and np.all(ours <= g["bins"].max(-1, keepdims=True) + 1e-6)
    det = ops.sample_pdf_op(bins.to(dev), w.to(dev), 64, None)
    close_but_for_flat_spots(det.cpu().numpy(), g["samples_det"], 0.005)
    rnd = ops.sample_pdf_op(bins.to(dev), w.to(dev), 64, torch.from_numpy(g["u_rand"]).to(dev))
    close_but_for_flat_spots(rnd.cpu().numpy(), g["samples_rand"], 0.005)
    # fused composite + resample + merge vs the oracle chain
    rs = np.random.RandomState(3)
    u = torch.from_numpy(rs.uniform(0, 1, size=(raw.shape[0], 64)).astype(np.float32))
    for uu in (None, u):
        o = ops.composite(raw.to(dev), z.to(dev), rd.to(dev), n_importance=64, u=None if uu is None else uu.to(dev))
        wref = O.raw2outputs(raw, z, rd)[4]
        zs = O.sample_pdf(0.5 * (z[:, 1:] + z[:, :-1]), wref[:, 1:-1], uu if uu is not None else O.det_u(raw.shape[0], 64))
        zf = torch.sort(torch.cat([z, zs], -1), -1)[0]
        bad = np.abs(o["z_vals_out"].cpu().numpy() - zf.numpy()) > 5e-6 + 1e-5 * np.abs(zf.numpy())
        assert bad.mean() <= 0.005, bad.mean()
        np.testing.assert_allclose(o["z_std"].cpu().numpy(), torch.std(zs, -1, unbiased=False).numpy(), atol=2e-3, rtol=1e-3)
        assert bool((o["z_vals_out"][:, 1:] >= o["z_vals_out"][:, :-1]).all())


def test_composite_backward_matches_autograd():
    from nonrigid_nerf_b200 import ops
    dev = _dev()
    rs = np.random.RandomState(11)
    n, s = 33, 128
    raw = torch.from_numpy(rs.randn(n, s, 5).astype(np.float32) * 2).requires_grad_(True)
    z = torch.from_numpy(np.sort(rs.uniform(0.1, 2.0, size=(n, s)).astype(np.float32), -1))
    rd = torch.from_numpy(rs.randn(n, 3).astype(np.float32))
    noise = torch.from_numpy(rs.randn(n, s).astype(np.float32))
    g_rgb = torch.from_numpy(rs.randn(n, 3).astype(np.float32))
    rgb = O.raw2outputs(raw, z, rd, noise)[0]
    (rgb * g_rgb).sum().backward()
    d_raw = ops.composite_backward(raw.detach().to(dev), z.to(dev), rd.to(dev), noise.to(dev), False, g_rgb.to(dev))
    np.testing.assert_allclose(d_raw.cpu().numpy(), raw.grad.numpy(), atol=2e-6, rtol=2e-4)


def test_sample_coarse_matches_oracle():
    from nonrigid_nerf_b200 import ops
    dev = _dev()
    n, s = 129, 64
    r = O.make_rays(5, n)
    t_rand = torch.rand(n, s, generator=torch.Generator().manual_seed(1))
    # torch.linspace rounds differently per backend / CPU vector width (aten RangeFactoriesKernel), so the
    # depths agree to an ulp, not bit for bit
    z = ops.sample_coarse(helpers.rays8(r, dev), s, t_rand.to(dev), False)
    near = torch.full((n, 1), float(r["near"])); far = torch.full((n, 1), float(r["far"]))
    np.testing.assert_allclose(z.cpu().numpy(), O.stratified_z(near, far, s, t_rand).numpy(), atol=2.5e-7, rtol=0)
    z0 = ops.sample_coarse(helpers.rays8(r, dev), s, None, False)
    np.testing.assert_allclose(z0.cpu().numpy(), O.stratified_z(near, far, s, None).numpy(), atol=1.3e-7, rtol=0)
    zl = ops.sample_coarse(helpers.rays8(r, dev), s, None, True)
    np.testing.assert_allclose(zl.cpu().numpy(), O.stratified_z(near, far, s, None, lindisp=True).numpy(), rtol=2e-6)

"""nonrigid_nerf_b200.optim.Adam (one launch over a flat parameter buffer) against torch.optim.Adam."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _make(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(32,)] * 5 + [(64, 38), (64,), (3, 64), (256, 63), (256,), (256, 256), (256, 319), (5, 256), (5,), (4097,), (1,)]
    return [torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for s in shapes]


def test_matches_torch_adam_including_skipped_tensors_and_lr_changes():
    from nonrigid_nerf_b200 import optim, _lib
    ours = _make(3)
    ref = [p.detach().clone().requires_grad_(True) for p in ours]
    o1 = optim.Adam(ours, lr=5e-4, betas=(0.9, 0.999))
    o2 = torch.optim.Adam(ref, lr=5e-4, betas=(0.9, 0.999))
    assert all(p.data_ptr() >= o1._flat.data_ptr() for p in ours)       # parameters are views into the flat buffer
    g = torch.Generator().manual_seed(4)
    for it in range(7):
        for i, (a, b) in enumerate(zip(ours, ref)):
            if (i + it) % 5 == 0:               # no gradient this step: both optimizers must leave the tensor (and its moments) alone
                a.grad = b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g).to(DEV) * (10.0 ** ((i % 3) - 2))
            a.grad, b.grad = gr.clone(), gr.clone()
        lr = 5e-4 * (0.1 ** (it / 3.0))          # train.py:1611-1616 rewrites the learning rate every iteration
        for grp in o1.param_groups + o2.param_groups:
            grp["lr"] = lr
        o1.step()
        o2.step()
    _lib.device_error_check()
    for a, b in zip(ours, ref):
        torch.testing.assert_close(a.detach(), b.detach(), rtol=2e-5, atol=1e-7)
    # checkpoints interchange with torch.optim.Adam (train.py:682 loads, :1692 saves optimizer.state_dict())
    sd, sd_ref = o1.state_dict(), o2.state_dict()
    assert set(sd.keys()) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == list(range(len(ours)))
    steps = sorted(int(float(v["step"])) for v in sd["state"].values())
    assert steps[0] == 5 and steps[-1] == 6
    for i in sd_ref["state"]:
        assert int(float(sd["state"][i]["step"])) == int(float(sd_ref["state"][i]["step"]))
        m_ref, v_ref = sd_ref["state"][i]["exp_avg"], sd_ref["state"][i]["exp_avg_sq"]
        torch.testing.assert_close(sd["state"][i]["exp_avg"], m_ref, rtol=1e-4, atol=1e-6 * float(m_ref.abs().max()))
        torch.testing.assert_close(sd["state"][i]["exp_avg_sq"], v_ref, rtol=1e-4, atol=1e-6 * float(v_ref.abs().max()))
    # torch -> ours and ours -> torch, then one more identical step on both
    o3 = optim.Adam([p.detach().clone().requires_grad_(True) for p in ref], lr=5e-4)
    o3.load_state_dict(sd_ref)
    o4 = torch.optim.Adam([p.detach().clone().requires_grad_(True) for p in ours], lr=5e-4)
    o4.load_state_dict(sd)
    gr = [torch.randn(p.shape, generator=g).to(DEV) for p in ours]
    for opt in (o3, o4):
        for p, gi in zip(opt.param_groups[0]["params"], gr):
            p.grad = gi.clone()
        opt.step()
    for a, b in zip(o3.param_groups[0]["params"], o4.param_groups[0]["params"]):
        torch.testing.assert_close(a.detach(), b.detach(), rtol=3e-5, atol=1e-7)


def test_gradient_arena_zero_grad_and_in_place_accumulation():
    from nonrigid_nerf_b200 import optim
    ps = _make(5)
    opt = optim.Adam(ps, lr=1e-3)
    assert opt.grads_in_arena and all(p.grad is not None for p in ps)
    arena = opt.gradient_arena()
    (ps[5] ** 2).sum().backward()                    # ordinary autograd accumulates in place into the arena view
    assert opt.grads_in_arena and float(arena.abs().sum()) > 0
    torch.testing.assert_close(ps[5].grad, 2 * ps[5].detach())
    opt.zero_grad(set_to_none=True)                  # keeps the views, one memset
    assert opt.grads_in_arena and float(arena.abs().sum()) == 0
    ps[3].grad = None                                # a caller may still re-bind: step() falls back to the pointer table
    assert not opt.grads_in_arena
    before = ps[3].detach().clone()
    opt.step()
    assert torch.equal(ps[3].detach(), before)
    opt.zero_grad()
    assert opt.grads_in_arena


def test_training_trajectory_matches_torch_adam_and_refreshes_the_packed_weights():
    """The update happens behind autograd's version counters: the cached fp16 weight images must still follow.
    Two identically initialised model sets, one stepped by optim.Adam, one by torch.optim.Adam: same loss curve."""
    import oracle.nrnerf_oracle as O
    from tests import helpers
    from nonrigid_nerf_b200 import optim, train as T
    seed, n = 21, 64
    r = O.make_rays(seed, n)
    tgt = r["target"].to(DEV)
    curves = []
    for which in ("ours", "torch"):
        coarse, fine, bender, _ = helpers.build_models(O, seed, DEV)
        params = list(bender.parameters()) + list(coarse.parameters()) + list(fine.parameters())
        opt = optim.Adam(params, lr=5e-4) if which == "ours" else torch.optim.Adam(params, lr=5e-4)
        kw = dict(network_query_fn=None, perturb=0.0, N_importance=64, network_fine=fine, N_samples=64, network_fn=coarse, ray_bender=bender,
                  use_viewdirs=False, white_bkgd=False, raw_noise_std=0.0, ndc=False, lindisp=False)
        losses = []
        for _ in range(4):
            rgb = T.render(r["rays_o"].to(DEV), r["rays_d"].to(DEV), chunk=32768, near=r["near"], far=r["far"],
                           additional_pixel_information={"ray_bending_latents": r["latents"].to(DEV)}, **kw)[0]
            loss = ((rgb - tgt) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        assert all(p.grad is not None for p in bender.parameters())
        curves.append(losses)
    assert len(set(round(x, 9) for x in curves[0])) == 4, curves      # stale packed weights would repeat the first loss
    for a, b in zip(*curves):
        assert abs(a - b) <= 2e-3 * abs(b), curves                     # fp16 forward: the two runs differ only by Adam rounding

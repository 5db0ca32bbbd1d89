"""CPU-only tests: C-ABI surface, host-side mirror of the reference interface, ray-sharding plumbing
(gloo, world_size 2).  No kernel is launched here."""
import ctypes
import os
import re
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from nonrigid_nerf_b200 import _lib
    header = open(os.path.join(ROOT, "include", "nrnerf_b200.h")).read()
    declared = set(re.findall(r"\b(nrn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS.keys()), declared ^ set(_lib.SYMBOLS.keys())
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.nrn_abi_version() == _lib.ABI_VERSION
    assert loaded.nrn_packed_nerf_bytes() > 2 * 990_000 and loaded.nrn_packed_bender_bytes() > 2 * 53_000
    assert loaded.nrn_nerf_grad_floats(5) == 527_237 - 32_896 and loaded.nrn_bender_grad_floats() == 16_193
    assert loaded.nrn_stash_bytes(1024, 64) == 512 * 634_880
    assert loaded.nrn_stash_bytes(1, 7) == 2 * 634_880      # one ragged tile, rounded up to a tile pair (two slots per CTA)


def test_argument_validation_returns_error_codes_without_a_gpu():
    from nonrigid_nerf_b200 import _lib
    lib = _lib.load()
    assert lib.nrn_field_forward(None) == -1
    assert b"null args" in lib.nrn_last_error()
    assert lib.nrn_sample_pdf(None, None, None, 4, 1, 8, None, None) == -1
    assert lib.nrn_sample_coarse(None, None, -1, 64, 0, None, None) == -1
    a = _lib.NrnCompositeArgs()
    a.n_rays, a.n_samples, a.channels = 4, 64, 3
    assert lib.nrn_composite(ctypes.byref(a)) == -1
    assert lib.nrn_adam_step(None) == -1 and b"null args" in lib.nrn_last_error()
    ad = _lib.NrnAdamArgs()
    ad.n_tensors, ad.n_blocks = 3, 4
    assert lib.nrn_adam_step(ctypes.byref(ad)) == -1 and b"null buffer" in lib.nrn_last_error()
    assert lib.nrn_divergence_backward(None) == -1
    with pytest.raises(RuntimeError):
        _lib.check(-1, "unit test")


def test_every_entry_point_validates_before_touching_the_device():
    """Bad sizes / null pointers come back as NRN_E_INVALID with a message, empty inputs as NRN_OK -- all without a GPU
    (no CUDA call is made before validation)."""
    from nonrigid_nerf_b200 import _lib
    lib = _lib.load()
    bad = {
        "get_rays sizes": lambda: lib.nrn_get_rays(None, None, -1, 4, None, None, None),
        "get_rays null": lambda: lib.nrn_get_rays(None, None, 4, 4, None, None, None),
        "pack_rays size": lambda: lib.nrn_pack_rays(None, None, 0.0, 1.0, -3, None, None),
        "pack_rays null": lambda: lib.nrn_pack_rays(None, None, 0.0, 1.0, 3, None, None),
        "ray_batch sizes": lambda: lib.nrn_ray_batch(None, 4, None, None, None, None, 0, 8, None, None, None, None),
        "ray_batch null": lambda: lib.nrn_ray_batch(None, 4, None, None, None, None, 8, 8, None, None, None, None),
        "median sizes": lambda: lib.nrn_median_visibility_index(None, 4, 0, None, None),
        "median null": lambda: lib.nrn_median_visibility_index(None, 4, 64, None, None),
        "scale_rows sizes": lambda: lib.nrn_scale_rows(None, None, None, 4, 0, None),
        "scale_rows null": lambda: lib.nrn_scale_rows(None, None, None, 4, 3, None),
        "field_backward null": lambda: lib.nrn_field_backward(None),
        "composite_backward null": lambda: lib.nrn_composite_backward(None),
        "divergence_forward null": lambda: lib.nrn_divergence_forward(None),
        "ray_loss null": lambda: lib.nrn_ray_loss(None),
        "ray_loss_backward null": lambda: lib.nrn_ray_loss_backward(None),
        "peer_alloc null": lambda: lib.nrn_peer_alloc(4096, None, None),
        "peer_open null": lambda: lib.nrn_peer_open(None, None),
        "peer_reduce_adam null": lambda: lib.nrn_peer_reduce_adam(None, None),
        "peer_gather_rows null": lambda: lib.nrn_peer_gather_rows(None, None, 4, None, None),
    }
    for what, call in bad.items():
        assert call() == -1, what
        assert len(lib.nrn_last_error()) > 0, what
    empty = {
        "get_rays": lambda: lib.nrn_get_rays(None, None, 0, 4, None, None, None),
        "pack_rays": lambda: lib.nrn_pack_rays(None, None, 0.0, 1.0, 0, None, None),
        "ray_batch": lambda: lib.nrn_ray_batch(None, 0, None, None, None, None, 8, 8, None, None, None, None),
        "median": lambda: lib.nrn_median_visibility_index(None, 0, 64, None, None),
        "scale_rows": lambda: lib.nrn_scale_rows(None, None, None, 0, 3, None),
        "sample_coarse": lambda: lib.nrn_sample_coarse(None, None, 0, 64, 0, None, None),
        "peer_close": lambda: lib.nrn_peer_close(None),
        "peer_free": lambda: lib.nrn_peer_free(None),
    }
    for what, call in empty.items():
        assert call() == 0, what
    # struct-level checks
    rl = _lib.NrnRayLossArgs()
    rl.n_rays, rl.n_samples = 4, 0
    assert lib.nrn_ray_loss(ctypes.byref(rl)) == -1 and b"bad sizes" in lib.nrn_last_error()
    rl.n_samples = 64
    assert lib.nrn_ray_loss(ctypes.byref(rl)) == -1 and b"null argument" in lib.nrn_last_error()
    rb = _lib.NrnRayLossBwdArgs()
    rb.n_rays, rb.n_samples = 4, 64
    assert lib.nrn_ray_loss_backward(ctypes.byref(rb)) == -1 and b"upstream" in lib.nrn_last_error()
    ctx = _lib.NrnPeerCtx()
    ctx.world, ctx.rank = 9, 0
    assert lib.nrn_peer_reduce_adam(ctypes.byref(ctx), None) == -1 and b"at most" in lib.nrn_last_error()
    ctx.world, ctx.rank = 2, 2
    assert lib.nrn_peer_gather_rows(ctypes.byref(ctx), None, 4, None, None) == -1
    # the window layout the Python side assumes (peer.py): 1 KB of flags, two slots, then the arena; 256-byte granules
    assert lib.nrn_peer_window_bytes(0, 0) == 1024
    assert lib.nrn_peer_window_bytes(100, 10) == 1024 + 2 * 256 + 512
    assert lib.nrn_peer_window_bytes(-1, 10) == 0


def test_state_dict_keys_match_the_reference_checkpoint_layout():
    from nonrigid_nerf_b200 import run_nerf_helpers as H
    embed_fn, ch = H.get_embedder(10, 0)
    assert ch == 63
    bender = H.ray_bending(ch, 32, "simple_neural", embed_fn)
    net = H.NeRF(D=8, W=256, input_ch=ch, output_ch=5, skips=[4], input_ch_views=0, use_viewdirs=False, ray_bender=bender,
                 ray_bending_latent_size=32)
    keys = list(net.state_dict().keys())
    expect = [f"pts_linears.{i}.{k}" for i in range(8) for k in ("weight", "bias")] + \
             ["views_linears.0.weight", "views_linears.0.bias", "output_linear.weight", "output_linear.bias"]
    assert keys == expect
    assert net.pts_linears[5].weight.shape == (256, 319) and net.output_linear.weight.shape == (5, 256)
    assert sum(p.numel() for p in net.parameters()) == 527_237
    bk = list(bender.state_dict().keys())
    assert bk == [f"network.{i}.{k}" for i in range(4) for k in ("weight", "bias")] + ["network.4.weight"] + \
        [f"rigidity_network.{i}.{k}" for i in range(3) for k in ("weight", "bias")]
    assert sum(p.numel() for p in bender.parameters()) == 16_193
    assert isinstance(net.ray_bender, tuple) and net.ray_bender[0] is bender       # 1-tuple API (run_nerf_helpers.py:213)
    assert not any(p is q for p in net.parameters() for q in bender.parameters())
    assert float(bender.network[4].weight.abs().sum()) == 0.0                      # straight rays at init
    assert net.test_time_nonrigid_object_removal_threshold is None and bender.rigidity_test_time_cutoff is None


def test_unsupported_configurations_raise_loudly():
    from nonrigid_nerf_b200 import run_nerf_helpers as H, train as T
    with pytest.raises(RuntimeError, match="use_viewdirs"):
        H.NeRF(D=8, W=256, input_ch=63, use_viewdirs=True)
    with pytest.raises(RuntimeError, match="time_conditioned"):
        H.NeRF(D=8, W=256, input_ch=63, time_conditioned_baseline=True)
    with pytest.raises(RuntimeError):
        H.NeRF(D=4, W=128, input_ch=63)
    with pytest.raises(RuntimeError):
        H.get_embedder(10, -1)
    o, d = torch.zeros(4, 3), torch.ones(4, 3)
    with pytest.raises(RuntimeError, match="not implemented. change H, W, focal"):     # same message as train.py:386
        T.render(o, d, ndc=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        T.render(o, d, ndc=False, additional_pixel_information={"ray_bending_latents": torch.zeros(4, 32)})
    with pytest.raises(RuntimeError, match="pytest"):
        T.raw2outputs(torch.zeros(2, 4, 5), torch.zeros(2, 4), torch.ones(2, 3), pytest=True)
    from nonrigid_nerf_b200 import optim
    with pytest.raises(RuntimeError, match="CUDA"):
        optim.Adam([torch.zeros(3, requires_grad=True)], lr=1e-3)          # no CPU path
    with pytest.raises(RuntimeError, match="flat list"):
        optim.Adam([{"params": [torch.zeros(3, requires_grad=True)]}], lr=1e-3)


def test_shard_bounds_follow_dataparallel_chunking():
    from nonrigid_nerf_b200.parallel import shard_bounds
    assert [shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [shard_bounds(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(1024, 8, 7) == (896, 1024)


# ---- world_size-2 gloo test of the DataParallel replacement -------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _ToyStep(torch.nn.Module):
    """Stand-in for training_wrapper_class: per-ray loss from rays, a nested dict tensor and a shared net."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, rays, info, scale, target):
        pred = self.net(torch.cat([rays, info["lat"]], -1))
        return ((pred - target) ** 2).mean(-1) * scale


def _worker(rank, world, port, n, out_path, uniform):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from nonrigid_nerf_b200 import parallel as P
    P.UNIFORM_GRADS = uniform
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    dead = torch.nn.Parameter(torch.zeros(3))            # never used: its grad must stay None on every rank
    opt = torch.optim.Adam(list(net.parameters()) + [dead], lr=1e-2)
    P._register_sharded_params(list(net.parameters()) + [dead])   # what get_parallelized_training_function does
    P._install_optimizer_hook()
    fn = P.RayShardedFunction(_ToyStep(net))
    # an unrelated optimizer in the same process (bench.py's CPU baseline, a user's second model) must not be drawn into
    # the collective: only rank 0 steps it -- if the hook all-reduced its gradients, rank 1 would never answer
    other = torch.nn.Linear(3, 1)
    other_opt = torch.optim.Adam(other.parameters(), lr=1e-2)
    if rank == 0:
        other(torch.ones(2, 3)).sum().backward()
        other_opt.step()
    g = torch.Generator().manual_seed(7 + rank)          # ranks draw DIFFERENT batches; rank 0's must win
    for it in range(3):
        rays, lat, tgt = torch.randn(n, 3, generator=g), torch.randn(n, 2, generator=g), torch.randn(n, 3, generator=g)
        losses = fn(rays, {"lat": lat}, 2.0, tgt)
        assert losses.shape == (n,)
        opt.zero_grad()
        losses.mean().backward()
        opt.step()
    assert dead.grad is None
    if rank == 0:
        torch.save({"w": net[0].weight.detach().clone(), "losses": losses.detach().clone()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("uniform", [False, True])
def test_ray_sharded_function_matches_single_process(tmp_path, uniform):
    n, world = 11, 2          # uneven shards: 6 + 5 rows
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), n, out, uniform), nprocs=world, join=True)
    got = torch.load(out)
    # single-process run on rank 0's batches
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    step = _ToyStep(net)
    g = torch.Generator().manual_seed(7)
    for it in range(3):
        rays, lat, tgt = torch.randn(n, 3, generator=g), torch.randn(n, 2, generator=g), torch.randn(n, 3, generator=g)
        losses = step(rays, {"lat": lat}, 2.0, tgt)
        opt.zero_grad()
        losses.mean().backward()
        opt.step()
    torch.testing.assert_close(got["losses"], losses.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["w"], net[0].weight.detach(), rtol=1e-5, atol=1e-6)


class _ArenaSGD(torch.optim.Optimizer):
    """CPU stand-in for optim.Adam's multi-GPU contract: one flat gradient arena whose views are the .grad tensors,
    `gradient_arena()`, `reduces_gradients_itself`, and a step() that lets the attached reducer sum the arena in place."""

    def __init__(self, params, lr):
        params = list(params)
        super().__init__(params, {"lr": lr})
        self._gflat = torch.zeros(sum(p.numel() for p in params))
        self._reducer, self.grads_in_arena, o = None, True, 0
        for p in params:
            p.grad = self._gflat[o:o + p.numel()].view_as(p)
            o += p.numel()

    @property
    def reduces_gradients_itself(self):
        return self._reducer is not None and self.grads_in_arena

    def gradient_arena(self):
        return self._gflat

    def zero_grad(self, set_to_none=False):
        self._gflat.zero_()

    def step(self):
        if self.reduces_gradients_itself:
            assert self._reducer.step(self, None) is False      # the NCCL-style reducer leaves the update to the optimizer
        for g in self.param_groups:
            for p in g["params"]:
                p.data.add_(p.grad, alpha=-g["lr"])


def _arena_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from nonrigid_nerf_b200 import parallel as P
    torch.manual_seed(0)
    net = torch.nn.Linear(5, 3)
    opt = _ArenaSGD(net.parameters(), lr=0.1)
    P._register_sharded_params(list(net.parameters()))
    P._install_optimizer_hook()
    fn = P.RayShardedFunction(_ToyStep(net))
    g = torch.Generator().manual_seed(11)
    arenas = []
    for it in range(2):
        rays, lat, tgt = torch.randn(9, 3, generator=g), torch.randn(9, 2, generator=g), torch.randn(9, 3, generator=g)
        opt.zero_grad()
        fn(rays, {"lat": lat}, 1.0, tgt).mean().backward()
        assert net.weight.grad.data_ptr() == opt.gradient_arena().data_ptr()      # autograd accumulated INTO the arena views
        opt.step()
        arenas.append(opt.gradient_arena().clone())
    assert isinstance(opt._reducer, P.NcclArenaReducer)          # attached by the hook on the first step
    if rank == 0:
        torch.save({"arenas": arenas, "w": net.weight.detach().clone()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_arena_is_reduced_exactly_once_per_step(tmp_path):
    """The optimizer hook hands an arena optimizer its reducer on the first step and then stays out of the way: the arena
    is summed across ranks ONCE (not by the hook and again by step()), from the very first iteration on."""
    out = str(tmp_path / "arena.pt")
    mp.spawn(_arena_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    net = torch.nn.Linear(5, 3)
    step = _ToyStep(net)
    g = torch.Generator().manual_seed(11)
    for it in range(2):
        rays, lat, tgt = torch.randn(9, 3, generator=g), torch.randn(9, 2, generator=g), torch.randn(9, 3, generator=g)
        net.zero_grad()
        step(rays, {"lat": lat}, 1.0, tgt).mean().backward()
        full = torch.cat([net.weight.grad.reshape(-1), net.bias.grad.reshape(-1)])
        torch.testing.assert_close(got["arenas"][it], full, rtol=1e-5, atol=1e-6)
        with torch.no_grad():
            for p in net.parameters():
                p.add_(p.grad, alpha=-0.1)
    torch.testing.assert_close(got["w"], net.weight.detach(), rtol=1e-5, atol=1e-6)


def test_single_process_wrappers_call_straight_through():
    from nonrigid_nerf_b200 import parallel as P
    net = torch.nn.Linear(5, 3)
    fn = P.RayShardedFunction(_ToyStep(net))
    out = fn(torch.randn(4, 3), {"lat": torch.randn(4, 2)}, 1.0, torch.randn(4, 3))
    assert out.shape == (4,)

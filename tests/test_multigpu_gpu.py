"""Ray-sharded training over NCCL on 2 GPUs (skipped on a 1-GPU box): the DataParallel replacement of train.py:290-297.

Each rank renders its row block of the SAME global batch; the per-ray losses are all-gathered, every rank back-propagates
the global mean, the optimizer's gradient arena is summed across ranks in place and Adam steps replicated.  Checked:
  * the gathered per-ray loss [N] equals a single-GPU run bit for bit (rays are independent) and the executed reference
    (golden case H) within the forward tolerance;
  * the reduced gradients equal the single-GPU gradients up to summation order, and the reference's within fp16 tolerance;
  * after the optimizer step the weights are identical on both ranks (replicated Adam needs no broadcast);
  * an unrelated optimizer stepping on one rank only is not drawn into the collective.
Both reducers are exercised: NCCL all-reduce over the arena, and the peer-memory kernel (csrc/peer.cu) that sums the
ranks' arenas over NVLink inside the optimizer launch.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

import oracle.nrnerf_oracle as O
from tests import helpers

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev, g):
    import types
    from nonrigid_nerf_b200 import optim
    seed, n = int(g["seed"]), int(g["n"])
    torch.manual_seed(seed)      # the dead views_linears are default-initialised from the global generator: same on every rank
    coarse, fine, bender, _ = helpers.build_models(O, seed, dev)
    r = O.make_rays(seed, n)
    # per-ray random draws on the device, leading dimension = rays: RayShardedFunction shards them with their rays
    rnd = {k: v.to(dev) for k, v in O.make_randomness(seed, n, 64, 64).items()}
    rnd["e"] = torch.from_numpy(g["e"]).to(dev).view(n, 64, 3)
    latents = [torch.from_numpy(row.copy()).to(dev).requires_grad_(True) for row in g["latent_table"]]
    params = latents + list(bender.parameters()) + list(coarse.parameters()) + list(fine.parameters())
    opt = optim.Adam(params, lr=5e-4)
    targs = types.SimpleNamespace(chunk=32768, N_samples=64, N_importance=64, N_iters=int(g["N_iters"]),
                                  offsets_loss_weight=float(g["offsets_w"]), divergence_loss_weight=float(g["divergence_w"]),
                                  rigidity_loss_weight=float(g["rigidity_w"]), ray_bending_latent_size=32)
    kw = {"network_query_fn": None, "perturb": 1.0, "N_importance": 64, "network_fine": fine, "N_samples": 64, "network_fn": coarse,
          "ray_bender": bender, "use_viewdirs": False, "white_bkgd": False, "raw_noise_std": 1.0, "ndc": False, "lindisp": False,
          "near": r["near"], "far": r["far"], "randomness": rnd}
    call = (targs, r["rays_o"].to(dev), r["rays_d"].to(dev), 100, kw, r["target"].to(dev), int(g["global_step"]), 0,
            {"imageid_to_timestepid": [int(v) for v in g["i2t"]]}, torch.from_numpy(g["pix"]).to(dev))
    return coarse, fine, bender, latents, opt, call


def _step(train_fn, opt, call):
    opt.zero_grad()
    losses = train_fn(*call)
    losses.mean().backward()
    grads = opt.gradient_arena().clone()
    opt.step()
    return losses.detach(), grads


def _worker(rank, world, port, out_path, reducer_kind):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from nonrigid_nerf_b200 import _lib, parallel
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = np.load(os.path.join(GOLD, "caseH_training_wrapper.npz"))
    coarse, fine, bender, latents, opt, call = _build(dev, g)
    train_fn = parallel.get_parallelized_training_function(coarse, latents, fine_model=fine, ray_bender=bender)
    if reducer_kind == "peer":
        from nonrigid_nerf_b200 import peer
        parallel.attach_optimizer(opt, peer.PeerArenaReducer(opt))
    other = torch.nn.Linear(3, 1)                      # an unrelated CPU optimizer on rank 0 only must pass through untouched
    other_opt = torch.optim.Adam(other.parameters(), lr=1e-2)
    if rank == 0:
        other(torch.ones(2, 3)).sum().backward()
        other_opt.step()
    def across_ranks(t):
        """max |t_rank - t_rank0| over ranks (0.0 = bit-identical replicas)"""
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return max(float((q - parts[0]).abs().max()) for q in parts[1:])

    diag = {"params_before": across_ranks(opt._flat)}
    losses, _ = _step(train_fn, opt, call)
    # the arena after step() holds the all-reduced gradient
    reduced = opt.gradient_arena().clone()
    diag["reduced_grad_step1"] = across_ranks(reduced)
    diag["params_step1"] = across_ranks(opt._flat)
    diag["reducer"] = type(getattr(opt, "_reducer", None)).__name__
    diag["grads_in_arena"] = bool(opt.grads_in_arena)
    losses2, _ = _step(train_fn, opt, call)            # a second step: the collective is re-entrant
    _lib.device_error_check()
    diag["params_step2"] = across_ranks(opt._flat)
    if rank == 0:
        torch.save({"losses": losses.cpu(), "losses2": losses2.cpu(), "reduced": reduced.cpu(), "diag": diag,
                    "replicas_equal": diag["params_step1"] == 0.0 and diag["params_step2"] == 0.0}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("reducer_kind", ["nccl", "peer"])
def test_two_rank_training_step_equals_single_rank_and_reference(tmp_path, reducer_kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from nonrigid_nerf_b200 import parallel
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, reducer_kind), nprocs=2, join=True)
    got = torch.load(out)
    # single-GPU run of the same two steps
    g = np.load(os.path.join(GOLD, "caseH_training_wrapper.npz"))
    dev = torch.device("cuda", 0)
    coarse, fine, bender, latents, opt, call = _build(dev, g)
    train_fn = parallel.get_parallelized_training_function(coarse, latents, fine_model=fine, ray_bender=bender)
    losses, grads = _step(train_fn, opt, call)
    losses2, _ = _step(train_fn, opt, call)
    print(f"[{reducer_kind}] across-rank diagnostics: {got['diag']}")
    assert got["replicas_equal"], f"replicated Adam diverged between ranks: {got['diag']}"
    assert torch.equal(got["losses"], losses.cpu()), float((got["losses"] - losses.cpu()).abs().max())
    rel = float((got["reduced"] - grads.cpu()).norm() / grads.cpu().norm())
    print(f"[{reducer_kind}] reduced gradient vs single GPU: rel {rel:.3e}")
    assert rel <= 1e-5, rel
    d = float((got["losses"].numpy() - g["loss"]).__abs__().max())
    print(f"[{reducer_kind}] sharded per-ray loss vs executed reference: L-inf {d:.3e}")
    assert d <= 2e-3, d
    rel2 = float((got["losses2"] - losses2.cpu()).norm() / losses2.cpu().norm())
    assert rel2 <= 1e-4, rel2          # second step: weights went through one (order-dependent) reduced Adam update

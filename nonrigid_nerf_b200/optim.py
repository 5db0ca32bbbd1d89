"""Optimizer of the training step: a drop-in for `torch.optim.Adam(params=grad_vars, lr=..., betas=(0.9, 0.999))`
(train.py:656-658; stepped at train.py:1610, learning rate rewritten per iteration at train.py:1631-1642).

Memory layout (all fp32, one allocation each, in the order the parameters were given):
  * parameters  -- every nn.Parameter becomes a view into ONE flat buffer (names, shapes, state_dict untouched)
  * gradients   -- ONE flat arena; every `p.grad` is a view into it ("gradients are written /
                   accumulated into fp32 buffers that alias param.grad").  The fused backward kernels (WGRAD reduce,
                   latent scatter) add straight into the arena, `zero_grad()` is one memset, the multi-GPU gradient
                   all-reduce runs IN PLACE over the arena (no gather / scatter copies), and the Adam kernel reads it.
  * exp_avg, exp_avg_sq -- flat, same offsets.
One CUDA launch updates all tensors (csrc/adam.cu) instead of PyTorch's six multi-tensor launches for this model's
137 tensors.  Step counts and the learning rate are device scalars, so a whole iteration can be captured in a CUDA graph.

A caller may still assign its own gradient tensors (or None = skip, like torch.optim.Adam): step() then reads them
through a freshly uploaded pointer table.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import numpy as np
import torch

from . import _lib, ops

_BLOCK_ELEMS = 2048   # csrc/adam.cuh: kAdamBlockElems


class Adam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, amsgrad: bool = False):
        params = list(params)
        if not params or any(isinstance(p, dict) for p in params):
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam takes one flat list of tensors (grad_vars, train.py:650-658)")
        if weight_decay != 0.0 or amsgrad:
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam: weight_decay / amsgrad are not implemented (the reference uses neither)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=True, differentiable=False, fused=None))
        ps: List[torch.Tensor] = self.param_groups[0]["params"]
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam: parameters must live on a CUDA device (no CPU path)")
        for p in ps:
            if p.device != dev or p.dtype != torch.float32:
                raise RuntimeError("nonrigid_nerf_b200.optim.Adam: all parameters must be float32 tensors on one device")
        self._dev = dev
        sizes = [p.numel() for p in ps]
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        total = int(offs[-1])
        if total >= 2 ** 31:
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam: more than 2^31 parameters")
        self._offs, self._sizes, self._total = offs, sizes, total
        with torch.no_grad():
            self._flat = torch.empty(total, dtype=torch.float32, device=dev)
            for p, o, n in zip(ps, offs, sizes):
                self._flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self._flat[o:o + n].view(p.shape)      # the parameter is now a view into the flat buffer
        self._pviews_ptr = [p.data_ptr() for p in ps]
        self._m = torch.zeros(total, dtype=torch.float32, device=dev)
        self._v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._gflat = self._allocate_arena(total)
        self._gviews = [self._gflat[o:o + n].view(p.shape) for p, o, n in zip(ps, offs, sizes)]
        self._seat_grads()
        blocks = []
        for i, (o, n) in enumerate(zip(offs, sizes)):
            for s in range(0, n, _BLOCK_ELEMS):
                blocks.append((i, s, min(_BLOCK_ELEMS, n - s), int(o) + s))
        self._n_blocks = len(blocks)
        self._blocks = torch.tensor(np.asarray(blocks, dtype=np.int32).reshape(-1, 4), device=dev)
        self._step = torch.zeros(len(ps), dtype=torch.int64, device=dev)   # per tensor, like torch.optim.Adam
        self._lr_dev = torch.full((), float(lr), dtype=torch.float32, device=dev)
        self._lr_pushed = float(lr)
        self._arena_ptrs = [v.data_ptr() for v in self._gviews]
        self._ptr_arena_dev = torch.tensor(self._arena_ptrs, dtype=torch.int64, device=dev)   # never rewritten
        self._ptr_dev = torch.zeros(len(ps), dtype=torch.int64, device=dev)                   # custom / None gradients
        self._graph_hosts = []      # pinned pointer tables referenced by captured copies: kept alive, never rewritten
        self._reducer = None        # multi-GPU: object with .step(optimizer, adam_args) -> bool (parallel.py, peer.py)

    # ------------------------------------------------------------------------------------------------
    def _allocate_arena(self, total: int) -> torch.Tensor:
        return torch.zeros(total, dtype=torch.float32, device=self._dev)

    def _seat_grads(self) -> None:
        for p, v in zip(self.param_groups[0]["params"], self._gviews):
            if p.grad is not v:
                p.grad = v

    @property
    def grads_in_arena(self) -> bool:
        """True when every parameter's .grad is its view of the flat arena (the state zero_grad() establishes)."""
        return all(p.grad is v for p, v in zip(self.param_groups[0]["params"], self._gviews))

    @property
    def reduces_gradients_itself(self) -> bool:
        """parallel.py's generic optimizer hook skips this optimizer when it all-reduces its arena in step()."""
        return self._reducer is not None and self.grads_in_arena

    def gradient_arena(self) -> torch.Tensor:
        return self._gflat

    def rebind_arena(self, arena: torch.Tensor) -> None:
        """Move the gradient arena into caller-provided memory (peer.PeerArenaReducer: a CUDA-IPC window the other ranks of
        the node can read).  Current gradient values are carried over."""
        if arena.numel() != self._total or arena.dtype != torch.float32 or arena.device != self._dev or not arena.is_contiguous():
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam.rebind_arena: need a contiguous fp32 tensor of the arena's size on the same device")
        ps = self.param_groups[0]["params"]
        was_seated = self.grads_in_arena
        with torch.no_grad():
            arena.copy_(self._gflat)
        self._gflat = arena
        self._gviews = [arena[o:o + n].view(p.shape) for p, o, n in zip(ps, self._offs, self._sizes)]
        self._arena_ptrs = [v.data_ptr() for v in self._gviews]
        self._ptr_arena_dev = torch.tensor(self._arena_ptrs, dtype=torch.int64, device=self._dev)
        if was_seated:
            self._seat_grads()

    def zero_grad(self, set_to_none: bool = True) -> None:
        """One memset over the arena.  `set_to_none` is accepted for signature compatibility: gradients stay tensors
        (views of the arena) because the backward kernels accumulate into them in place.  For this model the result
        equals torch's None semantics: a tensor whose gradient stays zero from the start (the dead views_linears,
        SURVEY.md 7.3-6) has zero moments and is never moved."""
        self._gflat.zero_()
        self._seat_grads()

    # -- checkpoints in torch.optim.Adam's format (train.py:682 / :1692 load, :1648-1650 save) ---------
    def state_dict(self):
        ps = self.param_groups[0]["params"]
        steps = self._step.tolist()
        state = {}
        for i, (o, n, p) in enumerate(zip(self._offs, self._sizes, ps)):
            if steps[i] == 0:
                continue    # torch creates the state lazily at a tensor's first step
            state[i] = {"step": torch.tensor(float(steps[i])), "exp_avg": self._m[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": self._v[o:o + n].view(p.shape).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(ps)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        ps = self.param_groups[0]["params"]
        if "state" not in sd or "param_groups" not in sd:
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam.load_state_dict: expected torch.optim.Adam's format "
                               "({'state': {index: {step, exp_avg, exp_avg_sq}}, 'param_groups': [...]})")
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(ps):
            raise ValueError("loaded state dict contains a different number of parameters")
        steps = [0] * len(ps)
        self._m.zero_()
        self._v.zero_()
        for pos, idx in enumerate(order):
            st = sd["state"].get(idx, sd["state"].get(str(idx)))
            if st is None:
                continue
            o, n = int(self._offs[pos]), self._sizes[pos]
            if st["exp_avg"].numel() != n:
                raise ValueError(f"loaded optimizer state of parameter {idx} has {st['exp_avg'].numel()} elements, expected {n}")
            self._m[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self._v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps[pos] = int(float(st["step"]))
        self._step.copy_(torch.tensor(steps, dtype=torch.int64))
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v

    # ------------------------------------------------------------------------------------------------
    def _push_lr(self, lr: float) -> None:
        self._lr_dev.fill_(lr)          # the value travels as a kernel argument: no host staging buffer to race on
        self._lr_pushed = lr

    def set_lr(self, lr: float) -> None:
        """Push a new learning rate now (use between CUDA-graph replays; eager code can simply assign
        param_groups[0]['lr'] like train.py:1641-1642)."""
        self.param_groups[0]["lr"] = float(lr)
        self._push_lr(float(lr))

    def _grad_table(self, ps, capturing: bool) -> int:
        """Device address of the gradient pointer table for this step."""
        if self.grads_in_arena:
            return self._ptr_arena_dev.data_ptr()
        ptrs = []
        for p in ps:
            g = p.grad
            if g is None:
                ptrs.append(0)
                continue
            if g.dtype != torch.float32 or g.device != self._dev or g.is_sparse:
                raise RuntimeError("nonrigid_nerf_b200.optim.Adam: gradients must be dense float32 tensors on the parameters' device")
            if not g.is_contiguous():
                g = g.contiguous()
                p.grad = g
            ptrs.append(g.data_ptr())
        # a fresh pinned table per upload: PyTorch's pinned-memory allocator recycles it only after the queued copy
        # has executed, so a host that runs ahead can never rewrite a table the device has not read yet
        host = torch.tensor(ptrs, dtype=torch.int64).pin_memory()
        if capturing:
            self._graph_hosts.append(host)      # a replayed copy node reads this memory again: keep it, never touch it
        self._ptr_dev.copy_(host, non_blocking=True)
        return self._ptr_dev.data_ptr()

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam: closures are not supported")
        group = self.param_groups[0]
        ps = group["params"]
        for p, ptr in zip(ps, self._pviews_ptr):
            if p.data_ptr() != ptr:
                raise RuntimeError("nonrigid_nerf_b200.optim.Adam: a parameter no longer aliases the flat buffer (its .data was "
                                   "re-bound after the optimizer was built, e.g. by model.to()); rebuild the optimizer")
        capturing = torch.cuda.is_current_stream_capturing()
        lr = float(group["lr"])
        if lr != self._lr_pushed:
            if capturing:
                raise RuntimeError("nonrigid_nerf_b200.optim.Adam: change the learning rate outside the captured region")
            self._push_lr(lr)
        a = _lib.NrnAdamArgs()
        a.params, a.exp_avg, a.exp_avg_sq = self._flat.data_ptr(), self._m.data_ptr(), self._v.data_ptr()
        a.grad_ptrs, a.blocks, a.n_tensors, a.n_blocks = self._grad_table(ps, capturing), self._blocks.data_ptr(), len(ps), self._n_blocks
        a.lr, a.step = self._lr_dev.data_ptr(), self._step.data_ptr()
        b1, b2 = group["betas"]
        a.beta1, a.beta2, a.eps = float(b1), float(b2), float(group["eps"])
        a.stream = torch.cuda.current_stream().cuda_stream
        # multi-GPU: the reducer either sums the ranks' arenas in place (NCCL) and leaves the update to the launch below,
        # or does both in one fused sequence over peer memory (peer.PeerArenaReducer) and returns True
        done = self._reducer.step(self, a) if self.reduces_gradients_itself else False
        if not done:
            with torch.cuda.device(self._dev):
                _lib.check(_lib.load().nrn_adam_step(C.byref(a)), "adam_step")
        ops.note_parameters_changed()
        return None


def arena_destination(params: List[torch.Tensor]) -> Optional[int]:
    """If the .grad tensors of `params` lie back to back (in this order) in one fp32 buffer -- the layout
    optim.Adam's arena gives a module's parameters -- return the address of the first one, else None."""
    if not params:
        return None
    g0 = params[0].grad
    if g0 is None:
        return None
    expect = g0.data_ptr()
    for p in params:
        g = p.grad
        if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != expect or g.shape != p.shape:
            return None
        expect += 4 * g.numel()
    return g0.data_ptr()

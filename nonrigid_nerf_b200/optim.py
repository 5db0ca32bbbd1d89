"""Optimizer of the training step: a drop-in for `torch.optim.Adam(params=grad_vars, lr=..., betas=(0.9, 0.999))`
(train.py:656-658; stepped at train.py:1608, learning rate rewritten per iteration at train.py:1611-1616).

All parameters are moved into ONE flat fp32 buffer (every nn.Parameter becomes a view into it: names, shapes and
state_dict layout are untouched) and one CUDA launch updates all of them (csrc/adam.cu) instead of PyTorch's six
multi-tensor launches for this model's 137 tensors.  Gradients are read where autograd left them, through a
device table of pointers.  Step count and learning rate are device scalars, so the iteration can be captured in a
CUDA graph; `param_groups[0]["lr"] = x` keeps working (the value is pushed to the device by step()).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List

import numpy as np
import torch

from . import _lib, ops

_BLOCK_ELEMS = 2048   # csrc/adam.cuh: kAdamBlockElems


class Adam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        params = list(params)
        if not params or any(isinstance(p, dict) for p in params):
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam takes one flat list of tensors (grad_vars, train.py:650-658)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
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
        with torch.no_grad():
            self._flat = torch.empty(total, dtype=torch.float32, device=dev)
            for p, o, n in zip(ps, offs, sizes):
                self._flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self._flat[o:o + n].view(p.shape)      # the parameter is now a view into the flat buffer
        self._m = torch.zeros(total, dtype=torch.float32, device=dev)
        self._v = torch.zeros(total, dtype=torch.float32, device=dev)
        blocks = []
        for i, (o, n) in enumerate(zip(offs, sizes)):
            for s in range(0, n, _BLOCK_ELEMS):
                blocks.append((i, s, min(_BLOCK_ELEMS, n - s), int(o) + s))
        self._n_blocks = len(blocks)
        self._blocks = torch.tensor(np.asarray(blocks, dtype=np.int32).reshape(-1, 4), device=dev)
        self._step = torch.zeros(len(ps), dtype=torch.int64, device=dev)   # per tensor, like torch.optim.Adam
        self._lr_host = torch.empty((), dtype=torch.float32).pin_memory()
        self._lr_host.fill_(float(lr))
        self._lr_dev = torch.full((), float(lr), dtype=torch.float32, device=dev)
        self._lr_pushed = float(lr)
        self._ptr_host = torch.zeros(len(ps), dtype=torch.int64).pin_memory()
        self._ptr_dev = torch.zeros(len(ps), dtype=torch.int64, device=dev)
        self._ptr_last = None

    # -- state the reference's checkpoint code touches (train.py:1648-1650 saves optimizer.state_dict()) --
    def state_dict(self):
        return {"step": self._step.clone(), "exp_avg": self._m.clone(), "exp_avg_sq": self._v.clone(),
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        self._step.copy_(sd["step"])
        self._m.copy_(sd["exp_avg"])
        self._v.copy_(sd["exp_avg_sq"])
        for k, v in sd["param_groups"][0].items():
            self.param_groups[0][k] = v

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise RuntimeError("nonrigid_nerf_b200.optim.Adam: closures are not supported")
        group = self.param_groups[0]
        ps = group["params"]
        capturing = torch.cuda.is_current_stream_capturing()
        lr = float(group["lr"])
        if lr != self._lr_pushed:
            if capturing:
                raise RuntimeError("nonrigid_nerf_b200.optim.Adam: change the learning rate outside the captured region")
            self._lr_host.fill_(lr)
            self._lr_dev.copy_(self._lr_host, non_blocking=True)
            self._lr_pushed = lr
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
        if ptrs != self._ptr_last:
            self._ptr_host.copy_(torch.tensor(ptrs, dtype=torch.int64))
            self._ptr_dev.copy_(self._ptr_host, non_blocking=True)     # replayed with the graph if captured
            self._ptr_last = None if capturing else ptrs
        a = _lib.NrnAdamArgs()
        a.params, a.exp_avg, a.exp_avg_sq = self._flat.data_ptr(), self._m.data_ptr(), self._v.data_ptr()
        a.grad_ptrs, a.blocks, a.n_tensors, a.n_blocks = self._ptr_dev.data_ptr(), self._blocks.data_ptr(), len(ps), self._n_blocks
        a.lr, a.step = self._lr_dev.data_ptr(), self._step.data_ptr()
        b1, b2 = group["betas"]
        a.beta1, a.beta2, a.eps = float(b1), float(b2), float(group["eps"])
        a.stream = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(self._dev):
            _lib.check(_lib.load().nrn_adam_step(C.byref(a)), "adam_step")
        ops.note_parameters_changed()
        return None

    def set_lr(self, lr: float) -> None:
        """Push a new learning rate now (use between CUDA-graph replays; eager code can simply assign
        param_groups[0]['lr'] like train.py:1614-1616)."""
        self.param_groups[0]["lr"] = float(lr)
        self._lr_host.fill_(float(lr))
        self._lr_dev.copy_(self._lr_host, non_blocking=True)
        self._lr_pushed = float(lr)

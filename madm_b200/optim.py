"""Optimizer side of MADM's training step (what follows the backward pass) on the device, behind the C ABI
(``madm_op_grad_norm`` / ``madm_op_adamw_step`` / ``madm_op_ema_update``, ``csrc/optim.cu``), plus the one collective of the training
configuration: the all-reduce of the trainable gradients.

Mirrors, with the same names and argument meaning:

* ``torch.optim.AdamW`` as ``config_files/common/optim.py:9-18`` instantiates it (param groups with ``lr`` / ``weight_decay``),
* ``torch.nn.utils.clip_grad_norm_`` as ``engine/train_loop.py:123-124, :201-210`` call it on the optimizer's parameters
  (folded into the step: the norm stays a device scalar, there is no ``.item()``),
* ``CMDISE._update_ema`` (``modeling/meta_arch/cmdise.py:337-349``),
* DDP's gradient averaging (``main.py:290``) as ONE all-reduce over a flat buffer, with zero gradients materialised for the
  parameters that did not take part in the step — the reference's ``add_zero_gead_on_unused_lora`` trick (``mtmadise.py:149-157``).

The backward pass itself lives in ``madm_b200/train.py`` (``madm_backward`` behind an ``autograd.Function``); these consume ``p.grad``.
torch tensors are device memory only; there is no CPU path for the kernels (the all-reduce helper is plain torch.distributed).
"""
import ctypes as C
from typing import Iterable, List, Optional, Sequence

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _table(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _numel(tensors: Sequence[torch.Tensor]):
    arr = (C.c_int64 * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.numel()
    return arr


def _check_tensors(tensors: Sequence[torch.Tensor], what: str):
    for t in tensors:
        if t.device.type != "cuda":
            raise _lib.MadmError(f"{what}: tensors must be CUDA tensors (madm_b200 has no CPU path)")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.MadmError(f"{what}: tensors must be contiguous fp32")


def _bump_versions(tensors: Sequence[torch.Tensor]):
    """The kernels write behind autograd's back: bump the version counters (what the engine watches to repack; no launch)."""
    torch.autograd.graph.increment_version(list(tensors))


@torch.no_grad()
def update_ema(ema_params: Iterable[torch.Tensor], params: Iterable[torch.Tensor], iter: int, ema_alpha: float = 0.999) -> float:
    """``CMDISE._update_ema``: ``alpha = min(1 - 1/(iter+1), ema_alpha)``; ``ema = alpha*ema + (1-alpha)*param`` for every pair, in one
    launch per 48 tensors.  Returns alpha.  In-place updates bump the tensors' version counters, so an engine that holds them repacks."""
    ema_params, params = list(ema_params), list(params)
    if len(ema_params) != len(params):
        raise _lib.MadmError("update_ema: parameter lists differ in length")
    alpha = min(1 - 1 / (iter + 1), ema_alpha)
    keep = [(e, p) for e, p in zip(ema_params, params) if e.numel()]
    if not keep:
        return alpha
    es, ps = [e for e, _ in keep], [p for _, p in keep]  # (bumping a Parameter's version needs the Parameter, not its .data alias)
    for e, p in keep:
        if e.shape != p.shape:
            raise _lib.MadmError("update_ema: shape mismatch")
    _check_tensors(es + ps, "update_ema")
    lib = _lib.load()
    _lib.check(lib.madm_op_ema_update(_table(es), _table(ps), _numel(es), len(es), float(alpha), float(1 - alpha), _stream()), None,
               "madm_op_ema_update")
    _bump_versions(es)
    return alpha


class FusedAdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` (amsgrad=False, maximize=False) with ``clip_grad_norm_`` folded into the step.

    A ``torch.optim.Optimizer`` subclass: parameters or param-group dicts, ``add_param_group``, ``state_dict`` / ``load_state_dict``
    (state keys ``step`` / ``exp_avg`` / ``exp_avg_sq`` like torch's AdamW, so the reference's checkpointer saves and resumes it) and
    LR schedulers (``LambdaLR`` of ``config_files/common/optim.py``) work unchanged.  ``step(clip_grad=None)`` does
    norm -> clip -> AdamW with two multi-tensor launches per group and no host synchronisation, and returns the total gradient norm
    as a device scalar (what the reference logs as ``grad_norm``).  A non-finite gradient norm skips the update on the device
    (parameters and moments untouched), as ``GradScaler.step`` does under the reference's AMP trainer; the host-side ``step``
    counters still advance, since nothing is read back."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._norm = None
        self._scratch = None

    @torch.no_grad()
    def step(self, closure=None, clip_grad: Optional[float] = None) -> Optional[torch.Tensor]:
        if closure is not None:
            with torch.enable_grad():
                closure()
        lib = _lib.load()
        active = [[p for p in g["params"] if p.grad is not None] for g in self.param_groups]
        allp = [p for ps in active for p in ps]
        if not allp:
            return None
        grads = [p.grad for p in allp]
        _check_tensors([p.data for p in allp] + grads, "FusedAdamW.step")
        dev = allp[0].device
        if self._norm is None or self._norm.device != dev:
            self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        need = lib.madm_op_grad_norm_scratch_floats(len(allp))
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.empty(need, dtype=torch.float32, device=dev)
        _lib.check(lib.madm_op_grad_norm(_table(grads), _numel(grads), len(grads), C.c_void_p(self._scratch.data_ptr()),
                                         C.c_void_p(self._norm.data_ptr()), _stream()), None, "madm_op_grad_norm")
        for g, ps in zip(self.param_groups, active):
            if not ps:
                continue
            for p in ps:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1  # (a state_dict written by torch.optim.AdamW carries a tensor here)
            steps = {self.state[p]["step"] for p in ps}
            for s in sorted(steps):  # parameters that joined later have their own step count (bias correction)
                sel = [p for p in ps if self.state[p]["step"] == s]
                _check_tensors([self.state[p]["exp_avg"] for p in sel] + [self.state[p]["exp_avg_sq"] for p in sel], "FusedAdamW.step (state)")
                _lib.check(lib.madm_op_adamw_step(
                    _table([p.data for p in sel]), _table([p.grad for p in sel]), _table([self.state[p]["exp_avg"] for p in sel]),
                    _table([self.state[p]["exp_avg_sq"] for p in sel]), _numel(sel), len(sel), float(g["lr"]), float(g["betas"][0]),
                    float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), int(s),
                    C.c_void_p(self._norm.data_ptr()), float(clip_grad) if clip_grad is not None else 0.0,
                    _stream()), None, "madm_op_adamw_step")
            _bump_versions(ps)
        return self._norm


@torch.no_grad()
def allreduce_grads(params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> torch.Tensor:
    """The one collective of the training configuration: every rank contributes the gradients of ALL trainable
    parameters — zeros where a parameter took no part in this rank's step, e.g. the LoRA adapters that were not active
    (``mtmadise.py:149-157``) — through ONE all-reduce (SUM, then / world) over a flat fp32 buffer.  Afterwards every
    ``p.grad`` is a view into that buffer (zeros for the parameters that had no gradient); the flat buffer is returned.  NCCL on the GPU box, gloo in the tests."""
    import torch.distributed as dist
    params = [p for p in params if p.requires_grad]
    if not params:
        raise ValueError("allreduce_grads: no trainable parameters")
    dev, sizes = params[0].device, [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    views = [v.view_as(p) for v, p in zip(flat.split(sizes), params)]
    have = [(v, p.grad) for v, p in zip(views, params) if p.grad is not None]
    if have:  # one multi-tensor copy instead of one small kernel per parameter (567 tensors: 4 ms of launches per step)
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
    world = 1
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average and world > 1:
        flat.div_(world)
    for p, v in zip(params, views):  # the averaged gradients ARE the flat buffer: p.grad becomes a view of it (no copy back)
        p.grad = v
    return flat

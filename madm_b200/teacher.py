"""Teacher post-processing and DACS mixing of MADM's self-training step on the device, behind the C ABI
(``madm_op_pseudo_labels`` / ``madm_op_class_mask`` / ``madm_op_one_mix``, ``csrc/teacher.cu``).

Mirrors ``modeling/meta_arch/mtmadise.py:339-352`` and ``utils/dacs_transforms.py:98-112`` with the same names and argument meaning;
unlike the reference nothing leaves the GPU (its ``pseudo_label.cpu()`` / ``.item()`` pair costs a sync per step): the confidence
ratio stays a device scalar and is folded into ``pseudo_weight`` by a second kernel.  torch tensors are device memory only.
"""
import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor, what: str):
    if t.device.type != "cuda":
        raise _lib.MadmError(f"{what} must be a CUDA tensor (madm_b200 has no CPU path)")


def pseudo_labels(ema_logits: torch.Tensor, size: Sequence[int], pseudo_threshold: float, psweight_ignore_top: int = 0
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """ema_logits [B,C,h,w] -> (pseudo_label int64 [B,H,W], pseudo_prob fp32 [B,H,W], pseudo_weight fp32 [B,H,W], confident-pixel
    count as a device int32 scalar).  pseudo_weight = count / (B*H*W), zero in the top ``psweight_ignore_top`` rows (``pl_crop``)."""
    _need_cuda(ema_logits, "ema_logits")
    lib = _lib.load()
    x = ema_logits.detach().to(torch.float32).contiguous()
    B, Cc, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    dev = x.device
    label = torch.empty(B, H, W, dtype=torch.int64, device=dev)
    prob = torch.empty(B, H, W, dtype=torch.float32, device=dev)
    weight = torch.empty(B, H, W, dtype=torch.float32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.check(lib.madm_op_pseudo_labels(_ptr(x), B, Cc, h, w, H, W, float(pseudo_threshold), int(psweight_ignore_top), _ptr(label), _ptr(prob),
                                         _ptr(weight), _ptr(count), _stream()), None, "madm_op_pseudo_labels")
    return label, prob, weight, count


def generate_class_mask(label: torch.Tensor, classes: torch.Tensor) -> torch.Tensor:
    """label [H,W] (or any shape) int64, classes [k] int64 -> mask [1, *label.shape] int64 (dacs_transforms.generate_class_mask)."""
    _need_cuda(label, "label")
    lib = _lib.load()
    lab = label.to(torch.int64).contiguous()
    cls = classes.to(device=lab.device, dtype=torch.int64).contiguous()
    mask = torch.empty_like(lab)
    _lib.check(lib.madm_op_class_mask(_ptr(lab), lab.numel(), _ptr(cls), cls.numel(), _ptr(mask), _stream()), None, "madm_op_class_mask")
    return mask.unsqueeze(0)


def one_mix(mask: torch.Tensor, target: Optional[torch.Tensor] = None, weight: Optional[torch.Tensor] = None):
    """DACS mixing ``mask * t[0] + (1 - mask) * t[1]`` (dacs_transforms.one_mix) of stacked int64 labels ``target`` [2,H,W] and / or
    stacked fp32 pixel weights ``weight`` [2,H,W] in one pass; returns ([1,H,W] mixed label or None, [1,H,W] mixed weight or None)."""
    _need_cuda(mask, "mask")
    lib = _lib.load()
    m = mask[0].to(torch.int64).contiguous()
    n = m.numel()
    la = lb = lo = wa = wb = wo = None
    if target is not None:
        t = target.to(torch.int64).contiguous()
        if t.shape[0] != 2 or t[0].numel() != n:
            raise _lib.MadmError("one_mix: target must be [2, *mask.shape[1:]]")
        la, lb, lo = t[0], t[1], torch.empty_like(t[0])
    if weight is not None:
        wt = weight.to(torch.float32).contiguous()
        if wt.shape[0] != 2 or wt[0].numel() != n:
            raise _lib.MadmError("one_mix: weight must be [2, *mask.shape[1:]]")
        wa, wb, wo = wt[0], wt[1], torch.empty_like(wt[0])
    if lo is None and wo is None:
        raise _lib.MadmError("one_mix: nothing to mix")
    _lib.check(lib.madm_op_one_mix(_ptr(m), n, _ptr(la), _ptr(lb), _ptr(lo), _ptr(wa), _ptr(wb), _ptr(wo), _stream()), None, "madm_op_one_mix")
    return (lo.unsqueeze(0) if lo is not None else None), (wo.unsqueeze(0) if wo is not None else None)

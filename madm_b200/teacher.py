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


def image_mix(mask: torch.Tensor, data: torch.Tensor) -> torch.Tensor:
    """``one_mix`` on images (dacs_transforms.py:101-104): ``mask[0] * data[0] + (1 - mask[0]) * data[1]`` with the [1,H,W] class mask
    broadcast over channels; ``data`` is the stacked pair [2,C,H,W] fp32.  Returns [1,C,H,W]."""
    _need_cuda(data, "data")
    lib = _lib.load()
    d = data.to(torch.float32).contiguous()
    m = mask[0].to(device=d.device, dtype=torch.int64).contiguous()
    if d.dim() != 4 or d.shape[0] != 2 or m.numel() != d.shape[2] * d.shape[3]:
        raise _lib.MadmError("image_mix: data must be [2,C,H,W] and mask [1,H,W]")
    out = torch.empty_like(d[0])
    _lib.check(lib.madm_op_image_mix(_ptr(m), _ptr(d[0]), _ptr(d[1]), d.shape[1], m.numel(), _ptr(out), _stream()), None, "madm_op_image_mix")
    return out.unsqueeze(0)


def blur_kernel_size(h: int, w: int) -> Tuple[int, int]:
    """Kernel size rule of dacs_transforms.gaussian_blur (:66-75): ~10 % of the image side, forced odd."""
    import math

    def k(n):
        c = math.ceil(0.1 * n)
        return int(math.floor(c - 0.5 + c % 2))
    return k(h), k(w)


def gaussian_blur(blur: float, data: torch.Tensor, sigma: Optional[float] = None) -> torch.Tensor:
    """dacs_transforms.gaussian_blur (:62-84): if ``blur > 0.5`` blur the 3-channel images [B,3,H,W] with
    ``kornia.filters.GaussianBlur2d(kernel_size, (sigma, sigma))`` (separable, border 'reflect'); ``sigma`` defaults to the reference's
    ``np.random.uniform(0.15, 1.15)`` draw.  Other inputs pass through unchanged, as in the reference."""
    if data is None or data.shape[1] != 3 or not blur > 0.5:
        return data
    _need_cuda(data, "data")
    import numpy as np
    if sigma is None:
        sigma = float(np.random.uniform(0.15, 1.15))
    lib = _lib.load()
    d = data.to(torch.float32).contiguous()
    B, Cc, H, W = d.shape
    ky, kx = blur_kernel_size(H, W)
    tmp, out = torch.empty_like(d), torch.empty_like(d)
    _lib.check(lib.madm_op_gaussian_blur(_ptr(d), B * Cc, H, W, ky, kx, float(sigma), float(sigma), _ptr(tmp), _ptr(out), _stream()), None,
               "madm_op_gaussian_blur")
    return out


def color_jitter_params(batch: int, s=0.25, generator: Optional[torch.Generator] = None):
    """What kornia's ColorJitterGenerator samples for ``ColorJitter(brightness=s, contrast=s, saturation=s, hue=s)`` (or the dict form the
    reference also accepts): per image a random order of the four adjustments, brightness / contrast / saturation factors uniform in
    ``[max(0, 1 - s), 1 + s]`` and a hue factor uniform in ``[-s, s]`` (bounded by 0.5).  Host-side: a handful of scalars per image."""
    cfg = dict(brightness=s, contrast=s, saturation=s, hue=s) if not isinstance(s, dict) else dict(brightness=0.0, contrast=0.0, saturation=0.0, hue=0.0, **s)

    def uni(lo, hi):
        return lo + (hi - lo) * torch.rand(batch, generator=generator)
    b, c, sa, h = (float(cfg[k]) for k in ("brightness", "contrast", "saturation", "hue"))
    return dict(order=torch.stack([torch.randperm(4, generator=generator) for _ in range(batch)]),
                brightness_factor=uni(max(0.0, 1 - b), 1 + b), contrast_factor=uni(max(0.0, 1 - c), 1 + c),
                saturation_factor=uni(max(0.0, 1 - sa), 1 + sa), hue_factor=uni(-min(h, 0.5), min(h, 0.5)))


def color_jitter(color_jitter: float, mean=None, std=None, data: Optional[torch.Tensor] = None, target=None, s=0.25, p=0.2, params=None):
    """dacs_transforms.color_jitter (:41-59): if ``color_jitter > p`` apply kornia's ColorJitter to the 3-channel images ``data`` [B,3,H,W]
    (between ``denorm_`` / ``renorm_`` when ``mean`` / ``std`` are given), in one kernel.  ``params`` (``color_jitter_params``) can be passed to
    make the draw reproducible.  Returns ``(data, target)`` like the reference; other inputs pass through unchanged."""
    if data is None or data.shape[1] != 3 or not color_jitter > p:
        return data, target
    _need_cuda(data, "data")
    import math
    lib = _lib.load()
    d = data.to(torch.float32).contiguous()
    B, _, H, W = d.shape
    if params is None:
        params = color_jitter_params(B, s)
    dev = d.device
    order = params["order"].to(device=dev, dtype=torch.int32).contiguous()
    fac = torch.stack([params["brightness_factor"].float() - 1.0, params["contrast_factor"].float(), params["saturation_factor"].float(),
                       params["hue_factor"].float() * (2 * math.pi)], dim=1).to(dev).contiguous()
    m = sd = None
    if mean is not None and std is not None:
        m = torch.as_tensor(mean, dtype=torch.float32, device=dev).reshape(-1).expand(3).contiguous() if torch.as_tensor(mean).numel() in (1, 3) else None
        sd = torch.as_tensor(std, dtype=torch.float32, device=dev).reshape(-1).expand(3).contiguous() if torch.as_tensor(std).numel() in (1, 3) else None
        if m is None or sd is None:
            raise _lib.MadmError("color_jitter: mean / std must be scalars or per-channel 3-vectors")
    out = torch.empty_like(d)
    _lib.check(lib.madm_op_color_jitter(_ptr(d), B, H * W, _ptr(order), _ptr(fac), _ptr(m), _ptr(sd), _ptr(out), _stream()), None,
               "madm_op_color_jitter")
    return out, target

"""fp32 restatement of the teacher post-processing and DACS mixing of MADM's self-training step.
TEST INFRASTRUCTURE; PARITY UNPINNED (the reference ships no fixtures for this step).

Follows ``modeling/meta_arch/mtmadise.py:339-352`` (bilinear upsampling of the EMA head's logits, softmax, max, confidence ratio,
``pl_crop``) and ``utils/dacs_transforms.py:98-112`` (``generate_class_mask``, ``one_mix``).
"""
import numpy as np
import torch
import torch.nn.functional as F


def pseudo_labels(ema_logits: torch.Tensor, size, pseudo_threshold: float, psweight_ignore_top: int = 0):
    _ema_logits = F.interpolate(ema_logits, size=size, mode="bilinear", align_corners=False)   # mtmadise.py:339
    ema_softmax = torch.softmax(_ema_logits.detach(), dim=1)                                   # :340
    pseudo_prob, pseudo_label = torch.max(ema_softmax, dim=1)                                  # :342
    ps_large_p = pseudo_prob.ge(pseudo_threshold).long() == 1                                  # :346
    ps_size = np.size(pseudo_label.cpu().numpy())                                              # :347 (np.size(np.array(label.cpu())))
    pseudo_val = torch.sum(ps_large_p).item() / ps_size                                        # :348
    pseudo_weight = pseudo_val * torch.ones(pseudo_prob.shape, device=ema_softmax.device)      # :349
    if psweight_ignore_top > 0:                                                                # :351-352 (pl_crop)
        pseudo_weight[:, :psweight_ignore_top, :] = 0
    return pseudo_label, pseudo_prob, pseudo_weight, pseudo_val


def generate_class_mask(label: torch.Tensor, classes: torch.Tensor) -> torch.Tensor:           # dacs_transforms.py:98-103
    label, classes = torch.broadcast_tensors(label, classes.unsqueeze(1).unsqueeze(2))
    return label.eq(classes).sum(0, keepdims=True)


def one_mix(mask: torch.Tensor, target: torch.Tensor) -> torch.Tensor:                         # dacs_transforms.py:106-112
    stacked, _ = torch.broadcast_tensors(mask[0], target[0])
    return (stacked * target[0] + (1 - stacked) * target[1]).unsqueeze(0)


# ---- kornia's classic ColorJitter (0.6.x - 0.7.0), restated from its published source: kornia.color.rgb_to_hsv / hsv_to_rgb,
# kornia.enhance.adjust_{brightness,contrast,saturation,hue} and ColorJitter.apply_transform.  kornia is not installed here and the reference
# does not pin a release (requirements.txt has none), so this is what dacs_transforms.color_jitter (:41-59) is assumed to compute.
def rgb_to_hsv(image: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    import math
    max_rgb, argmax_rgb = image.max(-3)
    min_rgb = image.min(-3)[0]
    v = max_rgb
    deltac = max_rgb - min_rgb
    s = deltac / (max_rgb + eps)
    deltac = torch.where(deltac == 0, torch.ones_like(deltac), deltac)
    rc, gc, bc = torch.unbind(max_rgb.unsqueeze(-3) - image, dim=-3)
    h = torch.stack([bc - gc, (rc - bc) + 2.0 * deltac, (gc - rc) + 4.0 * deltac], dim=-3) / deltac.unsqueeze(-3)
    h = torch.gather(h, dim=-3, index=argmax_rgb.unsqueeze(-3)).squeeze(-3)
    h = (h / 6.0) % 1.0
    return torch.stack([2.0 * math.pi * h, s, v], dim=-3)


def hsv_to_rgb(image: torch.Tensor) -> torch.Tensor:
    import math
    h = image[..., 0, :, :] / (2 * math.pi)
    s, v = image[..., 1, :, :], image[..., 2, :, :]
    hi = torch.floor(h * 6) % 6
    f = ((h * 6) % 6) - hi
    p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
    hi = hi.long()
    indices = torch.stack([hi, hi + 6, hi + 12], dim=-3)
    out = torch.stack((v, q, p, p, t, v, t, v, v, q, p, p, p, p, t, v, v, q), dim=-3)
    return torch.gather(out, -3, indices)


def color_jitter_apply(data: torch.Tensor, order, brightness_factor, contrast_factor, saturation_factor, hue_factor) -> torch.Tensor:
    """ColorJitter.apply_transform on [B,3,H,W] in [0,1]; `order` [B][4] and the four factors [B] as kornia's generator samples them."""
    import math
    out = []
    for b in range(data.shape[0]):
        x = data[b:b + 1]
        for idx in order[b]:
            if idx == 0:
                x = (x + (float(brightness_factor[b]) - 1)).clamp(0, 1)
            elif idx == 1:
                x = (x * float(contrast_factor[b])).clamp(0, 1)
            elif idx == 2:
                hsv = rgb_to_hsv(x)
                x = hsv_to_rgb(torch.stack([hsv[:, 0], (hsv[:, 1] * float(saturation_factor[b])).clamp(0, 1), hsv[:, 2]], dim=1))
            else:
                hsv = rgb_to_hsv(x)
                x = hsv_to_rgb(torch.stack([torch.fmod(hsv[:, 0] + float(hue_factor[b]) * 2 * math.pi, 2 * math.pi), hsv[:, 1], hsv[:, 2]], dim=1))
        out.append(x)
    return torch.cat(out, dim=0)

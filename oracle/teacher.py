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

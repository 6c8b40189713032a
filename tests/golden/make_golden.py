#!/usr/bin/env python
"""Generates tests/golden/config1_b1.npz from the fp32 oracle on CPU (BASELINE config 1: 1x3x512x512, seed 0,
input_modal='others', adapter Depth_r16_a16, synthetic weights of oracle/synthetic.py).

The reference (XiaRho/MADM) cannot be imported here (diffusers / peft / detectron2 absent), so these vectors are outputs of
the oracle restatement, not of the reference: PARITY UNPINNED (see oracle/__init__.py).  They pin the oracle against
silent drift and let the GPU tests check the product without re-running the oracle.

``s0_b1.npz`` is the same for the vae_decoder_loss / 's0' variant all shipped experiment files select
(config_files/SemSeg/MTMADISE/mtmadise_cityscapes_rgb_to_depth_11.py:47-55): UNet final output, decoded image,
s0 / s3 / s4 / s5.

``config5_lora_grads_b1.npz`` (``grads``) holds the oracle's gradients of the LoRA training step's trainable set: the parity target of
the backward pass that SURVEY §8 row f-3 still asks for.

Run:  python tests/golden/make_golden.py [base|s0|grads]    (about 30 s each on 8 cores; deterministic for a given torch build)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synthetic  # noqa: E402
from oracle.lora import set_adapter  # noqa: E402

SUB = {"s2": 4, "s3": 2, "s4": 1, "s5": 1, "enc_tap": 4, "unet_tap16": 1, "unet_tap32": 1, "unet_tap64": 2, "s0": 8, "decoded_raw": 2}


def subsample(name, t):
    s = SUB[name]
    return t[:, :, ::s, ::s].contiguous()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    bb = synthetic.build_backbone()
    set_adapter(bb.feature_extractor.ldm_extractor.unet, ["Depth"])
    img = synthetic.synthetic_images(1)
    with torch.no_grad():
        taps = bb.feature_extractor(dict(img=img), "others")
        feats = bb.forward_features(taps)["output_features"]
    inter = bb.feature_extractor.ldm_extractor.last_intermediates
    out = {"latents": inter["latents"].numpy(), "noisy_latents": inter["noisy_latents"].numpy()}
    for name, t in zip(["enc_tap", "unet_tap16", "unet_tap32", "unet_tap64"], taps):
        out[name] = subsample(name, t).numpy().astype(np.float16)
        out[name + "_absmax"] = np.float32(t.abs().max().item())
    for name, t in feats.items():
        out[name] = subsample(name, t).numpy().astype(np.float16)
        out[name + "_absmax"] = np.float32(t.abs().max().item())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1_b1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def main_s0():
    torch.set_num_threads(os.cpu_count() or 1)
    bb = synthetic.build_backbone(variant="s0")
    set_adapter(bb.feature_extractor.ldm_extractor.unet, ["Depth"])
    img = synthetic.synthetic_images(1)
    with torch.no_grad():
        taps, fin = bb.feature_extractor(dict(img=img), "others", return_unet_final_output=True)
        feats = bb.forward_features(taps)["output_features"]
    out = {"unet_sample": fin["before_vae.decoder"].numpy(), "decoded_raw": subsample("decoded_raw", taps[0]).numpy().astype(np.float16),
           "decoded_raw_absmax": np.float32(taps[0].abs().max().item())}
    for name, t in feats.items():
        out[name] = subsample(name, t).numpy().astype(np.float16)
        out[name + "_absmax"] = np.float32(t.abs().max().item())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "s0_b1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def main_grads():
    """config5_lora_grads_b1.npz: oracle gradients of the LoRA training step's trainable set (oracle.synthetic.training_gradients) at
    1x3x512x512 — per tensor the L2 norm and the first 8 values.  The parity target of the backward pass, not yet consumed
    by a product test."""
    torch.set_num_threads(os.cpu_count() or 1)
    bb = synthetic.build_backbone()
    loss, grads = synthetic.training_gradients(bb, synthetic.synthetic_images(1))
    names = sorted(grads)
    out = {"loss": np.float32(loss), "names": np.array(names),
           "norm": np.array([0.0 if grads[n] is None else float(grads[n].double().norm()) for n in names], dtype=np.float32),
           "reached": np.array([grads[n] is not None for n in names]),
           "head8": np.stack([np.zeros(8, np.float32) if grads[n] is None else
                              np.pad(grads[n].flatten()[:8].numpy(), (0, max(0, 8 - grads[n].numel()))) for n in names]).astype(np.float32)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config5_lora_grads_b1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(names), "tensors, total norm", float(np.sqrt((out["norm"].astype(np.float64) ** 2).sum())), "loss", loss)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    if which in ("base", "both"):
        main()
    if which in ("s0", "both"):
        main_s0()
    if which == "grads":
        main_grads()

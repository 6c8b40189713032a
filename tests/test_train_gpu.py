"""SURVEY §8 row f-3 / BASELINE config 5 on the GPU: the backward pass of the drop-in backbone.

`loss.backward()` through the product backbone (autograd.Function around madm_extract(MADM_FLAG_TRAIN) / madm_backward, C ABI) against
torch.autograd through the fp32 oracle on the same device, same seeded inputs, weights and loss (oracle.synthetic.training_gradients:
a fixed linear functional of the feature dict), for the LoRA training step's trainable set: the active adapter's 256 LoRA factors, the
four GN-bottleneck projections (conv weights + GroupNorm affines) and the learned prompt / time parameters.

Gates: see GATES below (cosine per parameter family / overall / per tensor, gradient-norm ratio within 3 %).  The floor of this
comparison is not the backward's arithmetic: the path contains 12 ReLUs in the GN bottlenecks, and wherever the product's forward (16-bit
operands) and the oracle's (fp32) land on different sides of zero the two gradients differ by a full-size term.  With fp16 operands
~0.25 % of the projections' activations flip (measured and printed by test_backward_matches_oracle_autograd), i.e. sqrt(0.0025) = 5 %
relative gradient difference = cosine 0.9988 — what every family shows, including the projections' own last-layer weights; an fp32
pipeline whose weights are merely rounded to fp16 shows the same (test_gradient_noise_floor).  bf16 operands: 7x the forward noise.
against the committed oracle fixture (tests/golden/config5_lora_grads_b1.npz) per-tensor norms within 5 %.
"""
import os

import numpy as np
import pytest
import torch

from helpers import build_product_backbone, cosine, set_lora_adapter

pytestmark = pytest.mark.gpu


def make_trainable(bb, adapters=("default", "Depth")):
    """The LoRA training step's trainable set: LoRA factors, feature projections, prompt / time conditioning; base weights frozen."""
    for n, p in bb.named_parameters():
        p.requires_grad_(("lora_" in n) or n.startswith("feature_projections.") or "clip_project_" in n and not n.startswith("feature_extractor.ema_"))
        p.grad = None


def product_loss(out, seed=99):
    g = torch.Generator().manual_seed(seed)  # the oracle's functional: R_k drawn on the CPU in dict order
    return sum((v * torch.randn(v.shape, generator=g).to(v.device)).sum() for v in out.values()) / 1e3


def family(n):
    if "lora_A" in n:
        return "lora_A"
    if "lora_B" in n:
        return "lora_B"
    if n.startswith("feature_projections."):
        return "proj.norm" if ".norm." in n else "proj.conv"
    return "conditioning"


# gates per operand dtype: (per-tensor cosine, per-family / overall cosine); see the module docstring for what sets the floor
# (per-tensor, per-family) cosine gates.  The comparison has a noise band of its own: two runs of the PRODUCT that differ only in an fp32
# summation order (MADM_NO_SPLITK=1, or the TMA-store epilogue's column statistics against the coalesced epilogue's) differ from each
# other by 9e-4 of the feature norm and by cosine 0.9980 - 0.9989 per gradient family (tools/check_grads.py, profiles/r02_gradient_noise_band.txt):
# the random-init UNet amplifies rounding-level perturbations (near one-hot attention rows, ReLU / rounding boundaries).  Against the
# oracle the fp16 families therefore land anywhere in 0.9977 - 0.9990 depending on the rounding realisation; the gate sits below that band.
GATES = {"fp16": (0.99, 0.997), "bf16": (0.93, 0.985)}


def compare(grads_ref, params, tag, mode, gates=None):
    cos_tensor, cos_family = gates or GATES[mode]
    fam, per = {}, []
    for n, g in sorted(grads_ref.items()):
        if g is None:
            assert params[n].grad is None or float(params[n].grad.abs().max()) == 0.0, f"{n}: gradient where the oracle has none"
            continue
        got = params[n].grad
        assert got is not None, f"{n}: no gradient"
        assert torch.isfinite(got).all(), f"{n}: non-finite gradient"
        per.append((cosine(got, g), n, float(g.norm()), float(got.norm())))
        a, b = fam.setdefault(family(n), ([], []))
        a.append(got.flatten().double()); b.append(g.flatten().double())
    per.sort()
    for c, n, rn, gn in per[:6]:
        print(f"[{tag}] worst tensors: cosine {c:.5f}  |ref| {rn:.3e} |got| {gn:.3e}  {n}")
    allg, allr, fails = [], [], []
    for k, (a, b) in sorted(fam.items()):
        ga, gb = torch.cat(a), torch.cat(b)
        c = torch.nn.functional.cosine_similarity(ga, gb, dim=0).item()
        ratio = (ga.norm() / gb.norm()).item()
        print(f"[{tag}] {k:13s}: {len(a):4d} tensors  cosine {c:.6f}  |got|/|ref| {ratio:.4f}")
        if c < cos_family or abs(ratio - 1) > 3e-2:
            fails.append((k, c, ratio))
        allg.append(ga); allr.append(gb)
    c = torch.nn.functional.cosine_similarity(torch.cat(allg), torch.cat(allr), dim=0).item()
    print(f"[{tag}] all          : {len(per)} tensors  cosine {c:.6f}")
    assert not fails, fails
    assert c >= cos_family
    assert per[0][0] >= cos_tensor, per[0]


@pytest.fixture(scope="module", params=["bf16", "fp16"])
def pair(request, cuda_device):
    from oracle import synthetic
    ob = synthetic.build_backbone().to(cuda_device)
    pb = build_product_backbone(cuda_device, compute_dtype=request.param)
    pb.load_state_dict(ob.state_dict(), strict=True)
    make_trainable(pb)
    yield ob, pb, request.param
    del pb, ob
    torch.cuda.empty_cache()


def test_backward_matches_oracle_autograd(pair, cuda_device):
    from oracle import synthetic
    ob, pb, mode = pair
    img = synthetic.synthetic_images(2, seed=7).to(cuda_device)
    loss_ref, grads_ref = synthetic.training_gradients(ob, img, adapter="Depth", input_modal="others")
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    out = pb(img, input_modal="others")["output_features"]
    assert all(v.requires_grad for v in out.values())
    with torch.no_grad():  # how many of the final ReLUs land on different sides of zero in the two forwards
        ref_out = ob(img, input_modal="others")["output_features"]
        flips = sum(((out[k] > 0) != (ref_out[k] > 0)).sum().item() for k in out) / sum(v.numel() for v in out.values())
    print(f"[others/Depth/{mode}] final-ReLU sign flips between product and oracle forward: {100 * flips:.3f} % of the feature values "
          f"-> expected relative gradient difference ~ sqrt = {100 * flips ** 0.5:.1f} %")
    loss = product_loss(out)
    # (the functional is a random-sign sum over 22 M feature values: its value carries the forward's operand-rounding noise)
    assert abs(float(loss.detach()) - loss_ref) <= (5e-2 if mode == "bf16" else 1e-2) * max(1.0, abs(loss_ref))
    loss.backward()
    params = dict(pb.named_parameters())
    compare(grads_ref, params, f"others/Depth/{mode}", mode)
    # the adapter that did not take part got no gradient; frozen base weights neither
    assert all(p.grad is None for n, p in params.items() if ".default." in n)
    assert all(p.grad is None for n, p in params.items() if not p.requires_grad)
    # deterministic: the same step again gives bit-identical gradients
    first = {n: p.grad.clone() for n, p in params.items() if p.grad is not None}
    for p in params.values():
        p.grad = None
    product_loss(pb(img, input_modal="others")["output_features"]).backward()
    for n, g in first.items():
        assert torch.equal(g, params[n].grad), n


def test_two_passes_in_flight_source_and_mixed(pair, cuda_device):
    """MTMADISE.forward runs the source pass ('rgb', adapter 'default') and the mixed pass ('mixed', target adapter) before ONE backward of
    the summed losses (mtmadise.py:240-302, train_loop.py:277-302): two training forwards in flight, different adapters, gradients of
    the shared parameters (projections) accumulated by autograd."""
    from oracle import synthetic
    from oracle.lora import set_adapter
    ob, pb, mode = pair
    src = synthetic.synthetic_images(1, seed=11).to(cuda_device)
    mix = synthetic.synthetic_images(1, seed=12).to(cuda_device)
    # oracle: both passes under grad, one backward
    ounet = ob.feature_extractor.ldm_extractor.unet
    train = [(n, p) for n, p in ob.named_parameters() if "lora_" in n or n.startswith("feature_projections.") or
             n.startswith("feature_extractor.clip_project_")]
    for p in ob.parameters():
        p.requires_grad_(False); p.grad = None
    for _, p in train:
        p.requires_grad_(True)
    set_adapter(ounet, ["default"])
    l1 = product_loss(ob(src, input_modal="rgb")["output_features"], seed=5)
    set_adapter(ounet, ["Depth"])
    l2 = product_loss(ob(mix, input_modal="mixed")["output_features"], seed=6)
    (l1 + l2).backward()
    grads_ref = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in train}
    for _, p in train:
        p.requires_grad_(False); p.grad = None
    # product
    params = dict(pb.named_parameters())
    for p in params.values():
        p.grad = None
    punet = pb.feature_extractor.ldm_extractor.unet
    set_lora_adapter(punet, "default")
    p1 = product_loss(pb(src, input_modal="rgb")["output_features"], seed=5)
    set_lora_adapter(punet, "Depth")
    p2 = product_loss(pb(mix, input_modal="mixed")["output_features"], seed=6)
    (p1 + p2).backward()
    assert abs(float((p1 + p2).detach()) - float((l1 + l2).detach())) <= (5e-2 if mode == "bf16" else 1e-2) * max(1.0, abs(float((l1 + l2).detach())))
    compare(grads_ref, params, f"rgb/default + mixed/Depth/{mode}", mode)


def test_golden_gradient_fixture(pair, cuda_device):
    """Against the committed oracle fixture (CPU fp32, tests/golden/make_golden.py grads): per-tensor gradient norms and leading values at
    BASELINE config 1's input (1x3x512x512, seed 0, 'others', Depth)."""
    from oracle import synthetic
    _, pb, mode = pair
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config5_lora_grads_b1.npz"))
    params = dict(pb.named_parameters())
    for p in params.values():
        p.grad = None
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    img = synthetic.synthetic_images(1).to(cuda_device)
    loss = product_loss(pb(img, input_modal="others")["output_features"])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= (5e-2 if mode == "bf16" else 1e-2) * max(1.0, abs(float(g["loss"])))
    bad, heads_got, heads_ref = [], [], []
    for i, n in enumerate(str(x) for x in g["names"]):
        if not bool(g["reached"][i]):
            continue
        got = params[n].grad
        assert got is not None, n
        ref_norm = float(g["norm"][i])
        if abs(float(got.double().norm()) - ref_norm) > (5e-2 if mode == "fp16" else 2e-1) * ref_norm + 1e-9:
            bad.append((n, float(got.norm()), ref_norm))
        head = got.flatten()[:8].double().cpu().numpy()
        scale = ref_norm / np.sqrt(got.numel())  # leading values of every tensor, in units of its RMS
        heads_got.append(head / scale)
        heads_ref.append(g["head8"][i][:head.size].astype(np.float64) / scale)
    assert not bad, bad[:5]
    hg, hr = np.concatenate(heads_got), np.concatenate(heads_ref)
    c = float((hg * hr).sum() / np.sqrt((hg * hg).sum() * (hr * hr).sum()))
    print(f"[golden/{mode}] {len(heads_got)} tensors: leading-value cosine {c:.5f}")
    assert c >= GATES[mode][1]


def test_zero_adapter_training(cuda_device):
    """`lora_configs=[]` with `same_cond_params=True` (what the shipped experiment files set, mtmadise_cityscapes_rgb_to_depth_11.py:10,41):
    no LoRA factors at all — the trainable set the engine back-propagates to is the projections and the (shared) prompt / time
    parameters; the plain (un-wrapped) attention projections take the no-adapter branch of the dgrad weight packer."""
    from oracle import synthetic
    ob = synthetic.build_backbone(lora_configs=(), same_cond_params=True).to(cuda_device)
    pb = build_product_backbone(cuda_device, lora_configs=(), same_cond_params=True, compute_dtype="fp16")
    pb.load_state_dict(ob.state_dict(), strict=True)
    make_trainable(pb)
    img = synthetic.synthetic_images(1, seed=23).to(cuda_device)
    train = [(n, p) for n, p in ob.named_parameters() if n.startswith("feature_projections.") or n.startswith("feature_extractor.clip_project_rgb.")]
    for p in ob.parameters():
        p.requires_grad_(False); p.grad = None
    for _, p in train:
        p.requires_grad_(True)
    product_loss(ob(img, input_modal="others")["output_features"]).backward()
    grads_ref = {n: p.grad.detach().clone() for n, p in train}
    for _, p in train:
        p.requires_grad_(False); p.grad = None
    out = pb(img, input_modal="others")["output_features"]
    product_loss(out).backward()
    params = dict(pb.named_parameters())
    assert not any("lora_" in n for n in params)
    compare(grads_ref, params, "others/no-adapter/fp16", "fp16", gates=(0.99, 0.996))  # (one image: the 5 conditioning tensors carry more flip noise)
    del pb, ob
    torch.cuda.empty_cache()


def test_gradient_noise_floor(cuda_device):
    """The floor of the gradient comparison, measured without the product: two fp32 autograd runs of the oracle, one with its >= 2-D
    weights rounded to fp16 (an ideal fp16-weight pipeline; activations stay fp32, so this UNDER-estimates the product's forward noise).
    Their gradients already differ at the cosine ~0.999 level through the ReLU sign flips alone."""
    import copy
    from oracle import synthetic
    ob = synthetic.build_backbone().to(cuda_device)
    ob16 = copy.deepcopy(ob)
    with torch.no_grad():
        for p in ob16.parameters():
            if p.dim() >= 2:
                p.copy_(p.half().float())
    img = synthetic.synthetic_images(1, seed=7).to(cuda_device)
    _, g32 = synthetic.training_gradients(ob, img)
    _, g16 = synthetic.training_gradients(ob16, img)
    a = torch.cat([g32[n].flatten().double() for n in sorted(g32) if g32[n] is not None])
    b = torch.cat([g16[n].flatten().double() for n in sorted(g32) if g32[n] is not None])
    c = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    print(f"[noise floor] fp32 autograd, exact weights vs fp16-rounded weights: overall gradient cosine {c:.6f}")
    assert 0.99 < c < 0.99999
    del ob, ob16
    torch.cuda.empty_cache()


def test_unsupported_trainable_sets_raise(cuda_device):
    """Whole-UNet fine-tuning (finetune_unet='all' with the base weights left trainable) is outside the engine's backward: it raises."""
    from oracle import synthetic
    pb = build_product_backbone(cuda_device, compute_dtype="bf16")
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    img = synthetic.synthetic_images(1).to(cuda_device)
    with pytest.raises(NotImplementedError, match="base UNet weights"):
        pb(img, input_modal="others")
    make_trainable(pb)
    with pytest.raises(NotImplementedError, match="ema_forward"):
        pb(img, input_modal="others", ema_forward=True)
    del pb
    torch.cuda.empty_cache()

"""GPU parity of the vae_decoder_loss / 's0' variant: the configuration all shipped experiment files
select (config_files/SemSeg/MTMADISE/mtmadise_cityscapes_rgb_to_depth_11.py:47-55).  The UNet runs to its final output
(ldm_diffusers.py:608-611), the VAE decoder decodes it (ldm_diffusers.py:314-346), the decoded 3 x 512 x 512 image is the first
feature and Bottleneck(3 -> 128 -> 128) projects it into 's0' [B,128,512,512].

Product (CUDA engine through the C ABI) against the fp32 oracle on the same device with TF32 off, same seeded synthetic weights
and inputs (oracle.synthetic.build_backbone(variant='s0')).  Gates as for the base path (north_star): per tensor cosine >= 0.999
and max|a-b|/max|b| <= 2e-2 with the default fp16 operands.
"""
import os

import numpy as np
import pytest
import torch

from helpers import build_product_backbone, cosine, max_rel, set_lora_adapter

pytestmark = pytest.mark.gpu

COS_MIN = 0.999
REL_MAX = 2e-2


@pytest.fixture(scope="module")
def pair(cuda_device):
    from oracle import synthetic
    ob = synthetic.build_backbone(variant="s0").to(cuda_device)
    pb = build_product_backbone(cuda_device, variant="s0")
    missing, unexpected = pb.load_state_dict(ob.state_dict(), strict=False)
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    yield ob, pb
    del pb, ob
    torch.cuda.empty_cache()


def _check(name, got, ref, rel_max=REL_MAX):
    c, r = cosine(got, ref), max_rel(got, ref)
    print(f"[s0] {name}: cos={c:.6f} max_rel={r:.5f}")
    assert c >= COS_MIN, f"{name}: cosine {c}"
    assert r <= rel_max, f"{name}: max rel err {r} > {rel_max}"


def test_s0_variant_parity(pair, cuda_device):
    from oracle import synthetic
    from oracle.lora import set_adapter
    ob, pb = pair
    img = synthetic.synthetic_images(2).to(cuda_device)
    set_adapter(ob.feature_extractor.ldm_extractor.unet, ["Depth"])
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    with torch.no_grad():
        taps, fin = ob.feature_extractor(dict(img=img), "others", return_unet_final_output=True)
        feats = ob.forward_features(taps, None)["output_features"]
        res = pb._extract(img, "others", False, None, want_taps=True, return_unet_final_output=True)
    _check("before_vae.decoder", res["unet_sample"], fin["before_vae.decoder"])
    _check("decoder_output", res["taps"][0], taps[0])
    # 'after_vae.decoder' is clip(decoder_output, -1, 1): the clip itself is bit-exact (asserted), so its error is the decoder
    # output's, measured on the decoder output's scale (the clipped copy's own max is 1 by construction)
    assert res["decoded"].min() >= -1.0 and res["decoded"].max() <= 1.0
    assert torch.equal(res["decoded"], res["taps"][0].clamp(-1.0, 1.0))
    err = (res["decoded"] - fin["after_vae.decoder"]).abs().max().item() / taps[0].abs().max().item()
    print(f"[s0] after_vae.decoder: cos={cosine(res['decoded'], fin['after_vae.decoder']):.6f} max_rel={err:.5f}")
    assert cosine(res["decoded"], fin["after_vae.decoder"]) >= COS_MIN and err <= REL_MAX
    for k, got in zip(["s0", "s3", "s4", "s5"], res["features"]):
        assert got.shape == feats[k].shape
        _check(k, got, feats[k])


@torch.no_grad()
def test_s0_public_forward_and_final_output_dict(pair, cuda_device):
    """backbone(img, return_unet_final_output=True) returns (feature dict, {'before_vae.decoder', 'after_vae.decoder'}) as
    feature_extractor.py:164-166 / ldm_diffusers.py:211-215; without the kwarg just the dict; graph replay == eager launch."""
    from oracle import synthetic
    ob, pb = pair
    img = synthetic.synthetic_images(1, seed=3).to(cuda_device)
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    out, fin = pb(img, input_modal="others", return_unet_final_output=True)
    assert list(out["output_features"].keys()) == ["s0", "s3", "s4", "s5"]
    assert out["output_features"]["s0"].shape == (1, 128, 512, 512) and out["output_features"]["s5"].shape == (1, 512, 16, 16)
    assert set(fin.keys()) == {"before_vae.decoder", "after_vae.decoder"}
    assert fin["before_vae.decoder"].shape == (1, 4, 64, 64) and fin["after_vae.decoder"].shape == (1, 3, 512, 512)
    out2 = pb(img, input_modal="others")
    assert isinstance(out2, dict)
    for k in out["output_features"]:
        assert torch.equal(out["output_features"][k], out2["output_features"][k]), k  # deterministic, graph vs graph
    eng = pb.feature_extractor.ldm_extractor.engine()
    gmax, eng.graph_max_batch = eng.graph_max_batch, 0
    try:
        out3 = pb(img, input_modal="others")
    finally:
        eng.graph_max_batch = gmax
    for k in out["output_features"]:
        assert torch.equal(out["output_features"][k], out3["output_features"][k]), k  # graph replay == eager launches


def test_s0_golden_fixture(pair, cuda_device):
    """Committed oracle fixture of the variant (tests/golden/s0_b1.npz, tests/golden/make_golden.py s0; CPU fp32)."""
    from oracle import synthetic
    ob, pb = pair
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "s0_b1.npz"))
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    img = synthetic.synthetic_images(1).to(cuda_device)
    with torch.no_grad():
        res = pb._extract(img, "others", False, None, want_taps=True, return_unet_final_output=True)
    _check("golden/unet_sample", res["unet_sample"].cpu(), torch.from_numpy(g["unet_sample"]))
    sub = {"s0": 8, "s3": 2, "s4": 1, "s5": 1}
    for (k, s), got in zip(sub.items(), res["features"]):
        ref = torch.from_numpy(g[k].astype(np.float32))
        got = got[:, :, ::s, ::s].cpu()
        c = cosine(got, ref)
        r = ((got.double() - ref.double()).abs().max() / float(g[k + "_absmax"])).item()
        print(f"[s0] golden/{k}: cos={c:.6f} max_rel={r:.5f}")
        assert c >= COS_MIN and r <= REL_MAX, (k, c, r)


def test_s0_ema_and_sliding_window(pair, cuda_device):
    """EMA projections and the sliding window (512x1024 -> 3 crops) in the variant: s0 is merged at stride 1."""
    from oracle import synthetic
    ob, pb = pair
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    img = synthetic.synthetic_images(1, h=512, w=1024, seed=11).to(cuda_device)
    with torch.no_grad():
        ref = ob.slide_forward(img, "others", ema_forward=True)["output_features"]
        got = pb.slide_forward(img, "others", ema_forward=True)["output_features"]
    assert got["s0"].shape == (1, 128, 512, 1024)
    for k in ref:
        _check("slide-ema/" + k, got[k], ref[k])


def test_s0_head_image_to_segmentation(pair, cuda_device):
    """Row f-2 in the variant: DAFormerHead with in_keys[0]='s0', in_channels[0]=128 fuses on the 512^2 grid
    (mtmadise_cityscapes_rgb_to_depth_11.py:51-55).  Logits parity on the oracle's features, and image -> product backbone ->
    product head against image -> oracle backbone -> oracle head: argmax >= 99.5 % pixel-identical (north_star gate)."""
    from madm_b200.head import DAFormerHead
    from oracle import synthetic
    from oracle.daformer_head import build_head
    from test_head_gpu import HEAD_KW
    ob, pb = pair
    oh = build_head(variant="s0").to(cuda_device)
    kw = dict(HEAD_KW, in_channels=[128, 512, 512, 512], in_keys=["s0", "s3", "s4", "s5"])
    ph = DAFormerHead(**kw, device=cuda_device).eval()
    assert set(ph.state_dict().keys()) == set(oh.state_dict().keys())
    ph.load_state_dict(oh.state_dict())
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    img = synthetic.synthetic_images(1, seed=21).to(cuda_device)
    with torch.no_grad():
        rf = ob(img, input_modal="others")
        ref = oh(rf)
        same_feats = ph(rf)
        got = ph(pb(img, input_modal="others"))
    assert got.shape == ref.shape == (1, 19, 512, 512)
    _check("head(s0)/logits on oracle features", same_feats, ref)
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"[s0] image -> segmentation argmax agreement {agree * 100:.3f} %")
    assert agree >= 0.995


def test_module_level_vae_encoder(pair, cuda_device):
    """`from modeling.meta_arch.ldm_diffusers import vae_encoder` (mtmadise.py:15): the meta-arch encodes colour targets in [-1,1] with
    the backbone's VAE for the vae_decoder_loss (mtmadise.py:254,345,398,463).  Drop-in: madm_b200.ldm.vae_encoder."""
    from madm_b200.ldm import vae_encoder
    from oracle import sd14, synthetic
    ob, pb = pair
    x = synthetic.synthetic_images(2, seed=31).to(cuda_device) * 2.0 - 1.0
    with torch.no_grad():
        ref, ref_feats = sd14.vae_encoder(ob.feature_extractor.ldm_extractor.vae, x, [])
        v0 = pb.feature_extractor.ldm_extractor.engine()._versions
        lat, feats = vae_encoder(pb.feature_extractor.ldm_extractor.vae, x, encoder_block_indices=[])
    assert feats == [] and ref_feats == [] and lat.shape == (2, 4, 64, 64)
    assert pb.feature_extractor.ldm_extractor.engine()._versions is v0  # no re-bind / repack between backbone calls and vae_encoder
    _check("vae_encoder/latents", lat, ref)
    with pytest.raises(NotImplementedError):
        vae_encoder(pb.feature_extractor.ldm_extractor.vae, x, encoder_block_indices=[5])  # no encoder tap in the s0 configuration


def test_s0_zero_adapter_shipped_configuration(cuda_device):
    """The configuration the three shipped experiment files actually run: the s0 variant with `model.lora_configs = []`
    (mtmadise_cityscapes_rgb_to_depth_11.py:10, :47-55) and `same_cond_params=True` (:41) — un-wrapped `to_q.weight` keys, no adapter
    fold, one shared prompt / time parameter set — against the oracle built the same way."""
    from oracle import synthetic
    ob = synthetic.build_backbone(lora_configs=(), variant="s0", same_cond_params=True).to(cuda_device)
    pb = build_product_backbone(cuda_device, lora_configs=(), variant="s0", same_cond_params=True)
    sd = ob.state_dict()
    assert not any("base_layer" in k or "lora_" in k for k in sd)
    pb.load_state_dict(sd, strict=True)
    img = synthetic.synthetic_images(1, seed=73).to(cuda_device)
    with torch.no_grad():
        ref, rfin = ob(img, input_modal="others", return_unet_final_output=True)
        out, fin = pb(img, input_modal="others", return_unet_final_output=True)
    _check("no-lora/before_vae.decoder", fin["before_vae.decoder"], rfin["before_vae.decoder"])
    for k in ("s0", "s3", "s4", "s5"):
        _check("no-lora/" + k, out["output_features"][k], ref["output_features"][k])
    del pb, ob
    torch.cuda.empty_cache()

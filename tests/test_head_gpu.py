"""the DAFormer head (reference modeling/sem_seg_head/daformer_head.py:702-749) as a CUDA stage behind the C ABI.

Parity against the fp32 oracle restatement (oracle/daformer_head.py) on the same seeded inputs / weights:
logits cosine >= 0.999, max|a-b|/max|b| <= 2e-2 (1e-2 expected: 16-bit operands, fp32 accumulate), argmax >= 99.5 % identical.
Byte/index kernels bit-exact, interpolation / depthwise kernels within 16-bit output rounding."""
import pytest
import torch
import torch.nn.functional as F

from helpers import cosine, max_rel

pytestmark = pytest.mark.gpu

HEAD_KW = dict(in_channels=[512] * 4, in_keys=["s2", "s3", "s4", "s5"], channels=256, num_classes=19, in_index=[0, 1, 2, 3],
               norm_cfg=dict(type="BN", requires_grad=True), align_corners=False,
               decoder_params=dict(embed_dims=256, embed_cfg=dict(type="mlp", act_cfg=None, norm_cfg=None),
                                   embed_neck_cfg=dict(type="mlp", act_cfg=None, norm_cfg=None),
                                   fusion_cfg=dict(type="aspp", sep=True, dilations=(1, 6, 12, 18), pool=False, act_cfg=dict(type="ReLU"),
                                                   norm_cfg=dict(type="BN", requires_grad=True))))


@pytest.fixture(scope="module")
def ops():
    from madm_b200 import ops as o
    return o


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_nchw_to_nhwc16(ops, cuda_device, dt):
    x = torch.randn(3, 70, 17, 19, device=cuda_device)
    y = ops.nchw_to_nhwc16(x, dtype=dt)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous().to(dt))


@pytest.mark.parametrize("hs,hd", [(64, 128), (32, 128), (16, 128), (24, 40)])
def test_bilinear_resize(ops, cuda_device, hs, hd):
    x = torch.randn(2, hs, hs, 64, device=cuda_device).half()
    cat = torch.zeros(2, hd, hd, 192, device=cuda_device, dtype=torch.float16)
    ops.bilinear_resize(x, hd, hd, out=cat[..., 64:], pitch=192)  # written into the middle slice of a wider concat buffer
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=(hd, hd), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert max_rel(cat[..., 64:128].float(), ref) < 2e-3  # fp16 output rounding
    assert torch.count_nonzero(cat[..., :64]) == 0 and torch.count_nonzero(cat[..., 128:]) == 0


@pytest.mark.parametrize("dil", [1, 6, 12, 18, 45])  # 45 > H: every tap but the centre falls into the zero padding
def test_depthwise3x3(ops, cuda_device, dil):
    g = torch.Generator(device="cuda").manual_seed(dil)
    Cc = 128
    x = torch.randn(2, 40, 48, Cc, device=cuda_device, generator=g).half()
    w = torch.randn(Cc, 1, 3, 3, device=cuda_device, generator=g) * 0.3
    shift = torch.randn(Cc, device=cuda_device, generator=g) * 0.1
    w9 = w.reshape(Cc, 9).t().contiguous()
    y = ops.depthwise3x3(x, w9, shift, dil)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=dil, dilation=dil, groups=Cc) + shift[None, :, None, None]).permute(0, 2, 3, 1)
    assert max_rel(y.float(), ref) < 2e-3


@pytest.fixture(scope="module")
def heads(cuda_device):
    from madm_b200.head import DAFormerHead
    from oracle.daformer_head import build_head
    oh = build_head().to(cuda_device)
    ph = DAFormerHead(**HEAD_KW, device=cuda_device).eval()
    assert set(ph.state_dict().keys()) == set(oh.state_dict().keys())  # reference key names (mmcv ConvModule / DepthwiseSeparableConvModule)
    ph.load_state_dict(oh.state_dict())
    return oh, ph


def _feats(cuda_device, B, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    # post-ReLU, O(1) magnitude like the GN-bottleneck projections' outputs
    return {k: F.relu(torch.randn(B, 512, s, s, device=cuda_device, generator=g)) for k, s in zip(("s2", "s3", "s4", "s5"), (128, 64, 32, 16))}


@pytest.mark.parametrize("B", [1, 3])
def test_head_parity_on_random_features(heads, cuda_device, B):
    oh, ph = heads
    feats = _feats(cuda_device, B, 100 + B)
    with torch.no_grad():
        ref = oh({"output_features": feats})
        out = ph({"output_features": feats})
    assert out.shape == ref.shape == (B, 19, 128, 128)
    c, r = cosine(out, ref), max_rel(out, ref)
    agree = (out.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"head logits: cosine {c:.6f} max-rel {r:.2e} argmax agreement {100 * agree:.3f} %")
    assert c >= 0.999 and r <= 2e-2 and agree >= 0.995
    with torch.no_grad():
        assert torch.equal(out, ph({"output_features": feats}))  # no atomics anywhere: bit-identical reruns


@pytest.mark.parametrize("hw", [(256, 512), (256, 256), (64, 128)])
def test_head_on_full_image_grids(heads, cuda_device, hw):
    """Sliding-window inference merges the crops' features into full-image maps before the head runs (feature_extractor.py:270-275):
    the head stage takes the grid as a call argument (madm_extract_args.head_h / head_w), e.g. 256 x 512 for a 1024 x 2048 image
    (BASELINE config 4) and 256 x 256 for 1024 x 1024 (config 3)."""
    oh, ph = heads
    g = torch.Generator(device="cuda").manual_seed(hw[0] + hw[1])
    feats = {k: F.relu(torch.randn(1, 512, hw[0] // r, hw[1] // r, device=cuda_device, generator=g)) for k, r in zip(("s2", "s3", "s4", "s5"), (1, 2, 4, 8))}
    with torch.no_grad():
        ref = oh({"output_features": feats})
        out = ph({"output_features": feats})
    assert out.shape == ref.shape == (1, 19, hw[0], hw[1])
    c, r = cosine(out, ref), max_rel(out, ref)
    agree = (out.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"head logits on {hw}: cosine {c:.6f} max-rel {r:.2e} argmax agreement {100 * agree:.3f} %")
    assert c >= 0.999 and r <= 2e-2 and agree >= 0.995


def test_head_bf16_operands(cuda_device):
    """The head follows the context's operand dtype: bf16 operands within the looser bf16 gate (cosine >= 0.999, 3e-2)."""
    from madm_b200.head import DAFormerHead
    from oracle.daformer_head import build_head
    oh = build_head().to(cuda_device)
    ph = DAFormerHead(**HEAD_KW, device=cuda_device, compute_dtype="bf16").eval()
    ph.load_state_dict(oh.state_dict())
    feats = _feats(cuda_device, 2, 77)
    with torch.no_grad():
        ref = oh({"output_features": feats})
        out = ph({"output_features": feats})
    assert cosine(out, ref) >= 0.999 and max_rel(out, ref) <= 3e-2
    assert (out.argmax(1) == ref.argmax(1)).float().mean().item() >= 0.97


def test_head_refolds_after_parameter_update(heads, cuda_device):
    """BatchNorm statistics / weights are folded at pack time: an in-place update must trigger a repack (version tracking)."""
    oh, ph = heads
    feats = _feats(cuda_device, 1, 7)
    with torch.no_grad():
        before = ph({"output_features": feats})
        for m in (oh, ph):
            m.fuse_layer.bottleneck.bn.running_var.mul_(1.5)
            m.conv_seg.bias.add_(0.25)
        ref = oh({"output_features": feats})
        out = ph({"output_features": feats})
        for m in (oh, ph):  # restore for the other tests of this module
            m.fuse_layer.bottleneck.bn.running_var.div_(1.5)
            m.conv_seg.bias.sub_(0.25)
    assert not torch.equal(before, out)
    assert cosine(out, ref) >= 0.999 and max_rel(out, ref) <= 2e-2


def test_image_to_segmentation_end_to_end(heads, cuda_device):
    """Product backbone -> product head against oracle backbone -> oracle head: the north-star argmax gate on the whole chain
    image -> logits, after the meta-arch's bilinear upsampling to the input size (mtmadise.py:685-688)."""
    from oracle import synthetic
    from oracle.lora import set_adapter
    from helpers import build_product_backbone, set_lora_adapter
    oh, ph = heads
    ob = synthetic.build_backbone().to(cuda_device).eval()
    pb = build_product_backbone(cuda_device)
    missing, unexpected = pb.load_state_dict(ob.state_dict(), strict=False)
    assert not missing and not unexpected
    set_adapter(ob.feature_extractor.ldm_extractor.unet, ["Depth"])
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    img = synthetic.synthetic_images(2, seed=43).to(cuda_device)
    with torch.no_grad():
        ref = oh(ob(img, input_modal="others"))
        out = ph(pb(img, input_modal="others"))
        seg_ref = F.interpolate(ref, size=img.shape[-2:], mode="bilinear", align_corners=False).argmax(1)
        seg_out = F.interpolate(out, size=img.shape[-2:], mode="bilinear", align_corners=False).argmax(1)
    agree = (seg_ref == seg_out).float().mean().item()
    print(f"image -> segmentation: logits cosine {cosine(out, ref):.6f} max-rel {max_rel(out, ref):.2e} argmax agreement {100 * agree:.3f} %")
    assert seg_ref.unique().numel() >= 5
    assert cosine(out, ref) >= 0.999 and agree >= 0.995

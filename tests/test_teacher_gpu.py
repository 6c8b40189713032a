"""teacher post-processing + DACS mixing kernels against the torch restatement of the reference
(oracle/teacher.py; mtmadise.py:339-352, utils/dacs_transforms.py:98-112).  Integer outputs bit-exact (arg-max labels up to
ties of the interpolated logits at float rounding level), probabilities to 1e-5, the confidence count within those near-ties."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,C,h,w,H,W,thr,top", [(2, 19, 128, 128, 512, 512, 0.968, 0), (3, 19, 64, 96, 256, 384, 0.6, 15),
                                                 (1, 11, 37, 53, 148, 212, 0.9, 0), (2, 19, 128, 128, 128, 128, 0.5, 3)])
def test_pseudo_labels(cuda_device, B, C, h, w, H, W, thr, top):
    from madm_b200 import teacher
    from oracle import teacher as ot
    g = torch.Generator(device="cuda").manual_seed(B * C + H)
    logits = torch.randn(B, C, h, w, device=cuda_device, generator=g) * 4
    lab_r, prob_r, w_r, val_r = ot.pseudo_labels(logits, (H, W), thr, top)
    lab, prob, wgt, count = teacher.pseudo_labels(logits, (H, W), thr, top)
    assert lab.dtype == torch.int64 and lab.shape == lab_r.shape
    agree = (lab == lab_r).float().mean().item()
    assert agree >= 0.9999, agree  # near-ties of interpolated logits may resolve differently
    same = lab == lab_r
    assert torch.allclose(prob[same], prob_r[same], rtol=0, atol=1e-5)
    n = B * H * W
    near_thr = ((prob_r - thr).abs() < 1e-5).sum().item()
    assert abs(int(count.item()) - round(val_r * n)) <= near_thr + (n - int(same.sum().item()))
    ratio = count.item() / n
    body = wgt[:, top:, :]
    assert torch.equal(body, torch.full_like(body, ratio))
    if top:
        assert torch.count_nonzero(wgt[:, :top, :]) == 0
    # reruns are bit-identical (integer atomics only)
    lab2, prob2, wgt2, count2 = teacher.pseudo_labels(logits, (H, W), thr, top)
    assert torch.equal(lab, lab2) and torch.equal(prob, prob2) and torch.equal(wgt, wgt2) and torch.equal(count, count2)


def test_class_mask_and_one_mix(cuda_device):
    from madm_b200 import teacher
    from oracle import teacher as ot
    g = torch.Generator(device="cuda").manual_seed(3)
    H, W = 96, 160
    gt = torch.randint(0, 19, (H, W), device=cuda_device, generator=g)
    gt[:5] = 255  # ignore label
    pl = torch.randint(0, 19, (H, W), device=cuda_device, generator=g)
    classes = torch.tensor([0, 3, 7, 18, 255], device=cuda_device)
    mask = teacher.generate_class_mask(gt, classes)
    assert torch.equal(mask, ot.generate_class_mask(gt, classes))
    gt_w = torch.ones(H, W, device=cuda_device)
    ps_w = torch.full((H, W), 0.37, device=cuda_device)
    mixed_lbl, mixed_w = teacher.one_mix(mask, target=torch.stack((gt, pl)), weight=torch.stack((gt_w, ps_w)))
    assert torch.equal(mixed_lbl, ot.one_mix(mask, torch.stack((gt, pl))))
    assert torch.equal(mixed_w, ot.one_mix(mask, torch.stack((gt_w, ps_w))))
    only_lbl, none_w = teacher.one_mix(mask, target=torch.stack((gt, pl)))
    assert none_w is None and torch.equal(only_lbl, mixed_lbl)


def test_image_to_pseudo_labels_chain(cuda_device):
    """image -> product backbone -> product head -> pseudo labels: the teacher path of mtmadise.py:335-352 without leaving the GPU,
    against oracle backbone -> oracle head -> torch post-processing."""
    from oracle import synthetic, teacher as ot
    from oracle.daformer_head import build_head
    from oracle.lora import set_adapter
    from helpers import build_product_backbone, set_lora_adapter
    from madm_b200 import teacher
    from madm_b200.head import DAFormerHead
    from test_head_gpu import HEAD_KW
    oh = build_head().to(cuda_device)
    ph = DAFormerHead(**HEAD_KW, device=cuda_device).eval()
    ph.load_state_dict(oh.state_dict())
    ob = synthetic.build_backbone().to(cuda_device).eval()
    pb = build_product_backbone(cuda_device)
    pb.load_state_dict(ob.state_dict(), strict=False)
    set_adapter(ob.feature_extractor.ldm_extractor.unet, ["Depth"])
    set_lora_adapter(pb.feature_extractor.ldm_extractor.unet, "Depth")
    img = synthetic.synthetic_images(2, seed=47).to(cuda_device)
    with torch.no_grad():
        lab_r, prob_r, w_r, val_r = ot.pseudo_labels(oh(ob(img, input_modal="others")), img.shape[2:], 0.3, 15)
        lab, prob, wgt, count = teacher.pseudo_labels(ph(pb(img, input_modal="others")), img.shape[2:], 0.3, 15)
    agree = (lab == lab_r).float().mean().item()
    print(f"pseudo-label agreement {100 * agree:.3f} %, confident ratio {count.item() / lab.numel():.4f} vs {val_r:.4f}")
    assert agree >= 0.995
    assert abs(count.item() / lab.numel() - val_r) < 5e-3

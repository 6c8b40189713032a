"""SURVEY §8 row f-3 (optimizer side) and f-4 (image side) on the GPU, through the C ABI:

* FusedAdamW + folded clip_grad_norm_ against torch.optim.AdamW + torch.nn.utils.clip_grad_norm_ (the reference's optimizer,
  config_files/common/optim.py:9-18, engine/train_loop.py:123-124) on the same tensors — fp32 arithmetic in the same order, tolerance
  1e-6 relative (fma contraction / reduction order),
* update_ema against CMDISE._update_ema's formula (cmdise.py:337-349) — bit-exact,
* image_mix against dacs_transforms.one_mix (bit-exact), gaussian_blur against a plain-torch restatement of
  kornia.filters.GaussianBlur2d (separable Gaussian, border 'reflect'; kornia itself is not installed: parity unpinned).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(128, 3, 1, 1), (128,), (128, 128, 3, 3), (512, 128, 1, 1), (1, 77, 768), (1280,), (1, 1, 1280), (16, 320), (320, 16), (7,), (1,),
          (1280, 16), (33, 5, 3)] * 5  # 65 tensors: more than one 48-tensor launch group


def _make(dev, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return [torch.randn(s, device=dev, generator=g) for s in SHAPES]


@pytest.mark.parametrize("clip", [None, 0.01, 1e9])
def test_fused_adamw_matches_torch(cuda_device, clip):
    from madm_b200.optim import FusedAdamW
    ref_p = [torch.nn.Parameter(t.clone()) for t in _make(cuda_device, 1)]
    our_p = [torch.nn.Parameter(t.detach().clone()) for t in ref_p]
    kw = dict(lr=5e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8)
    ref = torch.optim.AdamW([{"params": ref_p[:40]}, {"params": ref_p[40:], "weight_decay": 0.0, "lr": 1e-3}], **kw)
    ours = FusedAdamW([{"params": our_p[:40]}, {"params": our_p[40:], "weight_decay": 0.0, "lr": 1e-3}], **kw)
    for it in range(5):
        grads = _make(cuda_device, 100 + it)
        for k, (a, b) in enumerate(zip(ref_p, our_p)):
            if it == 0 and k % 7 == 3:  # some parameters join at the second step: their own bias correction
                a.grad, b.grad = None, None
                continue
            a.grad, b.grad = grads[k].clone(), grads[k].clone()
        ref_norm = None
        if clip is not None:
            ref_norm = torch.nn.utils.clip_grad_norm_([p for p in ref_p if p.grad is not None], clip)
        ref.step()
        v0 = our_p[0]._version
        norm = ours.step(clip_grad=clip)
        assert our_p[0]._version > v0  # engines watching the parameter repack
        if ref_norm is not None:
            assert abs(norm.item() - ref_norm.item()) <= 1e-5 * ref_norm.item()
    for a, b in zip(ref_p, our_p):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (a.shape, (a - b).abs().max().item())
    for a, b in zip(ref_p, our_p):
        if a in ref.state:
            assert torch.allclose(ref.state[a]["exp_avg"], ours.state[b]["exp_avg"], rtol=2e-6, atol=2e-7)  # lerp cancels: absolute rounding of O(1) terms
            assert torch.allclose(ref.state[a]["exp_avg_sq"], ours.state[b]["exp_avg_sq"], rtol=2e-6, atol=1e-8)


def test_update_ema_bit_exact(cuda_device):
    from madm_b200.optim import update_ema
    params = [torch.nn.Parameter(t) for t in _make(cuda_device, 3)] + [torch.nn.Parameter(torch.tensor(0.37, device=cuda_device))]
    ema = [torch.nn.Parameter(t) for t in _make(cuda_device, 4)] + [torch.nn.Parameter(torch.tensor(-1.5, device=cuda_device))]
    ref = [e.detach().clone() for e in ema]
    for it in (0, 1, 5, 5000):
        alpha = min(1 - 1 / (it + 1), 0.999)
        for r, p in zip(ref, params):  # cmdise.py:341-348
            r.copy_(alpha * r + (1 - alpha) * p.data)
        v0 = ema[0]._version
        a = update_ema(ema, params, it, 0.999)
        assert a == alpha and ema[0]._version > v0
        for r, e in zip(ref, ema):
            assert torch.equal(r, e.data)


def test_image_mix_bit_exact(cuda_device):
    from madm_b200.teacher import image_mix
    g = torch.Generator(device="cuda").manual_seed(5)
    data = torch.rand(2, 3, 96, 160, device=cuda_device, generator=g)
    mask = (torch.rand(1, 96, 160, device=cuda_device, generator=g) > 0.4).long()
    m, _ = torch.broadcast_tensors(mask[0], data[0])  # dacs_transforms.one_mix
    ref = (m * data[0] + (1 - m) * data[1]).unsqueeze(0)
    assert torch.equal(image_mix(mask, data), ref)


def _kornia_gaussian_blur(x, ky, kx, sigma):
    def k1(n):
        t = torch.arange(n, device=x.device, dtype=x.dtype) - n // 2
        g = torch.exp(-t.pow(2) / (2 * sigma ** 2))
        return g / g.sum()
    k2 = k1(ky)[:, None] * k1(kx)[None, :]
    c = x.shape[1]
    xp = F.pad(x, (kx // 2, kx // 2, ky // 2, ky // 2), mode="reflect")
    return F.conv2d(xp, k2.expand(c, 1, ky, kx).contiguous(), groups=c)


@pytest.mark.parametrize("hw,sigma", [((512, 512), 0.15), ((512, 512), 1.15), ((512, 1024), 0.7), ((64, 96), 0.5)])
def test_gaussian_blur(cuda_device, hw, sigma):
    from madm_b200.teacher import blur_kernel_size, gaussian_blur
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand(2, 3, *hw, device=cuda_device, generator=g)
    ky, kx = blur_kernel_size(*hw)
    assert ky % 2 == 1 and kx % 2 == 1 and (hw != (512, 512) or (ky, kx) == (51, 51))
    assert gaussian_blur(0.3, x, sigma) is x  # blur <= 0.5: untouched (dacs_transforms.py:65)
    out = gaussian_blur(0.9, x, sigma)
    ref = _kornia_gaussian_blur(x, ky, kx, sigma)
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-6


def test_color_jitter(cuda_device):
    """dacs_transforms.color_jitter = kornia ColorJitter (classic arithmetic) between denorm_ / renorm_, against the torch restatement of
    kornia's published formulas (oracle/teacher.py; kornia is not installed: parity unpinned).  Piecewise functions (argmax of RGB, floor of
    the hue sextant) may flip for pixels that sit on a boundary to within fp32 rounding, hence the tiny tolerated fraction of outliers."""
    from madm_b200.teacher import color_jitter, color_jitter_params
    from oracle import teacher as ot
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(6, 3, 96, 128, device=cuda_device, generator=g)
    x[0, :, :8, :8] = 0.5          # grey patch: zero saturation, delta == 0 branch
    x[1, 0] = x[1, 1]              # r == g ties in the arg-max
    params = color_jitter_params(6, 0.25, generator=torch.Generator().manual_seed(3))
    params["order"][0] = torch.tensor([3, 2, 1, 0])
    params["hue_factor"][1] = -0.25
    assert color_jitter(0.1, data=x, p=0.2)[0] is x  # not selected: untouched
    got, tgt = color_jitter(0.9, data=x, target="lbl", s=0.25, p=0.2, params=params)
    ref = ot.color_jitter_apply(x, params["order"].tolist(), params["brightness_factor"], params["contrast_factor"],
                                params["saturation_factor"], params["hue_factor"])
    assert tgt == "lbl" and got.shape == ref.shape
    err = (got - ref).abs()
    assert (err > 1e-5).float().mean().item() < 1e-4, ((err > 1e-5).float().mean().item(), err.max().item())
    assert got.min() >= 0 and got.max() <= 1
    # with normalisation constants: denorm_ -> jitter -> renorm_
    mean, std = torch.tensor([0.2, 0.1, 0.3]), torch.tensor([0.5, 0.6, 0.4])
    xn = (x - mean.view(1, 3, 1, 1).to(cuda_device)) / std.view(1, 3, 1, 1).to(cuda_device)
    got2, _ = color_jitter(0.9, mean=mean, std=std, data=xn, params=params)
    den = xn * std.view(1, 3, 1, 1).to(cuda_device) + mean.view(1, 3, 1, 1).to(cuda_device)
    ref2 = (ot.color_jitter_apply(den.clamp(0, 1), params["order"].tolist(), params["brightness_factor"], params["contrast_factor"],
                                  params["saturation_factor"], params["hue_factor"]) - mean.view(1, 3, 1, 1).to(cuda_device)) / std.view(1, 3, 1, 1).to(cuda_device)
    err2 = (got2 - ref2).abs()
    assert (err2 > 1e-4).float().mean().item() < 1e-3, ((err2 > 1e-4).float().mean().item(), err2.max().item())


def test_fused_adamw_is_a_torch_optimizer(cuda_device):
    """The reference's training loop drives the optimizer through torch.optim machinery (LambdaLR scheduler,
    config_files/common/optim.py; checkpointer state_dict / load_state_dict): FusedAdamW must behave as one, resume bit-exactly
    from a saved state — also from one written by torch.optim.AdamW (tensor `step`) — and skip the step on a non-finite norm."""
    from madm_b200.optim import FusedAdamW
    assert issubclass(FusedAdamW, torch.optim.Optimizer)
    kw = dict(lr=5e-3, weight_decay=0.05)
    p1 = [torch.nn.Parameter(t.clone()) for t in _make(cuda_device, 1)[:6]]
    o1 = FusedAdamW(p1, **kw)
    sched = torch.optim.lr_scheduler.LambdaLR(o1, lambda it: 1.0 / (1 + it))
    for it in range(3):
        for p, g in zip(p1, _make(cuda_device, 200 + it)):
            p.grad = g.clone()
        o1.step(clip_grad=1.0)
        sched.step()
    assert abs(o1.param_groups[0]["lr"] - 5e-3 / 4) < 1e-12
    import copy
    sd = copy.deepcopy(o1.state_dict())  # a checkpoint round trip (state_dict() itself hands out references to the live moments)
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    # resume into a fresh optimizer over copies of the parameters, and into torch.optim.AdamW: one more identical step each
    p2 = [torch.nn.Parameter(p.detach().clone()) for p in p1]
    p3 = [torch.nn.Parameter(p.detach().clone()) for p in p1]
    o2 = FusedAdamW(p2, **kw)
    o2.load_state_dict(copy.deepcopy(sd))
    o3 = torch.optim.AdamW(p3, **kw)
    sd3 = {"state": {k: dict(v, step=torch.tensor(float(v["step"]))) for k, v in sd["state"].items()}, "param_groups": sd["param_groups"]}
    o3.load_state_dict(sd3)
    grads = _make(cuda_device, 300)
    for ps in (p1, p2, p3):
        for p, g in zip(ps, grads):
            p.grad = g.clone()
    o1.step(); o2.step(); o3.step()
    for a, b, c in zip(p1, p2, p3):
        assert torch.equal(a, b)
        assert torch.allclose(a, c, rtol=2e-6, atol=1e-7)
    # and back: a state written by torch's AdamW (tensor step) resumes in FusedAdamW
    o4 = FusedAdamW([torch.nn.Parameter(p.detach().clone()) for p in p3], **kw)
    o4.load_state_dict(copy.deepcopy(o3.state_dict()))
    assert all(int(s["step"]) == 4 for s in o4.state.values())
    # non-finite gradient: the update is skipped on the device (GradScaler semantics), nothing becomes NaN
    before = [p.detach().clone() for p in p1]
    m_before = [o1.state[p]["exp_avg"].clone() for p in p1]
    for p, g in zip(p1, grads):
        p.grad = g.clone()
    p1[2].grad.view(-1)[0] = float("inf")
    norm = o1.step(clip_grad=1.0)
    assert not torch.isfinite(norm).item()
    for p, b, m in zip(p1, before, m_before):
        assert torch.equal(p, b) and torch.equal(o1.state[p]["exp_avg"], m)

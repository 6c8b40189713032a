"""SURVEY §8 row f-3, operator level: every backward kernel of the LoRA training step against torch.autograd of the same op in fp32,
through the C ABI (madm_op_*_bwd).  16-bit gradient tensors round to the operand dtype, so tolerances are relative to the tensor's max:
1e-2 for bf16, 3e-3 for fp16 on elementwise / norm kernels, 2e-2 on the tensor-core products."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DTS = [torch.bfloat16, torch.float16]


@pytest.fixture(scope="module")
def ops():
    from madm_b200 import ops as o
    return o


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def tol(dt, base=3e-3):
    return base * (4 if dt == torch.bfloat16 else 1)


def group_sums(x, groups=32):
    """[B,HW,C] -> [B,32,2] (sum, sum of squares) per group: the statistics format the forward kernels hand to the backward."""
    B, HW, C = x.shape
    xg = x.float().reshape(B, HW, groups, C // groups)
    return torch.stack([xg.sum(dim=(1, 3)), (xg * xg).sum(dim=(1, 3))], dim=-1).contiguous()


ACT = {0: lambda t: t, 1: F.silu, 3: F.relu}


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,HW,C0,C1,act,in16", [
    (2, 4096, 320, 0, 1, False), (2, 1024, 640, 320, 1, False), (1, 256, 1280, 1280, 1, False), (2, 64, 1280, 0, 0, False),
    (2, 4096, 320, 0, 1, True), (1, 1024, 128, 0, 3, True), (2, 256, 1920, 0, 1, False), (1, 4096, 512, 0, 0, False), (1, 64, 1280, 640, 1, False)])
def test_groupnorm_bwd(ops, cuda_device, dt, B, HW, C0, C1, act, in16):
    g = torch.Generator(device="cuda").manual_seed(B * HW + C0 + C1 + act)
    C = C0 + C1
    x = torch.randn(B, HW, C, device=cuda_device, generator=g) * 1.5 + 0.3
    if in16:
        x = x.to(dt).float()
    gamma = 1.0 + 0.2 * torch.randn(C, device=cuda_device, generator=g)
    beta = 0.2 * torch.randn(C, device=cuda_device, generator=g)
    dy = (torch.randn(B, HW, C, device=cuda_device, generator=g) * 0.05).to(dt)
    extra = torch.randn(B, HW, C, device=cuda_device, generator=g) * 0.01
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = ACT[act](F.group_norm(xr.permute(0, 2, 1), 32, gr, br, eps=1e-5)).permute(0, 2, 1)
    y.backward(dy.float())
    ref_dx = xr.grad + extra
    x0 = x[..., :C0].contiguous()
    x1 = x[..., C0:].contiguous() if C1 else None
    if in16:
        x0 = x0.to(dt)
    stats = group_sums(x)
    prev0 = torch.randn(B, HW, C0, device=cuda_device, generator=g)
    dx0 = prev0.clone()
    res = ops.groupnorm_bwd(x0, x1, stats, gamma, beta, 1e-5, act, dy, extra=extra, want16=True, want32=True, acc=(True, False), dx0=dx0,
                            want_affine=True)
    assert relerr(res["out16"].float(), ref_dx) < tol(dt)
    assert relerr(res["dx0"] - prev0, ref_dx[..., :C0]) < 1e-4  # fp32 output, accumulated onto what was there
    if C1:
        assert relerr(res["dx1"], ref_dx[..., C0:]) < 1e-4
    assert relerr(res["dgamma"], gr.grad) < 1e-4 and relerr(res["dbeta"], br.grad) < 1e-4
    res2 = ops.groupnorm_bwd(x0, x1, stats, gamma, beta, 1e-5, act, dy, extra=extra, want16=True)
    assert torch.equal(res["out16"], res2["out16"])  # no atomics: bit-identical reruns


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,C", [(8192, 320), (2048, 640), (512, 1280), (77, 1280)])
def test_layernorm_bwd(ops, cuda_device, dt, M, C):
    g = torch.Generator(device="cuda").manual_seed(M + C)
    x = torch.randn(M, C, device=cuda_device, generator=g) * 2 + 0.5
    gamma = 1.0 + 0.2 * torch.randn(C, device=cuda_device, generator=g)
    beta = 0.1 * torch.randn(C, device=cuda_device, generator=g)
    dy = (torch.randn(M, C, device=cuda_device, generator=g) * 0.1).to(dt)
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr, (C,), gamma, beta, 1e-5).backward(dy.float())
    prev = torch.randn(M, C, device=cuda_device, generator=g)
    dx = ops.layernorm_bwd(x, gamma, 1e-5, dy, dx=prev.clone(), accumulate=True)
    assert relerr(dx - prev, xr.grad) < 1e-4
    assert relerr(ops.layernorm_bwd(x, gamma, 1e-5, dy), xr.grad) < 1e-5


@pytest.mark.parametrize("dt", DTS)
def test_geglu_fwd_bwd(ops, cuda_device, dt):
    g = torch.Generator(device="cuda").manual_seed(5)
    M, H = 1000, 1280
    raw = (torch.randn(M, 2 * H, device=cuda_device, generator=g) * 1.5).to(dt)
    dout = (torch.randn(M, H, device=cuda_device, generator=g) * 0.1).to(dt)
    rr = raw.float().requires_grad_(True)
    h, gate = rr.chunk(2, dim=-1)
    ref = h * F.gelu(gate)  # diffusers GEGLU: hidden_states * gelu(gate), exact erf GELU
    ref.backward(dout.float())
    assert relerr(ops.geglu_fwd(raw).float(), ref) < tol(dt)
    assert relerr(ops.geglu_bwd(raw, dout).float(), rr.grad) < tol(dt)


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,heads,d,Nq,Nk", [(2, 8, 40, 1024, 1024), (1, 8, 80, 1024, 1024), (2, 8, 160, 256, 256), (1, 8, 160, 64, 64),
                                             (2, 8, 40, 4096, 77), (2, 8, 160, 64, 77), (1, 8, 80, 1024, 77), (1, 8, 40, 4096, 4096),
                                             (2, 8, 40, 256, 128), (1, 8, 40, 128, 384)])  # (d = 40 with token counts that are multiples of 128: the tcgen05 kernels)
def test_attention_bwd(ops, cuda_device, dt, B, heads, d, Nq, Nk):
    if dt == torch.float16 and Nq == 4096 and Nk == 4096:
        pytest.skip("one dtype is enough at the largest shape")
    g = torch.Generator(device="cuda").manual_seed(d + Nk)
    Cc = heads * d
    if Nq == Nk:  # fused qkv buffer [B, N, 3C] and a gradient buffer of the same layout, as the engine uses
        qkv = torch.randn(B, Nq, 3 * Cc, device=cuda_device, generator=g).to(dt)
        q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
        dqkv = torch.zeros_like(qkv)
        dq, dk, dv = dqkv[..., :Cc], dqkv[..., Cc:2 * Cc], dqkv[..., 2 * Cc:]
        ldq = ldk = lddq = lddk = 3 * Cc
        q_bs = kv_bs = dq_bs = dkv_bs = Nq * 3 * Cc
    else:  # q [B,N,C]; k, v (and their gradients) inside wide per-layer-stacked buffers [B,77,ldkv]
        ldkv = 2 * Cc + 256
        qb = torch.randn(B, Nq, Cc, device=cuda_device, generator=g).to(dt)
        kvb = torch.randn(B, Nk, ldkv, device=cuda_device, generator=g).to(dt)
        q, k, v = qb, kvb[..., 128:128 + Cc], kvb[..., 128 + Cc:128 + 2 * Cc]
        dq = torch.zeros_like(qb)
        dkvb = torch.zeros_like(kvb)
        dk, dv = dkvb[..., 128:128 + Cc], dkvb[..., 128 + Cc:128 + 2 * Cc]
        ldq, ldk, lddq, lddk = Cc, ldkv, Cc, ldkv
        q_bs, kv_bs, dq_bs, dkv_bs = Nq * Cc, Nk * ldkv, Nq * Cc, Nk * ldkv
    dout = (torch.randn(B, Nq, Cc, device=cuda_device, generator=g) * 0.1).to(dt)
    split = lambda t: t.float().reshape(B, -1, heads, d).transpose(1, 2)  # noqa: E731
    qr, kr, vr = (split(t).detach().requires_grad_(True) for t in (q, k, v))
    ref_o = F.scaled_dot_product_attention(qr, kr, vr)
    ref_o.backward(split(dout))
    o = ref_o.transpose(1, 2).reshape(B, Nq, Cc).to(dt).contiguous()  # the forward kernel's 16-bit output
    scale = 1.0 / math.sqrt(d)
    ops.attention_bwd(q, ldq, k, ldk, v, ldk, o, Cc, dout, Cc, dq, lddq, dk, lddk, dv, lddk, B, heads, d, Nq, Nk, q_bs, kv_bs, Nq * Cc, Nq * Cc,
                      dq_bs, dkv_bs, scale)
    merge = lambda t: t.transpose(1, 2).reshape(B, -1, Cc)  # noqa: E731
    for name, got, ref in (("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
        e = relerr(got.float(), merge(ref))
        print(f"attention_bwd {dt} d={d} {Nq}x{Nk} {name}: {e:.2e}")
        assert e < (3e-2 if dt == torch.bfloat16 else 1e-2), name
    if Nq != Nk:  # nothing outside the k / v gradient columns of the stacked buffer was touched
        assert torch.count_nonzero(dkvb[..., :128]) == 0 and torch.count_nonzero(dkvb[..., 128 + 2 * Cc:]) == 0


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,N,K", [(8192, 320, 320), (154, 640, 768), (2048, 640, 640), (512, 1280, 1280), (154, 1280, 768), (1000, 320, 768), (16384, 320, 320)])
def test_lora_grads_fused(ops, cuda_device, dt, M, N, K):
    """Both LoRA factor gradients of one wrapped linear in one kernel (+ reduce): dB = s dY^T (X A^T), dA = s (dY B)^T X against fp32 matmuls of the
    same 16-bit operands (the skinny products are rounded to 16 bits in between, as on the unfused path); dY / X are column slices of wider
    buffers, M is ragged against the 64-row chunks, and more chunks than CTAs (M = 16384) exercise the per-CTA chunk loop."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    wide_y = torch.randn(M, N + 64, device=cuda_device, generator=g).to(dt)
    wide_x = torch.randn(M, K + 128, device=cuda_device, generator=g).to(dt)
    dy, x = wide_y[:, 64:], wide_x[:, 128:]
    A = torch.randn(16, K, device=cuda_device, generator=g) / math.sqrt(K)
    B = torch.randn(N, 16, device=cuda_device, generator=g) * 0.05
    gA, gB = ops.lora_grads(x, dy, A, B, alpha=0.25)
    a16, b16 = A.to(dt).float(), B.to(dt).float()
    U = (x.float() @ a16.t()).to(dt).float()
    V = (dy.float() @ b16).to(dt).float()
    assert relerr(gB, 0.25 * dy.float().t() @ U) < 2e-3
    assert relerr(gA, 0.25 * V.t() @ x.float()) < 2e-3
    gA2, gB2 = ops.lora_grads(x, dy, A, B, alpha=0.25)
    assert torch.equal(gA, gA2) and torch.equal(gB, gB2)  # deterministic


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("M,N,K", [(8192, 320, 16), (154, 1280, 16), (2048, 640, 64), (32768, 512, 128), (8192, 128, 512), (100, 64, 64)])
def test_wgrad_linear(ops, cuda_device, dt, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    wide = torch.randn(M, N + 64, device=cuda_device, generator=g).to(dt)  # dY is a column slice of a wider buffer (fused qkv gradient)
    dy = wide[:, 64:]
    x = torch.randn(M, K, device=cuda_device, generator=g).to(dt)
    ref = 0.5 * dy.float().t() @ x.float()
    got = ops.wgrad(dy, x, N, K, alpha=0.5, lda=N + 64)
    assert relerr(got, ref) < 2e-3
    got_t = ops.wgrad(dy, x, N, K, alpha=0.5, lda=N + 64, transpose_out=True)
    assert torch.equal(got_t, got.t().contiguous())
    assert torch.equal(got, ops.wgrad(dy, x, N, K, alpha=0.5, lda=N + 64))  # deterministic split reduction


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 128, 128), (1, 16, 16, 128, 128), (2, 128, 128, 128, 128)])
def test_wgrad_conv3x3(ops, cuda_device, dt, B, H, W, Cin, Cout):
    g = torch.Generator(device="cuda").manual_seed(H + Cin)
    x = torch.randn(B, H, W, Cin, device=cuda_device, generator=g).to(dt)
    dy = (torch.randn(B, H, W, Cout, device=cuda_device, generator=g) * 0.1).to(dt)
    w = torch.zeros(Cout, Cin, 3, 3, device=cuda_device, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    got = ops.wgrad(dy.reshape(-1, Cout), x, Cout, Cin, taps=9, geom=(B, H, W), M=B * H * W, ldb=Cin)
    assert got.shape == (Cout, Cin, 3, 3)
    assert relerr(got, w.grad) < 2e-3


@pytest.mark.parametrize("dt", DTS)
def test_strided_conv_dgrad_via_zero_stuffing(ops, cuda_device, dt):
    """Input gradient of the UNet's stride-2 downsample conv (padding 1): zero-stuff dY to the input grid, then the stride-1 dgrad
    implicit GEMM with the mirrored-tap packed weight."""
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, C = 2, 32, 32, 64
    w = torch.randn(C, C, 3, 3, device=cuda_device, generator=g) * 0.05
    dy = (torch.randn(B, H // 2, W // 2, C, device=cuda_device, generator=g) * 0.1).to(dt)
    x = torch.zeros(B, C, H, W, device=cuda_device, requires_grad=True)
    F.conv2d(x, w, stride=2, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    z = ops.zero_stuff2x(dy)
    assert torch.equal(z[:, ::2, ::2], dy) and torch.count_nonzero(z[:, 1::2]) == 0 and torch.count_nonzero(z[:, :, 1::2]) == 0
    wd = ops.pack_conv_dgrad(w, dtype=dt)
    got = torch.empty(B * H * W, C, device=cuda_device)
    ops.gemm([ops.make_seg(z, B, H, W, C, taps=ops.taps_3x3())], B * H * W, C, wd, out_f32=got, ldo32=C)
    assert relerr(got.reshape(B, H, W, C), x.grad.permute(0, 2, 3, 1)) < tol(dt, 5e-3)


def test_small_movement_kernels(ops, cuda_device):
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(2, 4096, 320, device=cuda_device, generator=g).to(torch.bfloat16)
    out = torch.zeros(2, 1000, device=cuda_device)
    ops.colsum_per_image(x, out=out, col_off=400)
    assert relerr(out[:, 400:720], x.float().sum(1)) < 1e-5 and torch.count_nonzero(out[:, :400]) == 0 and torch.count_nonzero(out[:, 720:]) == 0
    u = torch.randn(2, 32, 32, 64, device=cuda_device, generator=g)
    ref = u.reshape(2, 16, 2, 16, 2, 64).sum(dim=(2, 4))
    prev = torch.randn(2, 16, 16, 64, device=cuda_device, generator=g)
    assert relerr(ops.sum2x2(u), ref) < 1e-6
    assert relerr(ops.sum2x2(u, out=prev.clone(), accumulate=True) - prev, ref) < 1e-5
    o = torch.randn(2, 70, 17, 19, device=cuda_device, generator=g)
    do = torch.randn(2, 70, 17, 19, device=cuda_device, generator=g)
    dz = ops.relu_bwd_nchw(do, o, scale=2.0, dtype=torch.float16)
    ref = (2.0 * do * (o > 0)).permute(0, 2, 3, 1).reshape(2, 17 * 19, 70)
    assert torch.equal(dz, ref.to(torch.float16))

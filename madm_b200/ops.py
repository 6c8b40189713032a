"""Operator-level wrappers over the C ABI (``madm_op_*``): torch tensors are used only as device memory.

These are the kernels the engine is composed of; the parity tests drive them through the same C ABI the
engine uses internally.  No PyTorch compute happens here.
"""
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import MadmGemmArgs, MadmGemmSeg


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _dt(dtype) -> int:
    """torch dtype of the 16-bit operand tensors -> MADM_DTYPE_*."""
    if dtype == torch.float16:
        return _lib.DTYPE_FP16
    if dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    raise _lib.MadmError(f"operand dtype must be torch.float16 or torch.bfloat16, got {dtype}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def taps_3x3():
    return [(kx - 1, ky - 1, 0) for ky in range(3) for kx in range(3)]  # (dx, dy, b_off)


def taps_stride2(B: int, pad1: bool):
    """Tap table of a stride-2 3x3 conv over the space-to-depth tensor [4][B][H/2][W/2][C]."""
    out = []
    for ky in range(3):
        for kx in range(3):
            if pad1:
                py, dy = (0 if ky == 1 else 1), (-1 if ky == 0 else 0)
                px, dx = (0 if kx == 1 else 1), (-1 if kx == 0 else 0)
            else:
                py, dy = (1 if ky == 1 else 0), (1 if ky == 2 else 0)
                px, dx = (1 if kx == 1 else 0), (1 if kx == 2 else 0)
            out.append((dx, dy, (py * 2 + px) * B))
    return out


def make_seg(a: torch.Tensor, Bt: int, H: int, W: int, Cc: int, ld: int = 0, taps: Optional[Sequence] = None) -> MadmGemmSeg:
    s = MadmGemmSeg()
    s.a = a.data_ptr()
    s.Bt, s.H, s.W, s.C, s.ld = Bt, H, W, Cc, ld
    taps = taps or [(0, 0, 0)]
    s.ntaps = len(taps)
    for i, (dx, dy, bo) in enumerate(taps):
        s.dx[i], s.dy[i], s.b_off[i] = dx, dy, bo
    return s


def gemm(segs: Sequence[MadmGemmSeg], M: int, N: int, w: torch.Tensor, *, Nw: int = 0, ldw: int = 0,
         bias=None, rowbias=None, rows_per_img: int = 1, ld_rowbias: int = 0, residual=None, ldr: int = 0,
         out_f32=None, ldo32: int = 0, out_bf16=None, ldo16: int = 0, act: int = 0, alpha: float = 1.0, bn: int = 0,
         colstats=None, stat_rows: int = 0, mt: int = 0, s2d_hw=None, pair: int = 0):
    a = MadmGemmArgs()
    a.nseg = len(segs)
    for i, s in enumerate(segs):
        a.seg[i] = s
    a.M, a.N, a.Nw, a.ldw = M, N, Nw or N, ldw
    a.w = w.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.rowbias = rowbias.data_ptr() if rowbias is not None else None
    a.rows_per_img, a.ld_rowbias = rows_per_img, ld_rowbias
    a.residual = residual.data_ptr() if residual is not None else None
    a.ldr = ldr
    a.res16 = 1 if (residual is not None and residual.dtype != torch.float32) else 0
    a.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
    a.ldo32 = ldo32
    a.out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
    a.ldo16 = ldo16
    a.act, a.alpha, a.bn = act, alpha, bn
    a.dtype = _dt(w.dtype)
    a.colstats = colstats.data_ptr() if colstats is not None else None
    a.stat_rows = stat_rows
    a.mt = mt
    a.pair = pair
    if s2d_hw is not None:
        a.s2d_H, a.s2d_W = s2d_hw
    lib = _lib.load()
    _lib.check(lib.madm_op_gemm(C.byref(a), _stream()), None, "madm_op_gemm")


def groupnorm(x0, x1, B, HW, gamma, beta, eps, act, y, raw=None):
    lib = _lib.load()
    C0 = x0.shape[-1]
    C1 = x1.shape[-1] if x1 is not None else 0
    stats = torch.empty(lib.madm_op_groupnorm_scratch_floats(B, HW, C0 + C1), dtype=torch.float32, device=x0.device)
    in16 = 0 if x0.dtype == torch.float32 else 1
    _lib.check(lib.madm_op_groupnorm(_ptr(x0), C0, _ptr(x1), C1, B, HW, in16, _ptr(gamma), _ptr(beta), eps, act, _ptr(stats),
                                     _ptr(y), _ptr(raw), _dt(y.dtype), _stream()), None, "madm_op_groupnorm")
    return stats


def groupnorm_from_colstats(x, B, HW, colstats, stat_rows, gamma, beta, eps, act, y):
    lib = _lib.load()
    Cc = x.shape[-1]
    scratch = torch.empty(B * 32 * 64, dtype=torch.float32, device=x.device)
    in16 = 0 if x.dtype == torch.float32 else 1
    _lib.check(lib.madm_op_groupnorm_from_colstats(_ptr(x), Cc, B, HW, in16, _ptr(colstats), stat_rows, _ptr(gamma), _ptr(beta), eps, act,
                                                   _ptr(scratch), _ptr(y), _dt(y.dtype), _stream()), None, "madm_op_groupnorm_from_colstats")


def layernorm(x, gamma, beta, eps, y):
    lib = _lib.load()
    M, Cc = x.shape
    in16 = 0 if x.dtype == torch.float32 else 1
    _lib.check(lib.madm_op_layernorm(_ptr(x), in16, M, Cc, _ptr(gamma), _ptr(beta), eps, _ptr(y), _dt(y.dtype), _stream()), None, "madm_op_layernorm")


def softmax_rows(s, p):
    lib = _lib.load()
    R, L = s.shape
    _lib.check(lib.madm_op_softmax_rows(_ptr(s), R, L, _ptr(p), _dt(p.dtype), _stream()), None, "madm_op_softmax_rows")


def attention(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, d, Nq, Nk, q_bs, kv_bs, o_bs, scale, impl=0):
    lib = _lib.load()
    _lib.check(lib.madm_op_attention(_ptr(q), ldq, _ptr(k), ldk, _ptr(v), ldv, _ptr(o), ldo, B, heads, d, Nq, Nk, q_bs, kv_bs,
                                     o_bs, scale, _dt(o.dtype), impl, _stream()), None, "madm_op_attention")


def pack_linear(w, lora_a=None, lora_b=None, scale=0.0, out=None, ldo=0, dtype=torch.float16):
    lib = _lib.load()
    N, K = w.shape
    r = lora_a.shape[0] if lora_a is not None else 0
    if out is None:
        out = torch.empty(N, K, dtype=dtype, device=w.device)
    _lib.check(lib.madm_op_pack_linear(_ptr(w), N, K, _ptr(lora_a), _ptr(lora_b), r, scale, _ptr(out), ldo, _dt(out.dtype), _stream()), None,
               "madm_op_pack_linear")
    return out


def pack_conv(w, Cpad=0, out=None, ldo=0, dtype=torch.float16):
    lib = _lib.load()
    N, Cc, kh, kw = w.shape
    taps = kh * kw
    Cpad = Cpad or (Cc + 63) // 64 * 64
    if out is None:
        out = torch.empty(N, taps * Cpad, dtype=dtype, device=w.device)
    _lib.check(lib.madm_op_pack_conv(_ptr(w), N, Cc, taps, Cpad, _ptr(out), ldo, _dt(out.dtype), _stream()), None, "madm_op_pack_conv")
    return out


def pack_conv_dgrad(w, CoPad=0, dtype=torch.float16):
    """[Cout,Cin,kh,kw] -> [Cin, taps*CoPad] with mirrored taps: the B operand of the conv's input-gradient GEMM."""
    lib = _lib.load()
    Cout, Cin, kh, kw = w.shape
    taps = kh * kw
    CoPad = CoPad or (Cout + 63) // 64 * 64
    out = torch.empty(Cin, taps * CoPad, dtype=dtype, device=w.device)
    _lib.check(lib.madm_op_pack_conv_dgrad(_ptr(w), Cout, Cin, taps, CoPad, _ptr(out), 0, _dt(dtype), _stream()), None, "madm_op_pack_conv_dgrad")
    return out


def pack_linear_dgrad(w, lora_a=None, lora_b=None, scale=0.0, dtype=torch.float16):
    """[N,K] (+ scale * B @ A) -> its transpose [K,N]: the B operand of the linear's input-gradient GEMM."""
    lib = _lib.load()
    N, K = w.shape
    out = torch.empty(K, N, dtype=dtype, device=w.device)
    r = lora_a.shape[0] if lora_a is not None else 0
    _lib.check(lib.madm_op_pack_linear_dgrad(_ptr(w), N, K, _ptr(lora_a), _ptr(lora_b), r, float(scale), _ptr(out), 0, _dt(dtype), _stream()),
               None, "madm_op_pack_linear_dgrad")
    return out


def pack_geglu(w, bias, dtype=torch.float16):
    lib = _lib.load()
    N2, K = w.shape
    out = torch.empty(N2, K, dtype=dtype, device=w.device)
    ob = torch.empty(N2, dtype=torch.float32, device=w.device)
    _lib.check(lib.madm_op_pack_geglu(_ptr(w), _ptr(bias), N2 // 2, K, _ptr(out), _ptr(ob), _dt(dtype), _stream()), None, "madm_op_pack_geglu")
    return out, ob


def space_to_depth(x, dtype=torch.float16):
    lib = _lib.load()
    B, H, W, Cc = x.shape
    out = torch.empty(4, B, H // 2, W // 2, Cc, dtype=dtype, device=x.device)
    _lib.check(lib.madm_op_space_to_depth(_ptr(x), B, H, W, Cc, _ptr(out), _dt(dtype), _stream()), None, "madm_op_space_to_depth")
    return out


def upsample2x(x, dtype=torch.float16):
    lib = _lib.load()
    B, H, W, Cc = x.shape
    out = torch.empty(B, 2 * H, 2 * W, Cc, dtype=dtype, device=x.device)
    _lib.check(lib.madm_op_upsample2x(_ptr(x), B, H, W, Cc, _ptr(out), _dt(dtype), _stream()), None, "madm_op_upsample2x")
    return out


def nchw_to_nhwc16(x, dtype=torch.float16):
    lib = _lib.load()
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, Cc, dtype=dtype, device=x.device)
    _lib.check(lib.madm_op_nchw_to_nhwc16(_ptr(x), B, Cc, H * W, _ptr(out), _dt(dtype), _stream()), None, "madm_op_nchw_to_nhwc16")
    return out


def bilinear_resize(x, Hd, Wd, out=None, pitch=0):
    """x: 16-bit NHWC [B,Hs,Ws,C] -> [B,Hd,Wd,C] (F.interpolate bilinear, align_corners=False), optionally into a wider buffer."""
    lib = _lib.load()
    B, Hs, Ws, Cc = x.shape
    if out is None:
        out = torch.empty(B, Hd, Wd, Cc, dtype=x.dtype, device=x.device)
    _lib.check(lib.madm_op_bilinear_resize(_ptr(x), B, Hs, Ws, Cc, _ptr(out), Hd, Wd, pitch or Cc, _dt(x.dtype), _stream()), None,
               "madm_op_bilinear_resize")
    return out


def depthwise3x3(x, w9, shift, dilation):
    """x: 16-bit NHWC; w9 fp32 [9,C]; shift fp32 [C] -> relu(depthwise_conv(x) + shift), 16-bit NHWC."""
    lib = _lib.load()
    B, H, W, Cc = x.shape
    out = torch.empty_like(x)
    _lib.check(lib.madm_op_depthwise3x3(_ptr(x), B, H, W, Cc, dilation, _ptr(w9), _ptr(shift), _ptr(out), _dt(x.dtype), _stream()), None,
               "madm_op_depthwise3x3")
    return out


def preprocess_image(img, size, divisibility=64):
    """T.Resize(size, BILINEAR) (no antialias: torchvision 0.16.1 on tensors) + zero pad to a multiple of `divisibility`
    (FeatureExtractorBackbone.preprocess_image, feature_extractor.py:140-146).  size=None: pad only."""
    lib = _lib.load()
    x = img.to(torch.float32).contiguous()
    B, Cc, Hs, Ws = x.shape
    Hr, Wr = (Hs, Ws) if size is None else (int(size[0]), int(size[1]))
    Hd, Wd = Hr + (-Hr) % divisibility, Wr + (-Wr) % divisibility
    out = torch.empty(B, Cc, Hd, Wd, dtype=torch.float32, device=x.device)
    _lib.check(lib.madm_op_preprocess_image(_ptr(x), B * Cc, Hs, Ws, Hr, Wr, Hd, Wd, _ptr(out), _stream()), None, "madm_op_preprocess_image")
    return out


def slide_merge(feats, nwin, wins_yx, Hf, Wf):
    """feats [nwin*n,C,hf,wf] fp32 (window-major), wins_yx [(y1,x1)] in feature pixels -> [n,C,Hf,Wf] mean over covering windows."""
    lib = _lib.load()
    NB, Cc, hf, wf = feats.shape
    n = NB // nwin
    wins = torch.tensor(list(wins_yx), dtype=torch.int32, device=feats.device).reshape(-1).contiguous()
    out = torch.empty(n, Cc, Hf, Wf, dtype=torch.float32, device=feats.device)
    _lib.check(lib.madm_op_slide_merge(_ptr(feats.contiguous()), nwin, n, Cc, hf, wf, _ptr(wins), Hf, Wf, _ptr(out), _stream()), None,
               "madm_op_slide_merge")
    return out


def image_im2col(img, range_flag=None, dtype=torch.float16):
    lib = _lib.load()
    B, _, H, W = img.shape
    out = torch.empty(B * H * W, 64, dtype=dtype, device=img.device)
    _lib.check(lib.madm_op_image_im2col(_ptr(img), B, H, W, _ptr(out), _ptr(range_flag), _dt(dtype), _stream()), None, "madm_op_image_im2col")
    return out


def gn_add_relu_nchw(a, ga, ba, s, gs, bs, eps, B, HW, Cc):
    lib = _lib.load()
    stats = torch.empty(lib.madm_op_groupnorm_scratch_floats(B, HW, Cc), dtype=torch.float32, device=a.device)
    out = torch.empty(B, Cc, HW, dtype=torch.float32, device=a.device)
    _lib.check(lib.madm_op_gn_add_relu_nchw(_ptr(a), _ptr(ga), _ptr(ba), _ptr(s), _ptr(gs), _ptr(bs), 1 if gs is not None else 0,
                                            eps, B, HW, Cc, _ptr(stats), _ptr(out), _stream()), None, "madm_op_gn_add_relu_nchw")
    return out


# ---------------------------------------------------------------------------------------------- backward pass
def groupnorm_bwd(x0, x1, stats, gamma, beta, eps, act, dy, *, extra=None, want16=True, want32=False, acc=(False, False),
                  dx0=None, dx1=None, want_affine=False):
    """GroupNorm(32)(+act) backward on NHWC inputs ([B,HW,C0] (+[B,HW,C1]) fp32 or 16-bit); stats [B,32,2] group sums of x.
    Returns dict(out16, dx0, dx1, dgamma, dbeta)."""
    lib = _lib.load()
    B, HW, C0 = x0.shape
    C1 = x1.shape[2] if x1 is not None else 0
    Cc = C0 + C1
    dev = x0.device
    in16 = 1 if x0.dtype in (torch.float16, torch.bfloat16) else 0
    scratch = torch.empty(lib.madm_op_groupnorm_bwd_scratch_floats(B, HW, Cc), dtype=torch.float32, device=dev)
    out16 = torch.empty(B, HW, Cc, dtype=dy.dtype, device=dev) if want16 else None
    if want32:
        dx0 = dx0 if dx0 is not None else torch.empty(B, HW, C0, dtype=torch.float32, device=dev)
        if C1:
            dx1 = dx1 if dx1 is not None else torch.empty(B, HW, C1, dtype=torch.float32, device=dev)
    dg = torch.empty(Cc, dtype=torch.float32, device=dev) if want_affine else None
    db = torch.empty(Cc, dtype=torch.float32, device=dev) if want_affine else None
    _lib.check(lib.madm_op_groupnorm_bwd(_ptr(x0), C0, _ptr(x1), C1, B, HW, in16, _ptr(stats), _ptr(gamma), _ptr(beta), float(eps), int(act), _ptr(dy),
                                         _ptr(extra), _ptr(scratch), _ptr(out16), _ptr(dx0), int(acc[0]), _ptr(dx1), int(acc[1]), _ptr(dg), _ptr(db),
                                         _dt(dy.dtype), _stream()), None, "madm_op_groupnorm_bwd")
    return dict(out16=out16, dx0=dx0, dx1=dx1, dgamma=dg, dbeta=db)


def layernorm_bwd(x, gamma, eps, dy, dx=None, accumulate=False):
    lib = _lib.load()
    M, Cc = x.shape
    if dx is None:
        dx = torch.empty(M, Cc, dtype=torch.float32, device=x.device)
    _lib.check(lib.madm_op_layernorm_bwd(_ptr(x), M, Cc, _ptr(gamma), float(eps), _ptr(dy), _ptr(dx), int(accumulate), _dt(dy.dtype), _stream()), None,
               "madm_op_layernorm_bwd")
    return dx


def geglu_fwd(raw):
    lib = _lib.load()
    M, H2 = raw.shape
    out = torch.empty(M, H2 // 2, dtype=raw.dtype, device=raw.device)
    _lib.check(lib.madm_op_geglu_fwd(_ptr(raw), M, H2 // 2, _ptr(out), _dt(raw.dtype), _stream()), None, "madm_op_geglu_fwd")
    return out


def geglu_bwd(raw, dout):
    lib = _lib.load()
    M, H2 = raw.shape
    draw = torch.empty_like(raw)
    _lib.check(lib.madm_op_geglu_bwd(_ptr(raw), _ptr(dout), M, H2 // 2, _ptr(draw), _dt(raw.dtype), _stream()), None, "madm_op_geglu_bwd")
    return draw


def attention_bwd(q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, dq, lddq, dk, lddk, dv, lddv, B, heads, d, Nq, Nk, q_bs, kv_bs, o_bs, do_bs, dq_bs,
                  dkv_bs, scale):
    lib = _lib.load()
    scratch = torch.empty(lib.madm_op_attention_bwd_scratch_floats(B, heads, d, Nq, Nk), dtype=torch.float32, device=o.device)
    _lib.check(lib.madm_op_attention_bwd(_ptr(q), ldq, _ptr(k), ldk, _ptr(v), ldv, _ptr(o), ldo, _ptr(dout), lddo, _ptr(dq), lddq, _ptr(dk), lddk,
                                         _ptr(dv), lddv, B, heads, d, Nq, Nk, q_bs, kv_bs, o_bs, do_bs, dq_bs, dkv_bs, float(scale), _ptr(scratch),
                                         _dt(o.dtype), _stream()), None, "madm_op_attention_bwd")


def wgrad(dy, x, N, K, *, taps=1, geom=(0, 0, 0), alpha=1.0, transpose_out=False, lda=None, ldb=None, M=None):
    """dW[n,k] = alpha * sum_m dY[m,n] X[m,k] (taps=9: X is NHWC [B,H,W,K], result in conv layout [N,K,3,3])."""
    lib = _lib.load()
    M = M if M is not None else dy.shape[0]
    lda = lda if lda is not None else dy.stride(0)
    ldb = ldb if ldb is not None else (x.stride(0) if x.dim() == 2 else x.shape[-1])
    shape = (N, K, 3, 3) if taps == 9 else ((K, N) if transpose_out else (N, K))
    out = torch.empty(shape, dtype=torch.float32, device=dy.device)
    scratch = torch.empty(lib.madm_op_wgrad_scratch_floats(M, N, K, taps), dtype=torch.float32, device=dy.device)
    _lib.check(lib.madm_op_wgrad(_ptr(dy), lda, _ptr(x), ldb, M, N, K, taps, geom[0], geom[1], geom[2], float(alpha), _ptr(out), int(transpose_out),
                                 _ptr(scratch), _dt(dy.dtype), _stream()), None, "madm_op_wgrad")
    return out


def lora_grads(x, dy, A, B, alpha=1.0, M=None, ldx=None, ldy=None):
    """Fused LoRA factor gradients of y = (W + s B A) x: returns (gA [16,K], gB [N,16]) = alpha * ((dY B)^T X, dY^T (X A^T)).
    x [M,K], dy [M,N] 16-bit (row pitches ldx / ldy); A [16,K], B [N,16] fp32 or 16-bit (converted to the operand dtype here)."""
    lib = _lib.load()
    M = M if M is not None else x.shape[0]
    K, N = A.shape[1], B.shape[0]
    a16 = A.to(x.dtype).contiguous()
    bt16 = B.t().to(x.dtype).contiguous()
    gA = torch.empty(16, K, dtype=torch.float32, device=x.device)
    gB = torch.empty(N, 16, dtype=torch.float32, device=x.device)
    n = lib.madm_op_lora_grads_scratch_floats(M, N, K)
    if n < 0:
        raise ValueError(f"lora_grads: unsupported N / K ({N}, {K})")
    scratch = torch.empty(n, dtype=torch.float32, device=x.device)
    _lib.check(lib.madm_op_lora_grads(_ptr(x), ldx if ldx is not None else x.stride(0), _ptr(dy), ldy if ldy is not None else dy.stride(0), _ptr(a16),
                                      _ptr(bt16), M, N, K, float(alpha), _ptr(gA), _ptr(gB), _ptr(scratch), _dt(x.dtype), _stream()), None,
               "madm_op_lora_grads")
    return gA, gB


def colsum_per_image(x, out=None, col_off=0):
    lib = _lib.load()
    B, HW, Cc = x.shape
    if out is None:
        out = torch.empty(B, Cc, dtype=torch.float32, device=x.device)
    _lib.check(lib.madm_op_colsum_per_image(_ptr(x), B, HW, Cc, C.c_void_p(out.data_ptr() + 4 * col_off), out.stride(0), _dt(x.dtype), _stream()), None,
               "madm_op_colsum_per_image")
    return out


def zero_stuff2x(x):
    lib = _lib.load()
    B, h, w, Cc = x.shape
    out = torch.empty(B, 2 * h, 2 * w, Cc, dtype=x.dtype, device=x.device)
    _lib.check(lib.madm_op_zero_stuff2x(_ptr(x), B, h, w, Cc, _ptr(out), _stream()), None, "madm_op_zero_stuff2x")
    return out


def sum2x2(x, out=None, accumulate=False):
    lib = _lib.load()
    B, H2, W2, Cc = x.shape
    if out is None:
        out = torch.empty(B, H2 // 2, W2 // 2, Cc, dtype=torch.float32, device=x.device)
    _lib.check(lib.madm_op_sum2x2(_ptr(x), B, H2 // 2, W2 // 2, Cc, _ptr(out), int(accumulate), _stream()), None, "madm_op_sum2x2")
    return out


def relu_bwd_nchw(dout, out, scale=1.0, dtype=torch.bfloat16):
    lib = _lib.load()
    B, Cc, H, W = out.shape
    dz = torch.empty(B, H * W, Cc, dtype=dtype, device=out.device)
    _lib.check(lib.madm_op_relu_bwd_nchw(_ptr(dout), _ptr(out), B, Cc, H * W, float(scale), _ptr(dz), _dt(dtype), _stream()), None,
               "madm_op_relu_bwd_nchw")
    return dz

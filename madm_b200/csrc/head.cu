// Data-movement kernels of the DAFormer head stage (reference modeling/sem_seg_head/daformer_head.py:702-749):
//   nchw_to_nhwc16        the public feature dict (fp32 NCHW s2..s5) -> 16-bit NHWC GEMM operands
//   bilinear_resize_nhwc  mmseg `resize(mode='bilinear', align_corners=False)` of the MLP embeds to the s2 grid, written straight
//                         into its channel slice of the concat buffer
//   depthwise3x3_nhwc     DepthwiseSeparableConvModule's depthwise 3x3 (dilation 6/12/18) with eval-mode BatchNorm folded in + ReLU
// All HBM-bound; the dense parts of the head (MLP embeds, ASPP 1x1 / pointwise convs, 3x3 bottleneck, classifier) run on
// gemm_tc_kernel with the BatchNorm scale folded into the packed weights.
#include "cvt.cuh"
#include "kernels.h"

namespace madm {

// ------------------------------------------------------------------ fp32 NCHW -> 16-bit NHWC (64 ch x 32 px smem transpose)
// reads: 32 consecutive pixels of one channel (128 B); writes: 64 consecutive channels of one pixel as 32 packed pairs (128 B)
__global__ void nchw_to_nhwc16_kernel(const float* __restrict__ x, int C, int HW, int fp16, uint16_t* __restrict__ out) {
  __shared__ float tile[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 64; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    tile[r][tx] = (c < C && p < HW) ? __ldg(x + (size_t(b) * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + 2 * tx;
    if (p >= HW) continue;
    uint16_t* dst = out + (size_t(b) * HW + p) * C + c;
    if (c + 1 < C) *reinterpret_cast<uint32_t*>(dst) = pack2_16(tile[2 * tx][r], tile[2 * tx + 1][r], fp16);
    else if (c < C) *dst = cvt_16(tile[2 * tx][r], fp16);
  }
}

const char* nchw_to_nhwc16(const float* x, int B, int C, int HW, void* out, int fp16, cudaStream_t st) {
  if (C % 2) return "nchw_to_nhwc16: C must be even";
  dim3 grid((HW + 31) / 32, (C + 63) / 64, B);
  nchw_to_nhwc16_kernel<<<grid, dim3(32, 8), 0, st>>>(x, C, HW, fp16, reinterpret_cast<uint16_t*>(out));
  return cudaGetLastError() == cudaSuccess ? nullptr : "nchw_to_nhwc16 launch failed";
}

// ------------------------------------------------------------------ bilinear resize, align_corners = False (F.interpolate semantics)
// One thread = 8 channels of one output pixel: four 16-byte loads, fp32 blend, one 16-byte store at channel pitch ldd.
__device__ __forceinline__ void unpack8(const uint4& u, int fp16, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f;
    if (fp16) f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

__global__ void bilinear_resize_nhwc_kernel(const uint16_t* __restrict__ src, int Hs, int Ws, int C, uint16_t* __restrict__ dst, int Hd,
                                            int Wd, int ldd, float sy, float sx, long total, int fp16) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Q = C >> 3;
  const int c = int(i % Q) * 8;
  const long pix = i / Q;
  const int xd = int(pix % Wd), yd = int((pix / Wd) % Hd), b = int(pix / (long(Wd) * Hd));
  // area_pixel_compute_source_index: src = (dst + 0.5) * scale - 0.5, clamped at 0
  const float fy = fmaxf((yd + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((xd + 0.5f) * sx - 0.5f, 0.f);
  const int y0 = min(int(fy), Hs - 1), x0 = min(int(fx), Ws - 1);
  const int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
  const float ly = fy - float(y0), lx = fx - float(x0);
  const uint16_t* base = src + size_t(b) * Hs * Ws * C + c;
  float a[8], bq[8], cq[8], d[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t(y0) * Ws + x0) * C)), fp16, a);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t(y0) * Ws + x1) * C)), fp16, bq);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t(y1) * Ws + x0) * C)), fp16, cq);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t(y1) * Ws + x1) * C)), fp16, d);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  float o[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) o[t] = w00 * a[t] + w01 * bq[t] + w10 * cq[t] + w11 * d[t];
  uint4 pk;
  pk.x = pack2_16(o[0], o[1], fp16); pk.y = pack2_16(o[2], o[3], fp16); pk.z = pack2_16(o[4], o[5], fp16); pk.w = pack2_16(o[6], o[7], fp16);
  *reinterpret_cast<uint4*>(dst + (size_t(b) * Hd * Wd + size_t(yd) * Wd + xd) * ldd + c) = pk;
}

const char* bilinear_resize_nhwc16(const void* src, int B, int Hs, int Ws, int C, void* dst, int Hd, int Wd, int ldd, int fp16, cudaStream_t st) {
  if (C % 8 || ldd % 8) return "bilinear_resize: C and the destination pitch must be multiples of 8";
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) return "bilinear_resize: pointers must be 16B aligned";
  const long total = long(B) * Hd * Wd * (C / 8);
  bilinear_resize_nhwc_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(src), Hs, Ws, C,
                                                                            reinterpret_cast<uint16_t*>(dst), Hd, Wd, ldd, float(Hs) / float(Hd),
                                                                            float(Ws) / float(Wd), total, fp16);
  return cudaGetLastError() == cudaSuccess ? nullptr : "bilinear_resize launch failed";
}

// ------------------------------------------------------------------ depthwise 3x3 (dilated, zero padding = dilation) + BN shift + ReLU
// w9: fp32 [9][C] with the BatchNorm scale folded in (tap index = ky*3 + kx); shift: fp32 [C].
// One thread = 8 channels of one image column x, for the rows of one residue class y = r (mod dil): walking y in steps of the
// dilation, output row y needs input rows y-dil, y, y+dil, two of which were already loaded for the previous output.  So each
// output costs 3 (not 9) 16-byte loads (issued one row ahead), and the 72 weights + 8 shifts of the thread's channels stay in registers for the walk.
__global__ void __launch_bounds__(128) depthwise3x3_nhwc_kernel(const uint16_t* __restrict__ src, int H, int W, int C, int dil,
                                                                const float* __restrict__ w9, const float* __restrict__ shift,
                                                                uint16_t* __restrict__ dst, long total, int classes, int fp16) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Q = C >> 3;
  const int c = int(i % Q) * 8;
  long rest = i / Q;
  const int x = int(rest % W); rest /= W;
  const int r = int(rest % classes);  // residue class of y (mod dil); classes = min(dil, H)
  const int b = int(rest / classes);
  float w[9][8], sh[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w9 + size_t(t) * C + c)), w1 = __ldg(reinterpret_cast<const float4*>(w9 + size_t(t) * C + c + 4));
    w[t][0] = w0.x; w[t][1] = w0.y; w[t][2] = w0.z; w[t][3] = w0.w; w[t][4] = w1.x; w[t][5] = w1.y; w[t][6] = w1.z; w[t][7] = w1.w;
  }
  {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(shift + c)), s1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
    sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
  }
  const uint16_t* base = src + size_t(b) * H * W * C + c;
  const bool xl = x - dil >= 0, xr = x + dil < W;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  auto load_row = [&](int yy, uint4 (&row)[3]) {  // the three x taps of input row yy (zero outside the image)
    if (yy < 0 || yy >= H) { row[0] = zero; row[1] = zero; row[2] = zero; return; }
    const uint16_t* p = base + (size_t(yy) * W + x) * C;
    row[0] = xl ? __ldg(reinterpret_cast<const uint4*>(p - size_t(dil) * C)) : zero;
    row[1] = __ldg(reinterpret_cast<const uint4*>(p));
    row[2] = xr ? __ldg(reinterpret_cast<const uint4*>(p + size_t(dil) * C)) : zero;
  };
  // Two rows are in flight beyond the three an output needs (6 x 16 B per thread): at ~150 registers only 12 warps are resident per SM, so
  // memory-level parallelism has to come from each thread (one row ahead: 2.0 TB/s on the 512^2 grid of the s0 head).
  uint4 up[3], mid[3], dn[3], n1[3], nx[3];
  load_row(r - dil, up);
  load_row(r, mid);
  load_row(r + dil, dn);
  load_row(r + 2 * dil, n1);
  for (int y = r; y < H; y += dil) {
    load_row(y + 3 * dil, nx);  // two rows ahead of the one this output needs: the latency hides behind two iterations of 72 FMAs
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = sh[t];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      float v[8];
      unpack8(up[kx], fp16, v);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaf(v[t], w[kx][t], acc[t]);
      unpack8(mid[kx], fp16, v);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaf(v[t], w[3 + kx][t], acc[t]);
      unpack8(dn[kx], fp16, v);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaf(v[t], w[6 + kx][t], acc[t]);
    }
    uint4 pk;
    pk.x = pack2_16(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fp16); pk.y = pack2_16(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f), fp16);
    pk.z = pack2_16(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fp16); pk.w = pack2_16(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f), fp16);
    *reinterpret_cast<uint4*>(dst + (size_t(b) * H * W + size_t(y) * W + x) * C + c) = pk;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) { up[kx] = mid[kx]; mid[kx] = dn[kx]; dn[kx] = n1[kx]; n1[kx] = nx[kx]; }
  }
}

const char* depthwise3x3_nhwc16(const void* src, int B, int H, int W, int C, int dil, const float* w9, const float* shift, void* dst, int fp16,
                                cudaStream_t st) {
  if (C % 8) return "depthwise3x3: C must be a multiple of 8";
  if (dil < 1) return "depthwise3x3: dilation must be >= 1";
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(w9) | reinterpret_cast<uintptr_t>(shift)) & 15)
    return "depthwise3x3: pointers must be 16B aligned";
  const int classes = dil < H ? dil : H;  // residue classes of y that contain at least one row
  const long total = long(B) * classes * W * (C / 8);
  depthwise3x3_nhwc_kernel<<<unsigned((total + 127) / 128), 128, 0, st>>>(reinterpret_cast<const uint16_t*>(src), H, W, C, dil, w9, shift,
                                                                         reinterpret_cast<uint16_t*>(dst), total, classes, fp16);
  return cudaGetLastError() == cudaSuccess ? nullptr : "depthwise3x3 launch failed";
}

// ------------------------------------------------------------------ eval-mode BatchNorm folding (pack time)
// scale[n] = gamma / sqrt(var + eps), shift[n] = beta - mean * scale (+ conv_bias * scale)  ->  out[0..N) = scale, out[N..2N) = shift
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, const float* __restrict__ conv_bias, float eps, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float s = gamma[n] * rsqrtf(var[n] + eps);
  out[n] = s;
  out[N + n] = beta[n] - mean[n] * s + (conv_bias ? conv_bias[n] * s : 0.f);
}
const char* bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* conv_bias, float eps, int N,
                    float* out, cudaStream_t st) {
  bn_fold_kernel<<<(N + 127) / 128, 128, 0, st>>>(gamma, beta, mean, var, conv_bias, eps, N, out);
  return cudaGetLastError() == cudaSuccess ? nullptr : "bn_fold launch failed";
}

// depthwise weight [C,1,3,3] fp32 * scale[c] -> [9][C] fp32
__global__ void pack_depthwise_kernel(const float* __restrict__ w, const float* __restrict__ scale, int C, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C) return;
  const int t = i / C, c = i - t * C;
  out[i] = w[size_t(c) * 9 + t] * scale[c];
}
const char* pack_depthwise(const float* w, const float* scale, int C, float* out, cudaStream_t st) {
  pack_depthwise_kernel<<<(9 * C + 255) / 256, 256, 0, st>>>(w, scale, C, out);
  return cudaGetLastError() == cudaSuccess ? nullptr : "pack_depthwise launch failed";
}

}  // namespace madm

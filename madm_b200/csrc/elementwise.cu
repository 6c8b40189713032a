// Small HBM-bound data-movement kernels around the GEMM path: input normalisation + first-layer im2col,
// q-sample (DDPM add_noise with the shared noise) fused with the UNet conv_in im2col, timestep sinusoid,
// space-to-depth for the stride-2 convs, nearest-2x upsample, and layout conversions.
#include "cvt.cuh"
#include "kernels.h"
#include "launch.cuh"

namespace madm {


// ------------------------------------------------------------------ image -> normalised 3x3 im2col rows (K padded to 64)
// One thread per output pixel writes its 128-byte row: k = (ky*3+kx)*3 + c for k < 27, zeros after.
__global__ void image_im2col_kernel(const float* __restrict__ img, int B, int H, int W, int fp16, uint16_t* __restrict__ out,
                                    int* __restrict__ range_flag, int normalised) {
  const long idx = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = long(B) * H * W;
  if (idx >= total) return;
  const int x = int(idx % W);
  const int y = int((idx / W) % H);
  const int b = int(idx / (long(W) * H));
  const float* base = img + size_t(b) * 3 * H * W;
  float v[28];
  bool bad = false;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t = 0.f;
        if (in) {
          t = __ldg(base + (size_t(c) * H + yy) * W + xx);
          if (!normalised) t = (t - 0.5f) / 0.5f;  // input_range '-1+1' (ldm_diffusers.py:145-146); vae_encoder() callers pass [-1,1] already
          if (ky == 1 && kx == 1 && !(t >= -1.0f && t <= 1.0f)) bad = true;
        }
        v[(ky * 3 + kx) * 3 + c] = t;
      }
    }
  v[27] = 0.f;
  if (bad && range_flag) atomicOr(range_flag, 1);
  uint4* o = reinterpret_cast<uint4*>(out + size_t(idx) * 64);
#pragma unroll
  for (int j = 0; j < 28; j += 8) {
    uint4 pk;
    const uint2 a = pack4_16(v[j], v[j + 1], v[j + 2], v[j + 3], fp16);
    uint2 c2 = make_uint2(0u, 0u);
    if (j + 4 < 28) c2 = pack4_16(v[j + 4], v[j + 5], v[j + 6], v[j + 7], fp16);
    pk.x = a.x; pk.y = a.y; pk.z = c2.x; pk.w = c2.y;
    o[j / 8] = pk;
  }
#pragma unroll
  for (int j = 4; j < 8; ++j) o[j] = make_uint4(0u, 0u, 0u, 0u);
}

const char* image_im2col(const float* img, int B, int H, int W, void* out, int* range_flag, int fp16, cudaStream_t st, int normalised) {
  const long total = long(B) * H * W;
  image_im2col_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(img, B, H, W, fp16, reinterpret_cast<uint16_t*>(out), range_flag, normalised);
  return cudaGetLastError() == cudaSuccess ? nullptr : "image_im2col launch failed";
}

// ------------------------------------------------------------------ q-sample + conv_in im2col (4 channels, K = 36 -> 64)
__global__ void qsample_kernel(const float* __restrict__ lat, const float* __restrict__ noise, const int64_t* __restrict__ t,
                               const float* __restrict__ ac, int B, int HW, float* __restrict__ noisy /*[B*HW,4]*/,
                               float* __restrict__ noisy_nchw) {
  const long idx = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= long(B) * HW) return;
  const int p = int(idx % HW);
  const int b = int(idx / HW);
  const float a = ac[t[b]];
  const float sa = sqrtf(a), sb = sqrtf(1.0f - a);
  const float4 l = *reinterpret_cast<const float4*>(lat + idx * 4);
  float4 r;
  r.x = sa * l.x + sb * noise[0 * HW + p];
  r.y = sa * l.y + sb * noise[1 * HW + p];
  r.z = sa * l.z + sb * noise[2 * HW + p];
  r.w = sa * l.w + sb * noise[3 * HW + p];
  *reinterpret_cast<float4*>(noisy + idx * 4) = r;
  if (noisy_nchw) {
    float* o = noisy_nchw + size_t(b) * 4 * HW + p;
    o[0] = r.x; o[HW] = r.y; o[2 * HW] = r.z; o[3 * HW] = r.w;
  }
}

__global__ void latent_im2col_kernel(const float* __restrict__ noisy /*[B,H,W,4]*/, int B, int H, int W, int fp16,
                                     uint16_t* __restrict__ out) {
  const long idx = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= long(B) * H * W) return;
  const int x = int(idx % W);
  const int y = int((idx / W) % H);
  const int b = int(idx / (long(W) * H));
  uint2* o = reinterpret_cast<uint2*>(out + size_t(idx) * 64);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = *reinterpret_cast<const float4*>(noisy + ((size_t(b) * H + yy) * W + xx) * 4);
      o[ky * 3 + kx] = pack4_16(v.x, v.y, v.z, v.w, fp16);
    }
#pragma unroll
  for (int j = 9; j < 16; ++j) o[j] = make_uint2(0u, 0u);
}

const char* qsample(const float* lat, const float* noise_nchw, const int64_t* t, const float* alphas_cumprod, int B, int HW,
                    float* noisy_nhwc, float* noisy_nchw_or_null, cudaStream_t st) {
  const long total = long(B) * HW;
  qsample_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(lat, noise_nchw, t, alphas_cumprod, B, HW, noisy_nhwc, noisy_nchw_or_null);
  return cudaGetLastError() == cudaSuccess ? nullptr : "qsample launch failed";
}

const char* latent_im2col(const float* noisy_nhwc, int B, int H, int W, void* out_bf16, int fp16, cudaStream_t st) {
  const long total = long(B) * H * W;
  latent_im2col_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(noisy_nhwc, B, H, W, fp16, reinterpret_cast<uint16_t*>(out_bf16));
  return cudaGetLastError() == cudaSuccess ? nullptr : "latent_im2col launch failed";
}

// ------------------------------------------------------------------ timestep sinusoid
__global__ void timestep_sinusoid_kernel(const int64_t* __restrict__ t, int B, int fp16, uint16_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 160) return;
  const int b = i / 160, k = i % 160;
  const float freq = expf(-9.210340371976184f * float(k) / 160.0f);  // ln(10000)
  const float ang = float(t[b]) * freq;
  out[size_t(b) * 320 + k] = cvt_16(cosf(ang), fp16);
  out[size_t(b) * 320 + 160 + k] = cvt_16(sinf(ang), fp16);
}

const char* timestep_sinusoid(const int64_t* t, int B, void* out_bf16, int fp16, cudaStream_t st) {
  timestep_sinusoid_kernel<<<(B * 160 + 127) / 128, 128, 0, st>>>(t, B, fp16, reinterpret_cast<uint16_t*>(out_bf16));
  return cudaGetLastError() == cudaSuccess ? nullptr : "timestep_sinusoid launch failed";
}

// ------------------------------------------------------------------ fp32 (+add) (+act) -> bf16 / fp32
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, const float* __restrict__ add, long n, int act, int fp16,
                                   uint16_t* __restrict__ y, float* __restrict__ yf) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (add) v += add[i];
  if (yf) yf[i] = v;
  if (act == ACT_SILU) v = v / (1.0f + __expf(-v));
  if (y) y[i] = cvt_16(v, fp16);
}
// four elements per thread (n % 4 == 0, 16-byte aligned pointers)
__global__ void f32_to_bf16_v4_kernel(const float4* __restrict__ x, const float4* __restrict__ add, long n4, int act, int fp16,
                                      uint2* __restrict__ y, float4* __restrict__ yf) {
  pdl_trigger();
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = x[i];
  if (add) { const float4 a = add[i]; v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
  if (yf) yf[i] = v;
  if (act == ACT_SILU) {
    v.x = v.x / (1.0f + __expf(-v.x)); v.y = v.y / (1.0f + __expf(-v.y)); v.z = v.z / (1.0f + __expf(-v.z)); v.w = v.w / (1.0f + __expf(-v.w));
  }
  if (y) y[i] = pack4_16(v.x, v.y, v.z, v.w, fp16);
}

const char* f32_to_bf16(const float* x, const float* add, long n, int act, void* y, float* yf, int fp16, cudaStream_t st) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(add) | reinterpret_cast<uintptr_t>(yf) | (reinterpret_cast<uintptr_t>(y) << 1);
  if (n % 4 == 0 && (al & 15) == 0)
    launch_k(f32_to_bf16_v4_kernel, dim3(unsigned((n / 4 + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(add),
             n / 4, act, fp16, reinterpret_cast<uint2*>(y), reinterpret_cast<float4*>(yf));
  else
    launch_k(f32_to_bf16_kernel, dim3(unsigned((n + 255) / 256)), dim3(256), 0, st, x, add, n, act, fp16, reinterpret_cast<uint16_t*>(y), yf);
  return cudaGetLastError() == cudaSuccess ? nullptr : "f32_to_bf16 launch failed";
}

// ------------------------------------------------------------------ space-to-depth (stride-2 conv input), fp32 -> bf16
// out[phase][b][y/2][x/2][c], phase = (y&1)*2 + (x&1)
__global__ void space_to_depth_kernel(const float* __restrict__ x, int B, int H, int W, int C, int fp16, uint16_t* __restrict__ out) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;  // one float4 per thread
  const int Q = C >> 2;
  const long total = long(B) * H * W * Q;
  if (i >= total) return;
  const int c = int(i % Q) * 4;
  const long pix = i / Q;
  const int xx = int(pix % W);
  const int yy = int((pix / W) % H);
  const int b = int(pix / (long(W) * H));
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  const int ph = (yy & 1) * 2 + (xx & 1);
  const int H2 = H >> 1, W2 = W >> 1;
  const size_t o = (((size_t(ph) * B + b) * H2 + (yy >> 1)) * W2 + (xx >> 1)) * C + c;
  *reinterpret_cast<uint2*>(out + o) = pack4_16(v.x, v.y, v.z, v.w, fp16);
}

const char* space_to_depth(const float* x, int B, int H, int W, int C, void* out, int fp16, cudaStream_t st) {
  if ((H | W) & 1 || C % 4) return "space_to_depth: H, W must be even and C % 4 == 0";
  const long total = long(B) * H * W * (C / 4);
  launch_k(space_to_depth_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), 0, st, x, B, H, W, C, fp16, reinterpret_cast<uint16_t*>(out));
  return cudaGetLastError() == cudaSuccess ? nullptr : "space_to_depth launch failed";
}

// ------------------------------------------------------------------ nearest 2x upsample, fp32 -> bf16
__global__ void upsample2x_kernel(const float* __restrict__ x, int B, int H, int W, int C, int fp16, uint16_t* __restrict__ out) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;  // one float4 of the INPUT per thread -> 4 outputs
  const int Q = C >> 2;
  const long total = long(B) * H * W * Q;
  if (i >= total) return;
  const int c = int(i % Q) * 4;
  const long pix = i / Q;
  const int xx = int(pix % W);
  const int yy = int((pix / W) % H);
  const int b = int(pix / (long(W) * H));
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  const uint2 pk = pack4_16(v.x, v.y, v.z, v.w, fp16);
  const int W2 = W * 2;
  const size_t row0 = ((size_t(b) * H * 2 + yy * 2) * W2 + xx * 2) * C + c;
  *reinterpret_cast<uint2*>(out + row0) = pk;
  *reinterpret_cast<uint2*>(out + row0 + C) = pk;
  *reinterpret_cast<uint2*>(out + row0 + size_t(W2) * C) = pk;
  *reinterpret_cast<uint2*>(out + row0 + size_t(W2) * C + C) = pk;
}

const char* upsample_nearest2x(const float* x, int B, int H, int W, int C, void* out, int fp16, cudaStream_t st) {
  if (C % 4) return "upsample_nearest2x: C % 4 != 0";
  const long total = long(B) * H * W * (C / 4);
  launch_k(upsample2x_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), 0, st, x, B, H, W, C, fp16, reinterpret_cast<uint16_t*>(out));
  return cudaGetLastError() == cudaSuccess ? nullptr : "upsample_nearest2x launch failed";
}

// ------------------------------------------------------------------ image preprocessing: bilinear resize + zero pad (fp32 NCHW)
// T.Resize(size, BILINEAR) on a tensor with torchvision 0.16.1 (the reference's pinned version: antialias defaults to "warn" = off) is
// F.interpolate(mode='bilinear', align_corners=False) = ATen upsample_bilinear2d: src = max((dst + 0.5) * in/out - 0.5, 0).  The output
// may be larger than the resized image: the rest is the zero padding of ImageList.from_tensors(size_divisibility).
__global__ void resize_bilinear_nchw_kernel(const float* __restrict__ src, int planes, int Hs, int Ws, int Hr, int Wr, int Hd, int Wd, float sy,
                                            float sx, float* __restrict__ dst) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = long(planes) * Hd * Wd;
  if (i >= total) return;
  const int x = int(i % Wd), y = int((i / Wd) % Hd);
  const long pl = i / (long(Wd) * Hd);
  float v = 0.f;
  if (y < Hr && x < Wr) {
    const float fy = fmaxf((y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((x + 0.5f) * sx - 0.5f, 0.f);
    const int y0 = min(int(fy), Hs - 1), x0 = min(int(fx), Ws - 1);
    const int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
    const float ly = fy - float(y0), lx = fx - float(x0), hy = 1.f - ly, hx = 1.f - lx;
    const float* p = src + pl * long(Hs) * Ws;
    // same association as ATen: hy * (hx * a + lx * b) + ly * (hx * c + lx * d)
    v = hy * (hx * __ldg(p + size_t(y0) * Ws + x0) + lx * __ldg(p + size_t(y0) * Ws + x1)) +
        ly * (hx * __ldg(p + size_t(y1) * Ws + x0) + lx * __ldg(p + size_t(y1) * Ws + x1));
  }
  dst[i] = v;
}

const char* resize_bilinear_nchw(const float* src, int planes, int Hs, int Ws, int Hr, int Wr, int Hd, int Wd, float* dst, cudaStream_t st) {
  if (Hs < 1 || Ws < 1 || Hr < 1 || Wr < 1 || Hd < Hr || Wd < Wr) return "resize_bilinear: bad geometry";
  const long total = long(planes) * Hd * Wd;
  resize_bilinear_nchw_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(src, planes, Hs, Ws, Hr, Wr, Hd, Wd, float(Hs) / float(Hr),
                                                                            float(Ws) / float(Wr), dst);
  return cudaGetLastError() == cudaSuccess ? nullptr : "resize_bilinear launch failed";
}

// ------------------------------------------------------------------ nearest 2x upsample of a 16-bit tensor (8 channels per thread)
__global__ void upsample2x_16_kernel(const uint4* __restrict__ x, int B, int H, int W, int C8, uint4* __restrict__ out) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = long(B) * H * W * C8;
  if (i >= total) return;
  const int c = int(i % C8);
  const long pix = i / C8;
  const int xx = int(pix % W);
  const int yy = int((pix / W) % H);
  const int b = int(pix / (long(W) * H));
  const uint4 v = __ldg(x + i);
  const int W2 = W * 2;
  const size_t row0 = ((size_t(b) * H * 2 + yy * 2) * W2 + xx * 2) * C8 + c;
  out[row0] = v;
  out[row0 + C8] = v;
  out[row0 + size_t(W2) * C8] = v;
  out[row0 + size_t(W2) * C8 + C8] = v;
}

const char* upsample_nearest2x_16(const void* x16, int B, int H, int W, int C, void* out16, cudaStream_t st) {
  if (C % 8) return "upsample_nearest2x_16: C % 8 != 0";
  const long total = long(B) * H * W * (C / 8);
  launch_k(upsample2x_16_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const uint4*>(x16), B, H, W, C / 8,
           reinterpret_cast<uint4*>(out16));
  return cudaGetLastError() == cudaSuccess ? nullptr : "upsample_nearest2x_16 launch failed";
}

// ------------------------------------------------------------------ VAE decoder entry / exit
__global__ void post_quant_conv_kernel(const float4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float inv_scale,
                                       long M, float4* __restrict__ z) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float4 v = __ldg(x + i);
  v.x *= inv_scale; v.y *= inv_scale; v.z *= inv_scale; v.w *= inv_scale;
  float o[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n] = __ldg(bias + n) + __ldg(w + n * 4) * v.x + __ldg(w + n * 4 + 1) * v.y + __ldg(w + n * 4 + 2) * v.z + __ldg(w + n * 4 + 3) * v.w;
  z[i] = make_float4(o[0], o[1], o[2], o[3]);
}

const char* post_quant_conv(const float* sample, const float* w, const float* bias, float inv_scale, long M, float* z, cudaStream_t st) {
  post_quant_conv_kernel<<<unsigned((M + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(sample), w, bias, inv_scale, M,
                                                                    reinterpret_cast<float4*>(z));
  return cudaGetLastError() == cudaSuccess ? nullptr : "post_quant_conv launch failed";
}

// 8 threads per pixel: each writes one 16-byte piece of the pixel's 128-byte operand row (piece 0 carries the 3 channels)
__global__ void decoder_image_pack_kernel(const float4* __restrict__ img4, int HW, long M, int fp16, uint4* __restrict__ rows, float* __restrict__ clipped,
                                          float* __restrict__ raw) {
  const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long pix = t >> 3;
  const int piece = int(t & 7);
  if (pix >= M) return;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (piece == 0) {
    const float4 v = __ldg(img4 + pix);
    const uint2 pk = pack4_16(v.x, v.y, v.z, 0.f, fp16);
    o.x = pk.x; o.y = pk.y;
    const long b = pix / HW, p = pix % HW;
    if (clipped) {
      float* dst = clipped + size_t(b) * 3 * HW + p;
      dst[0] = fminf(fmaxf(v.x, -1.f), 1.f);
      dst[HW] = fminf(fmaxf(v.y, -1.f), 1.f);
      dst[2 * size_t(HW)] = fminf(fmaxf(v.z, -1.f), 1.f);
    }
    if (raw) {
      float* dst = raw + size_t(b) * 3 * HW + p;
      dst[0] = v.x; dst[HW] = v.y; dst[2 * size_t(HW)] = v.z;
    }
  }
  if (rows) rows[t] = o;
}

const char* decoder_image_pack(const float* img4, int B, int HW, void* rows16, float* clipped, float* raw, int fp16, cudaStream_t st) {
  const long M = long(B) * HW;
  decoder_image_pack_kernel<<<unsigned((M * 8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(img4), HW, M, fp16,
                                                                           reinterpret_cast<uint4*>(rows16), clipped, raw);
  return cudaGetLastError() == cudaSuccess ? nullptr : "decoder_image_pack launch failed";
}

// ------------------------------------------------------------------ split-K reduction + fused epilogue
// out = act( sum_s partial[s] (fixed order) + bias + rowbias[img] + residual ) -> fp32 and/or 16-bit; one float4 per thread
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, long split_stride, int M, int N, const float* __restrict__ bias,
                                     const float* __restrict__ rowbias, int rows_per_img, int ld_rowbias, const float* residual, int ldr,
                                     float* out32, int ldo32, uint16_t* out16, int ldo16, int act, int fp16, int res16) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const int Q = N >> 2;
  if (i >= long(M) * Q) return;
  const int m = int(i / Q), n = int(i % Q) * 4;
  float4 v = *reinterpret_cast<const float4*>(part + size_t(m) * N + n);
  for (int s = 1; s < splits; ++s) {
    const float4 t = *reinterpret_cast<const float4*>(part + size_t(s) * split_stride + size_t(m) * N + n);
    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  }
  if (bias) { const float4 t = __ldg(reinterpret_cast<const float4*>(bias + n)); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
  if (rowbias) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(rowbias + size_t(m / rows_per_img) * ld_rowbias + n));
    v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
  }
  if (residual && res16) {  // 16-bit residual stream (may alias out16: read before the store below)
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(residual) + size_t(m) * ldr + n);
    float2 lo, hi;
    if (fp16) { lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x)); hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y)); }
    else { lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)); hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y)); }
    v.x += lo.x; v.y += lo.y; v.z += hi.x; v.w += hi.y;
  } else if (residual) { const float4 t = *reinterpret_cast<const float4*>(residual + size_t(m) * ldr + n); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
  if (act == ACT_SILU) { v.x = v.x / (1.f + __expf(-v.x)); v.y = v.y / (1.f + __expf(-v.y)); v.z = v.z / (1.f + __expf(-v.z)); v.w = v.w / (1.f + __expf(-v.w)); }
  else if (act == ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  if (out32) *reinterpret_cast<float4*>(out32 + size_t(m) * ldo32 + n) = v;
  if (out16) *reinterpret_cast<uint2*>(out16 + size_t(m) * ldo16 + n) = pack4_16(v.x, v.y, v.z, v.w, fp16);
}

const char* splitk_reduce(const float* part, int splits, long split_stride, int M, int N, const float* bias, const float* rowbias,
                          int rows_per_img, int ld_rowbias, const float* residual, int ldr, float* out32, int ldo32, void* out16, int ldo16,
                          int act, int fp16, cudaStream_t st, int res16) {
  if (N % 4) return "splitk_reduce: N % 4 != 0";
  const long total = long(M) * (N / 4);
  launch_k(splitk_reduce_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), 0, st, part, splits, split_stride, M, N, bias, rowbias,
           rows_per_img > 0 ? rows_per_img : 1, ld_rowbias ? ld_rowbias : N, residual, ldr, out32, ldo32, reinterpret_cast<uint16_t*>(out16), ldo16,
           act, fp16, res16);
  return cudaGetLastError() == cudaSuccess ? nullptr : "splitk_reduce launch failed";
}

// ------------------------------------------------------------------ NHWC fp32 -> NCHW fp32 (32x32 tiles through smem)
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int HW, int C, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (p < HW && c < C) ? x[(size_t(b) * HW + p) * C + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (p < HW && c < C) out[(size_t(b) * C + c) * HW + p] = tile[tx][r];
  }
}

// the same from rows of pitch ld >= C (padded GEMM output, e.g. 19 class logits in 32 columns)
__global__ void nhwc_to_nchw_strided_kernel(const float* __restrict__ x, int HW, int C, int ld, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (p < HW && c < C) ? x[(size_t(b) * HW + p) * ld + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (p < HW && c < C) out[(size_t(b) * C + c) * HW + p] = tile[tx][r];
  }
}
const char* nhwc_to_nchw_strided(const float* x, int B, int HW, int C, int ld, float* out, cudaStream_t st) {
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  nhwc_to_nchw_strided_kernel<<<grid, dim3(32, 8), 0, st>>>(x, HW, C, ld, out);
  return cudaGetLastError() == cudaSuccess ? nullptr : "nhwc_to_nchw_strided launch failed";
}

const char* nhwc_to_nchw(const float* x, int B, int HW, int C, float* out, cudaStream_t st) {
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, st>>>(x, HW, C, out);
  return cudaGetLastError() == cudaSuccess ? nullptr : "nhwc_to_nchw launch failed";
}

}  // namespace madm

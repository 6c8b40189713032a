// Optimizer side of MADM's training step (the part that follows the backward pass) and the image side of the
// DACS mixing (row f-4), as HBM-bound multi-tensor kernels behind the C ABI:
//   ema_update_kernel    CMDISE._update_ema (reference modeling/meta_arch/cmdise.py:337-349): ema = a * ema + (1 - a) * param over all
//                        EMA-tracked tensors (feature projections, clip_project_others, head) in one launch per 48 tensors
//   grad_sumsq_*         torch.nn.utils.clip_grad_norm_ (engine/train_loop.py:123-124, :201-210): global L2 norm of the gradients,
//                        two fixed-order stages (no atomics) -> a device scalar; nothing returns to the host
//   adamw_kernel         torch.optim.AdamW step (config_files/common/optim.py:9-18) with the clip coefficient read from that device
//                        scalar: p *= 1 - lr*wd; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g; p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps)
//   image_mix_kernel     dacs_transforms.one_mix for images: mask * a + (1 - mask) * b, mask broadcast over channels
//   gaussian_blur_*      kornia.filters.GaussianBlur2d as dacs_transforms.gaussian_blur calls it (separable, border 'reflect')
// Grid-stride coalesced accesses (tensor = blockIdx.y); every tensor is read and written once.
#include "kernels.h"

namespace madm {

static constexpr int kMtMax = 48;  // tensors per launch: 4 pointer tables + sizes stay under the 4 KB kernel-parameter limit

struct MtTable {
  void* p[4][kMtMax];
  long numel[kMtMax];
};

__global__ void __launch_bounds__(256) ema_update_kernel(MtTable t, float wa, float wb) {
  const int k = blockIdx.y;
  float* __restrict__ e = static_cast<float*>(t.p[0][k]);
  const float* __restrict__ p = static_cast<const float*>(t.p[1][k]);
  const long n = t.numel[k];
  for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x)
    e[i] = __fadd_rn(__fmul_rn(wa, e[i]), __fmul_rn(wb, p[i]));  // two rounded products + add, exactly like the reference's tensor ops
}

__global__ void __launch_bounds__(256) grad_sumsq_kernel(MtTable t, int first, float* __restrict__ partial /*[ntensors][32]*/) {
  const int k = blockIdx.y;
  const float* __restrict__ g = static_cast<const float*>(t.p[0][k]);
  const long n = t.numel[k];
  float acc = 0.f;
  for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) acc = fmaf(g[i], g[i], acc);
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {  // fixed-order tree
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[size_t(first + k) * 32 + blockIdx.x] = red[0];
}

__global__ void grad_norm_finish_kernel(const float* __restrict__ partial, int n, float* __restrict__ out_norm) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += double(partial[i]);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_norm = float(sqrt(red[0]));
}

struct AdamArgs {
  float decay;       // 1 - lr * weight_decay
  float w1;          // 1 - beta1
  float beta2, w2;   // beta2, 1 - beta2
  float step_size;   // lr / (1 - beta1^t)
  float bc2_sqrt;    // sqrt(1 - beta2^t)
  float eps;
  float max_norm;    // <= 0: no clipping
};

__global__ void __launch_bounds__(256) adamw_kernel(MtTable t, AdamArgs a, const float* __restrict__ grad_norm) {
  const int k = blockIdx.y;
  float* __restrict__ p = static_cast<float*>(t.p[0][k]);
  const float* __restrict__ g = static_cast<const float*>(t.p[1][k]);
  float* __restrict__ m = static_cast<float*>(t.p[2][k]);
  float* __restrict__ v = static_cast<float*>(t.p[3][k]);
  const long n = t.numel[k];
  float coef = 1.0f;
  if (grad_norm) {
    const float gn = *grad_norm;
    // a non-finite gradient norm skips the whole step, parameters and moments untouched: what GradScaler.step does under the reference's
    // AMP trainer (engine/train_loop.py:277-302) when unscale_ finds an inf / NaN; fminf(NaN, 1) would otherwise apply it unclipped
    if (!isfinite(gn)) return;
    if (a.max_norm > 0.f) coef = fminf(a.max_norm / (gn + 1e-6f), 1.0f);  // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max=1)
  }
  for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) {
    const float gi = __fmul_rn(g[i], coef);
    float pi = __fmul_rn(p[i], a.decay);
    const float mi = fmaf(a.w1, gi - m[i], m[i]);                       // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(a.w2 * gi, gi, __fmul_rn(v[i], a.beta2));     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), a.bc2_sqrt), a.eps);
    pi = fmaf(-a.step_size, __fdiv_rn(mi, denom), pi);                  // param.addcdiv_(exp_avg, denom, value = -step_size)
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

static int grid_x_for(long max_numel) {
  long g = (max_numel + 256 * 8 - 1) / (256 * 8);
  if (g < 1) g = 1;
  if (g > 32) g = 32;
  return int(g);
}

template <typename F>
static const char* for_each_group(int ntab, void* const* const* tabs, const long* numel, int n, F launch) {
  for (int first = 0; first < n; first += kMtMax) {
    const int cnt = n - first < kMtMax ? n - first : kMtMax;
    MtTable t;
    long mx = 1;
    for (int k = 0; k < cnt; ++k) {
      for (int a = 0; a < ntab; ++a) {
        t.p[a][k] = tabs[a][first + k];
        if (!t.p[a][k]) return "multi-tensor op: null tensor pointer";
      }
      t.numel[k] = numel[first + k];
      if (t.numel[k] < 0) return "multi-tensor op: negative size";
      if (t.numel[k] > mx) mx = t.numel[k];
    }
    launch(t, first, cnt, mx);
    if (cudaGetLastError() != cudaSuccess) return "multi-tensor kernel launch failed";
  }
  return nullptr;
}

const char* ema_update(float* const* ema, const float* const* param, const long* numel, int n, float wa, float wb, cudaStream_t st) {
  void* const* tabs[2] = {reinterpret_cast<void* const*>(ema), reinterpret_cast<void* const*>(const_cast<float* const*>(param))};
  return for_each_group(2, tabs, numel, n, [&](const MtTable& t, int, int cnt, long mx) {
    ema_update_kernel<<<dim3(grid_x_for(mx) * 4, cnt), 256, 0, st>>>(t, wa, wb);
  });
}

int grad_norm_scratch_floats(int n) { return n * 32; }

const char* grad_norm(const float* const* grad, const long* numel, int n, float* partial, float* out_norm, cudaStream_t st) {
  if (n < 1) return "grad_norm: no tensors";
  if (cudaMemsetAsync(partial, 0, size_t(n) * 32 * sizeof(float), st) != cudaSuccess) return "grad_norm: memset failed";
  void* const* tabs[1] = {reinterpret_cast<void* const*>(const_cast<float* const*>(grad))};
  const char* e = for_each_group(1, tabs, numel, n, [&](const MtTable& t, int first, int cnt, long mx) {
    grad_sumsq_kernel<<<dim3(grid_x_for(mx), cnt), 256, 0, st>>>(t, first, partial);
  });
  if (e) return e;
  grad_norm_finish_kernel<<<1, 256, 0, st>>>(partial, n * 32, out_norm);
  return cudaGetLastError() == cudaSuccess ? nullptr : "grad_norm_finish launch failed";
}

const char* adamw_step(float* const* param, const float* const* grad, float* const* exp_avg, float* const* exp_avg_sq, const long* numel, int n,
                       double lr, double beta1, double beta2, double eps, double weight_decay, int step, const float* grad_norm_dev, float max_norm,
                       cudaStream_t st) {
  if (step < 1) return "adamw_step: step counts from 1";
  // scalar prologue in double like torch.optim.adamw._single_tensor_adamw (python floats), cast to fp32 where the tensor ops take them
  const double bc1 = 1.0 - pow(beta1, double(step));
  const double bc2 = 1.0 - pow(beta2, double(step));
  AdamArgs a;
  a.decay = float(1.0 - lr * weight_decay);
  a.w1 = float(1.0 - beta1);
  a.beta2 = float(beta2);
  a.w2 = float(1.0 - beta2);
  a.step_size = float(lr / bc1);
  a.bc2_sqrt = float(sqrt(bc2));
  a.eps = float(eps);
  a.max_norm = max_norm;
  void* const* tabs[4] = {reinterpret_cast<void* const*>(param), reinterpret_cast<void* const*>(const_cast<float* const*>(grad)),
                          reinterpret_cast<void* const*>(exp_avg), reinterpret_cast<void* const*>(exp_avg_sq)};
  return for_each_group(4, tabs, numel, n, [&](const MtTable& t, int, int cnt, long mx) {
    adamw_kernel<<<dim3(grid_x_for(mx) * 4, cnt), 256, 0, st>>>(t, a, grad_norm_dev);
  });
}

// ------------------------------------------------------------------ DACS image mixing: mask [HW] (int64, 0/1) over [C,HW] images
__global__ void image_mix_kernel(const int64_t* __restrict__ mask, const float* __restrict__ a, const float* __restrict__ b, int C, long HW,
                                 float* __restrict__ out) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const float mk = float(mask[i]);
  for (int c = 0; c < C; ++c) {
    const size_t o = size_t(c) * HW + i;
    out[o] = __fadd_rn(__fmul_rn(mk, a[o]), __fmul_rn(1.0f - mk, b[o]));  // stackedMask0 * data[0] + (1 - stackedMask0) * data[1]
  }
}

const char* image_mix(const int64_t* mask, const float* a, const float* b, int C, long HW, float* out, cudaStream_t st) {
  image_mix_kernel<<<unsigned((HW + 255) / 256), 256, 0, st>>>(mask, a, b, C, HW, out);
  return cudaGetLastError() == cudaSuccess ? nullptr : "image_mix launch failed";
}

// ------------------------------------------------------------------ colour jitter (kornia.augmentation.ColorJitter.apply_transform)
// dacs_transforms.color_jitter (reference utils/dacs_transforms.py:41-59): the four adjustments of kornia's classic ColorJitter (0.6.x - 0.7.0:
// additive brightness, multiplicative contrast, saturation and hue through HSV), applied per image in the sampled order with the sampled
// factors, between the reference's denorm_ / renorm_ (x * std + mean ... (x - mean) / std).  One thread = one pixel; all arithmetic in fp32.
// torch semantics kept: clamp(0, 1) after brightness / contrast / saturation scaling, torch.fmod for the hue shift, torch.remainder ("%")
// in the RGB <-> HSV conversions (kornia.color.rgb_to_hsv with eps 1e-8, hue in [0, 2 pi]).
__device__ __forceinline__ float torch_remainder(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.f && ((b < 0.f) != (m < 0.f))) m += b;
  return m;
}
__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
  const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b));
  const int arg = (r >= g && r >= b) ? 0 : (g >= b ? 1 : 2);  // first maximum, like torch.max
  v = mx;
  float d = mx - mn;
  s = d / (mx + 1e-8f);
  if (d == 0.f) d = 1.f;
  const float rc = mx - r, gc = mx - g, bc = mx - b;
  const float hh = arg == 0 ? (bc - gc) : (arg == 1 ? (rc - bc) + 2.0f * d : (gc - rc) + 4.0f * d);
  h = 6.283185307179586f * torch_remainder((hh / d) / 6.0f, 1.0f);
}
__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float& r, float& g, float& b) {
  const float h6 = (h / 6.283185307179586f) * 6.0f;
  const float hi_f = torch_remainder(floorf(h6), 6.0f);
  const float f = torch_remainder(h6, 6.0f) - hi_f;
  const float p = v * (1.0f - s), q = v * (1.0f - f * s), t = v * (1.0f - (1.0f - f) * s);
  switch (int(hi_f)) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

__global__ void __launch_bounds__(256) color_jitter_kernel(const float* __restrict__ in, long HW, const int* __restrict__ order /*[B][4]*/,
                                                           const float* __restrict__ factors /*[B][4]*/, const float* __restrict__ mean,
                                                           const float* __restrict__ stdv, float* __restrict__ out) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  const int bi = blockIdx.y;
  const float* p = in + size_t(bi) * 3 * HW + i;
  float c[3] = {p[0], p[HW], p[2 * HW]};
  if (mean) {
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = __fadd_rn(__fmul_rn(c[k], stdv[k]), mean[k]);  // denorm_: img.mul_(std).add_(mean)
  }
  for (int k = 0; k < 4; ++k) {
    const int op = order[bi * 4 + k];
    const float fac = factors[bi * 4 + op];
    if (op == 0) {         // adjust_brightness(img, brightness_factor - 1): additive
      for (int j = 0; j < 3; ++j) c[j] = fminf(fmaxf(c[j] + fac, 0.f), 1.f);
    } else if (op == 1) {  // adjust_contrast(img, contrast_factor): multiplicative
      for (int j = 0; j < 3; ++j) c[j] = fminf(fmaxf(c[j] * fac, 0.f), 1.f);
    } else {
      float h, s, v;
      rgb_to_hsv(c[0], c[1], c[2], h, s, v);
      if (op == 2) s = fminf(fmaxf(s * fac, 0.f), 1.f);  // adjust_saturation
      else h = fmodf(h + fac, 6.283185307179586f);       // adjust_hue(img, hue_factor * 2 pi): torch.fmod
      hsv_to_rgb(h, s, v, c[0], c[1], c[2]);
    }
  }
  if (mean) {
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = __fdiv_rn(__fsub_rn(c[k], mean[k]), stdv[k]);  // renorm_: img.sub_(mean).div_(std)
  }
  float* o = out + size_t(bi) * 3 * HW + i;
  o[0] = c[0]; o[HW] = c[1]; o[2 * HW] = c[2];
}

const char* color_jitter(const float* in, int B, long HW, const int* order, const float* factors, const float* mean, const float* stdv, float* out,
                         cudaStream_t st) {
  if ((mean == nullptr) != (stdv == nullptr)) return "color_jitter: mean and std come together";
  color_jitter_kernel<<<dim3(unsigned((HW + 255) / 256), B), 256, 0, st>>>(in, HW, order, factors, mean, stdv, out);
  return cudaGetLastError() == cudaSuccess ? nullptr : "color_jitter launch failed";
}

// ------------------------------------------------------------------ separable Gaussian blur, reflect border (no edge repeat)
static constexpr int kBlurMaxK = 129;
struct BlurTaps { float w[kBlurMaxK]; };

__device__ __forceinline__ int reflect_idx(int i, int n) {  // torch 'reflect' padding: -1 -> 1, n -> n-2
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

template <bool ALONG_X>
__global__ void __launch_bounds__(256) gaussian_blur_pass_kernel(const float* __restrict__ src, int planes, int H, int W, int K, BlurTaps taps,
                                                                 float* __restrict__ dst) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = long(planes) * H * W;
  if (i >= total) return;
  const int x = int(i % W), y = int((i / W) % H);
  const float* p = src + (i / (long(W) * H)) * long(H) * W;
  const int r = K / 2;
  float acc = 0.f;
  if (ALONG_X) {
    const float* row = p + size_t(y) * W;
    for (int k = 0; k < K; ++k) acc = fmaf(taps.w[k], __ldg(row + reflect_idx(x + k - r, W)), acc);
  } else {
    for (int k = 0; k < K; ++k) acc = fmaf(taps.w[k], __ldg(p + size_t(reflect_idx(y + k - r, H)) * W + x), acc);
  }
  dst[i] = acc;
}

// kornia.filters.get_gaussian_kernel1d: x = arange(K) - K // 2 (odd K), g = exp(-x^2 / (2 sigma^2)), normalised to sum 1
static const char* gaussian_taps(int K, float sigma, BlurTaps* t) {
  if (K < 1 || K > kBlurMaxK || (K & 1) == 0) return "gaussian_blur: kernel size must be odd and <= 129";
  if (!(sigma > 0.f)) return "gaussian_blur: sigma must be positive";
  double sum = 0.0;
  for (int k = 0; k < K; ++k) {
    const float x = float(k - K / 2);
    t->w[k] = expf(-(x * x) / (2.0f * sigma * sigma));
    sum += double(t->w[k]);
  }
  for (int k = 0; k < K; ++k) t->w[k] = float(double(t->w[k]) / sum);
  return nullptr;
}

const char* gaussian_blur(const float* src, int planes, int H, int W, int ky, int kx, float sigma_y, float sigma_x, float* tmp, float* dst,
                          cudaStream_t st) {
  if (ky / 2 >= H || kx / 2 >= W) return "gaussian_blur: reflect padding needs kernel radius < image size";
  BlurTaps tx, ty;
  if (const char* e = gaussian_taps(kx, sigma_x, &tx)) return e;
  if (const char* e = gaussian_taps(ky, sigma_y, &ty)) return e;
  const long total = long(planes) * H * W;
  const unsigned grid = unsigned((total + 255) / 256);
  gaussian_blur_pass_kernel<true><<<grid, 256, 0, st>>>(src, planes, H, W, kx, tx, tmp);
  gaussian_blur_pass_kernel<false><<<grid, 256, 0, st>>>(tmp, planes, H, W, ky, ty, dst);
  return cudaGetLastError() == cudaSuccess ? nullptr : "gaussian_blur launch failed";
}

}  // namespace madm

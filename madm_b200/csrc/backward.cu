// Backward pass of the HBM-bound layers of the LoRA training step (reference engine/train_loop.py:277-302 runs
// torch.autograd through diffusers' GroupNorm / LayerNorm / GEGLU / nearest-upsample / strided conv and detectron2's GN bottleneck).
// Activation gradients travel as 16-bit tensors of the context's operand dtype between the dgrad GEMMs and these kernels and as fp32 on
// the residual stream, exactly mirroring the forward layouts (NHWC).  No atomics anywhere: every reduction is a fixed-order tree, so
// gradients are bit-reproducible run to run.
#include "cvt.cuh"
#include "kernels.h"
#include "launch.cuh"

#include <cooperative_groups.h>

namespace madm {

namespace {

__device__ __forceinline__ float2 ld2_16(const uint16_t* p, int fp16) {
  const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
  if (fp16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ float ld1_16(const uint16_t* p, int fp16) {
  const uint16_t w = *p;
  if (fp16) return __half2float(*reinterpret_cast<const __half*>(&w));
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&w));
}
__device__ __forceinline__ void st2_16(uint16_t* p, float a, float b, int fp16) { *reinterpret_cast<uint32_t*>(p) = pack2_16(a, b, fp16); }

// derivative of act(z) with respect to z
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == ACT_SILU) {
    const float s = 1.0f / (1.0f + __expf(-z));
    return s * (1.0f + z * (1.0f - s));
  }
  if (act == ACT_RELU) return z > 0.0f ? 1.0f : 0.0f;
  return 1.0f;
}

// ------------------------------------------------------------------------------------------------ GroupNorm(32) backward
// y = act(xhat * gamma + beta), xhat = (x - mean_g) * rstd_g over the (HW x C/32) elements of (image, group).
// With g = dy * act'(z):  dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)),  dgamma_c = sum g*xhat,  dbeta_c = sum g.
// Thread layout shared by the reduce and the apply pass: a CTA owns a slab of pixels of one image; thread (lane v, pixel lane pl)
// owns the channel pairs v, v + TX, ... (KMAX of them) and walks pixels pl, pl + P, ...
constexpr int kGnKMax = 5;       // channel pairs per thread: C <= 2 * 256 * 5 = 2560 (the widest concat input of the UNet)
constexpr int kGnThreads = 512;  // CTA size limit: TX * P <= 512

struct GnBwdGeo {
  int TX;  // threads along the channel-pair axis (<= 256)
  int P;   // pixel lanes (1..4)
  int K;   // channel pairs per thread
};
__host__ __device__ inline GnBwdGeo gn_bwd_geo(int C) {
  GnBwdGeo g;
  const int cv = C / 2;
  g.K = (cv + 255) / 256;
  while (g.K < kGnKMax && cv % g.K != 0) ++g.K;  // an exact split where one exists (C = 1280: 640 pairs = 4 x 160)
  g.TX = (cv + g.K - 1) / g.K;
  g.P = kGnThreads / g.TX;
  if (g.P > 4) g.P = 4;
  if (g.P < 1) g.P = 1;
  return g;
}

struct GnChan {  // per-channel constants of one channel pair
  float mean[2], rstd[2], ga[2], be[2];
};

__device__ __forceinline__ void gn_load_chan(const float* __restrict__ stats /*[32][2] sums of this image*/, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, int c, int cpg, float inv_n, float eps, GnChan& ch) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int g = (c + t) / cpg;
    const float mean = stats[g * 2] * inv_n;
    const float var = fmaxf(stats[g * 2 + 1] * inv_n - mean * mean, 0.0f);
    ch.mean[t] = mean;
    ch.rstd[t] = rsqrtf(var + eps);
    ch.ga[t] = gamma ? gamma[c + t] : 1.0f;
    ch.be[t] = beta ? beta[c + t] : 0.0f;
  }
}

template <bool IN16>
__device__ __forceinline__ float2 gn_load_x(const void* x0, int C0, const void* x1, int C1, size_t pix /* b*HW + p */, int c, int fp16) {
  const void* src; int ld, cc;
  if (c < C0) { src = x0; ld = C0; cc = c; } else { src = x1; ld = C1; cc = c - C0; }
  if constexpr (IN16) return ld2_16(reinterpret_cast<const uint16_t*>(src) + pix * ld + cc, fp16);
  else return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(src) + pix * ld + cc);
}

// pass 1: partial[b][slab][c][2] = sum over the slab's pixels of (g, g * xhat)
// KT = channel pairs per thread (geo.K, 1..5): compile-time, so a C = 320 launch carries one pair's constants and accumulators instead of five
// (128 registers at 512 threads = one CTA per SM before)
template <bool IN16, int KT>
__global__ void __launch_bounds__(kGnThreads) gn_bwd_reduce_kernel(const void* __restrict__ x0, int C0, const void* __restrict__ x1, int C1, int HW,
                                                                   int slab_pix, const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float eps, int act,
                                                                   const uint16_t* __restrict__ dy, int fp16, float* __restrict__ partial) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const int C = C0 + C1, cv = C / 2, cpg = C / 32;
  const GnBwdGeo geo = gn_bwd_geo(C);
  const int b = blockIdx.y, slab = blockIdx.x, slabs = gridDim.x;
  const int tx = threadIdx.x % geo.TX, pl = threadIdx.x / geo.TX;
  const bool active = pl < geo.P;
  const float inv_n = 1.0f / (float(HW) * float(cpg));
  GnChan ch[KT];
  float acc[KT][4];
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
    const int v = tx + k * geo.TX;
    if (k < geo.K && v < cv && active) gn_load_chan(stats + size_t(b) * 64, gamma, beta, 2 * v, cpg, inv_n, eps, ch[k]);
  }
  const int p0 = slab * slab_pix, p1 = min(HW, p0 + slab_pix);
  if (active) {
    for (int p = p0 + pl; p < p1; p += 2 * geo.P) {  // two pixels per trip: their loads are independent and stay in flight together
      const size_t pixa = size_t(b) * HW + p;
      const bool two = p + geo.P < p1;
      const size_t pixb = two ? pixa + geo.P : pixa;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const int v = tx + k * geo.TX;
        if (k < geo.K && v < cv) {
          const float2 xa = gn_load_x<IN16>(x0, C0, x1, C1, pixa, 2 * v, fp16);
          const float2 da = ld2_16(dy + pixa * C + 2 * v, fp16);
          const float2 xb = gn_load_x<IN16>(x0, C0, x1, C1, pixb, 2 * v, fp16);
          const float2 db = ld2_16(dy + pixb * C + 2 * v, fp16);
          {
            const float xh0 = (xa.x - ch[k].mean[0]) * ch[k].rstd[0], xh1 = (xa.y - ch[k].mean[1]) * ch[k].rstd[1];
            const float g0 = da.x * act_grad(fmaf(xh0, ch[k].ga[0], ch[k].be[0]), act);
            const float g1 = da.y * act_grad(fmaf(xh1, ch[k].ga[1], ch[k].be[1]), act);
            acc[k][0] += g0; acc[k][1] = fmaf(g0, xh0, acc[k][1]);
            acc[k][2] += g1; acc[k][3] = fmaf(g1, xh1, acc[k][3]);
          }
          if (two) {
            const float xh0 = (xb.x - ch[k].mean[0]) * ch[k].rstd[0], xh1 = (xb.y - ch[k].mean[1]) * ch[k].rstd[1];
            const float g0 = db.x * act_grad(fmaf(xh0, ch[k].ga[0], ch[k].be[0]), act);
            const float g1 = db.y * act_grad(fmaf(xh1, ch[k].ga[1], ch[k].be[1]), act);
            acc[k][0] += g0; acc[k][1] = fmaf(g0, xh0, acc[k][1]);
            acc[k][2] += g1; acc[k][3] = fmaf(g1, xh1, acc[k][3]);
          }
        }
      }
    }
  }
  // fold the pixel lanes (fixed order) through shared memory, then one (g, g*xhat) pair per channel
  __shared__ float red[kGnThreads * 4];
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    if (k < geo.K) {  // (uniform across the CTA)
      __syncthreads();
      red[threadIdx.x * 4 + 0] = acc[k][0]; red[threadIdx.x * 4 + 1] = acc[k][1];
      red[threadIdx.x * 4 + 2] = acc[k][2]; red[threadIdx.x * 4 + 3] = acc[k][3];
      __syncthreads();
      const int v = tx + k * geo.TX;
      if (pl == 0 && v < cv) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < geo.P; ++q)
#pragma unroll
          for (int t = 0; t < 4; ++t) s[t] += red[(q * geo.TX + tx) * 4 + t];
        float4* dst = reinterpret_cast<float4*>(partial + ((size_t(b) * slabs + slab) * C + 2 * v) * 2);
        *dst = make_float4(s[0], s[1], s[2], s[3]);
      }
    }
  }
}

// pass 2: one CTA (4 warps) per (group, image): chan[b][c] = (A_c, B_c) summed over the slabs in a fixed order (lane-strided partial sums,
// butterfly reduction), coef[b][g] = (sum_c gamma_c A_c, sum_c gamma_c B_c) / n.  Warp w owns the group's channels w, w + 4, ...
__global__ void __launch_bounds__(128) gn_bwd_finalize_kernel(const float* __restrict__ partial, int slabs, int C, int HW,
                                                              const float* __restrict__ gamma, float* __restrict__ coef, float* __restrict__ chan) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const int g = blockIdx.x, b = blockIdx.y, cpg = C / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ float sw[4][2];
  float s1 = 0.f, s2 = 0.f;
  for (int k = warp; k < cpg; k += 4) {
    const int c = g * cpg + k;
    float a = 0.f, bb = 0.f;
    for (int s = lane; s < slabs; s += 32) {
      const float2 v = *reinterpret_cast<const float2*>(partial + ((size_t(b) * slabs + s) * C + c) * 2);
      a += v.x; bb += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); bb += __shfl_xor_sync(0xffffffffu, bb, o); }
    const float ga = gamma ? gamma[c] : 1.0f;
    s1 = fmaf(ga, a, s1); s2 = fmaf(ga, bb, s2);
    if (lane == 0 && chan) { chan[(size_t(b) * C + c) * 2] = a; chan[(size_t(b) * C + c) * 2 + 1] = bb; }
  }
  if (lane == 0) { sw[warp][0] = s1; sw[warp][1] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float inv_n = 1.0f / (float(HW) * float(cpg));
    coef[(size_t(b) * 32 + g) * 2] = ((sw[0][0] + sw[1][0]) + (sw[2][0] + sw[3][0])) * inv_n;
    coef[(size_t(b) * 32 + g) * 2 + 1] = ((sw[0][1] + sw[1][1]) + (sw[2][1] + sw[3][1])) * inv_n;
  }
}

// dgamma_c = scale * sum_b B_c, dbeta_c = scale * sum_b A_c
__global__ void gn_bwd_affine_kernel(const float* __restrict__ chan, int B, int C, float scale, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, bb = 0.f;
  for (int b = 0; b < B; ++b) { a += chan[(size_t(b) * C + c) * 2]; bb += chan[(size_t(b) * C + c) * 2 + 1]; }
  if (dbeta) dbeta[c] = a * scale;
  if (dgamma) dgamma[c] = bb * scale;
}

// pass 3: dx (+ extra) -> 16-bit [B,HW,C], or fp32 split over the two concat sources (each stored or accumulated)
template <bool IN16, int KT>
__global__ void __launch_bounds__(kGnThreads) gn_bwd_apply_kernel(const void* __restrict__ x0, int C0, const void* __restrict__ x1, int C1, int HW,
                                                                  int slab_pix, const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float eps, int act,
                                                                  const uint16_t* __restrict__ dy, const float* __restrict__ coef,
                                                                  const float* __restrict__ extra, uint16_t* __restrict__ out16,
                                                                  float* __restrict__ dx0, int acc0, float* __restrict__ dx1, int acc1, int fp16) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const int C = C0 + C1, cv = C / 2, cpg = C / 32;
  const GnBwdGeo geo = gn_bwd_geo(C);
  const int b = blockIdx.y, slab = blockIdx.x;
  const int tx = threadIdx.x % geo.TX, pl = threadIdx.x / geo.TX;
  if (pl >= geo.P) return;
  const float inv_n = 1.0f / (float(HW) * float(cpg));
  GnChan ch[KT];
  float c1[KT][2], c2[KT][2];
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const int v = tx + k * geo.TX;
    if (k < geo.K && v < cv) {
      gn_load_chan(stats + size_t(b) * 64, gamma, beta, 2 * v, cpg, inv_n, eps, ch[k]);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int g = (2 * v + t) / cpg;
        c1[k][t] = coef[(size_t(b) * 32 + g) * 2];
        c2[k][t] = coef[(size_t(b) * 32 + g) * 2 + 1];
      }
    }
  }
  const int p0 = slab * slab_pix, p1 = min(HW, p0 + slab_pix);
  for (int p = p0 + pl; p < p1; p += 2 * geo.P) {  // two pixels per trip: all their loads are issued before the arithmetic of either
    const size_t pixs[2] = {size_t(b) * HW + p, size_t(b) * HW + p + geo.P};
    const bool two = p + geo.P < p1;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const int v = tx + k * geo.TX;
      if (k < geo.K && v < cv) {
        const int c = 2 * v;
        float2 xv[2], dv[2], ev[2], ov[2];
        float* dst[2] = {nullptr, nullptr};
        int accf = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u == 1 && !two) continue;
          const size_t pix = pixs[u];
          xv[u] = gn_load_x<IN16>(x0, C0, x1, C1, pix, c, fp16);
          dv[u] = ld2_16(dy + pix * C + c, fp16);
          ev[u] = extra ? *reinterpret_cast<const float2*>(extra + pix * C + c) : make_float2(0.f, 0.f);
          if (c < C0) { if (dx0) { dst[u] = dx0 + pix * C0 + c; accf = acc0; } }
          else if (dx1) { dst[u] = dx1 + pix * C1 + (c - C0); accf = acc1; }
          ov[u] = (dst[u] && accf) ? *reinterpret_cast<const float2*>(dst[u]) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u == 1 && !two) continue;
          const size_t pix = pixs[u];
          const float xh0 = (xv[u].x - ch[k].mean[0]) * ch[k].rstd[0], xh1 = (xv[u].y - ch[k].mean[1]) * ch[k].rstd[1];
          const float g0 = dv[u].x * act_grad(fmaf(xh0, ch[k].ga[0], ch[k].be[0]), act);
          const float g1 = dv[u].y * act_grad(fmaf(xh1, ch[k].ga[1], ch[k].be[1]), act);
          const float r0 = ch[k].rstd[0] * (g0 * ch[k].ga[0] - c1[k][0] - xh0 * c2[k][0]) + ev[u].x;
          const float r1 = ch[k].rstd[1] * (g1 * ch[k].ga[1] - c1[k][1] - xh1 * c2[k][1]) + ev[u].y;
          if (out16) st2_16(out16 + pix * C + c, r0, r1, fp16);
          if (dst[u]) *reinterpret_cast<float2*>(dst[u]) = make_float2(r0 + ov[u].x, r1 + ov[u].y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// one warp per row: dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)), g = dy; statistics recomputed from x
constexpr int kLnMaxPerLane = 40;  // C <= 1280
// PER = C / 32 elements per lane (10 / 20 / 40 for C = 320 / 640 / 1280): compile-time, so the row lives in exactly PER registers (a runtime bound
// made every launch pay for the 1280-channel case: 80 live values, 2 CTAs per SM)
template <int PER>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, int M, int C, const float* __restrict__ gamma, float eps,
                                                     const uint16_t* __restrict__ dy, int fp16, float* __restrict__ dx, int accumulate,
                                                     uint16_t* __restrict__ dx16) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int per = PER;  // C % 64 == 0: every lane owns `per` (even) elements, interleaved in pairs for coalescing
  const float* xr = x + size_t(row) * C;
  const uint16_t* dr = dy + size_t(row) * C;
  float xv[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; i += 2) {
    if (i < per) {
      const float2 t = *reinterpret_cast<const float2*>(xr + (i / 2) * 64 + lane * 2);
      xv[i] = t.x; xv[i + 1] = t.y;
      s += t.x + t.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / float(C);
  float vs = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if (i < per) { const float d = xv[i] - mean; vs = fmaf(d, d, vs); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
  const float rstd = rsqrtf(vs / float(C) + eps);
  float gg[PER];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < PER; i += 2) {
    if (i < per) {
      const int c = (i / 2) * 64 + lane * 2;
      const float2 d = ld2_16(dr + c, fp16);
      const float2 ga = *reinterpret_cast<const float2*>(gamma + c);
      xv[i] = (xv[i] - mean) * rstd; xv[i + 1] = (xv[i + 1] - mean) * rstd;
      gg[i] = d.x * ga.x; gg[i + 1] = d.y * ga.y;
      s1 += gg[i] + gg[i + 1];
      s2 = fmaf(gg[i], xv[i], s2); s2 = fmaf(gg[i + 1], xv[i + 1], s2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  s1 /= float(C); s2 /= float(C);
  float* orow = dx + size_t(row) * C;
#pragma unroll
  for (int i = 0; i < PER; i += 2) {
    if (i < per) {
      const int c = (i / 2) * 64 + lane * 2;
      float2 o = make_float2(rstd * (gg[i] - s1 - xv[i] * s2), rstd * (gg[i + 1] - s1 - xv[i + 1] * s2));
      if (accumulate) { const float2 old = *reinterpret_cast<const float2*>(orow + c); o.x += old.x; o.y += old.y; }
      *reinterpret_cast<float2*>(orow + c) = o;
      if (dx16) st2_16(dx16 + size_t(row) * C + c, o.x, o.y, fp16);  // the 16-bit operand copy of the updated gradient (next dgrad GEMM)
    }
  }
}

// ------------------------------------------------------------------------------------------------ GEGLU (natural column order)
// raw [M, 2H] = (hidden | gate) from ff.net.0.proj; out = hidden * gelu(gate) (exact erf GELU, as diffusers' GEGLU)
__global__ void geglu_fwd_kernel(const uint16_t* __restrict__ raw, long M, int H, int fp16, uint16_t* __restrict__ out) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = (long(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i >= M * H) return;
  const long m = i / H; const int j = int(i - m * H);
  const float2 h = ld2_16(raw + m * 2 * H + j, fp16), g = ld2_16(raw + m * 2 * H + H + j, fp16);
  const float a = h.x * 0.5f * g.x * (1.0f + erff(g.x * 0.70710678118654752f));
  const float b = h.y * 0.5f * g.y * (1.0f + erff(g.y * 0.70710678118654752f));
  st2_16(out + i, a, b, fp16);
}
__device__ __forceinline__ void geglu_grad(float h, float g, float d, float& dh, float& dg) {
  const float phi = 0.5f * (1.0f + erff(g * 0.70710678118654752f));           // Phi(g)
  const float pdf = 0.3989422804014327f * __expf(-0.5f * g * g);               // phi(g)
  dh = d * g * phi;
  dg = d * h * (phi + g * pdf);
}
__global__ void geglu_bwd_kernel(const uint16_t* __restrict__ raw, const uint16_t* __restrict__ dout, long M, int H, int fp16,
                                 uint16_t* __restrict__ draw) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = (long(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i >= M * H) return;
  const long m = i / H; const int j = int(i - m * H);
  const float2 h = ld2_16(raw + m * 2 * H + j, fp16), g = ld2_16(raw + m * 2 * H + H + j, fp16), d = ld2_16(dout + i, fp16);
  float dh0, dg0, dh1, dg1;
  geglu_grad(h.x, g.x, d.x, dh0, dg0);
  geglu_grad(h.y, g.y, d.y, dh1, dg1);
  st2_16(draw + m * 2 * H + j, dh0, dh1, fp16);
  st2_16(draw + m * 2 * H + H + j, dg0, dg1, fp16);
}

// ------------------------------------------------------------------------------------------------ small data-movement kernels
// per-image column sums of a 16-bit [B, HW, C] tensor -> out[b * ldo + c] (fp32): the gradient of a per-image row bias (time embedding).
// A thread-block cluster of S CTAs (gridDim.z = cluster dim z) splits the pixel axis; rank 0 adds the S partial sums in rank order through
// distributed shared memory (deterministic, no atomics, no second launch).  At 2 images the unsplit version ran on B * C/64 = 10 CTAs.
__global__ void __launch_bounds__(256) colsum_img_kernel(const uint16_t* __restrict__ x, int HW, int C, int fp16, float* __restrict__ out, int ldo) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int S = int(cluster.num_blocks()), slab = int(cluster.block_rank());
  const int b = blockIdx.y;
  const int cp = blockIdx.x * 32 + (threadIdx.x & 31);  // channel pair
  const int pl = threadIdx.x >> 5;                        // 8 pixel lanes
  const int per = (HW + S - 1) / S;
  const int p0 = slab * per, p1 = min(HW, p0 + per);
  float a0 = 0.f, a1 = 0.f;
  if (2 * cp < C)
    for (int p = p0 + pl; p < p1; p += 8) {
      const float2 v = ld2_16(x + (size_t(b) * HW + p) * C + 2 * cp, fp16);
      a0 += v.x; a1 += v.y;
    }
  __shared__ float red[256 * 2];
  __shared__ float tot[64];
  red[threadIdx.x * 2] = a0; red[threadIdx.x * 2 + 1] = a1;
  __syncthreads();
  if (pl == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int q = 0; q < 8; ++q) { s0 += red[(q * 32 + threadIdx.x) * 2]; s1 += red[(q * 32 + threadIdx.x) * 2 + 1]; }
    tot[threadIdx.x * 2] = s0; tot[threadIdx.x * 2 + 1] = s1;
  }
  cluster.sync();  // every CTA's totals are in its shared memory
  if (slab == 0 && pl == 0 && 2 * cp < C) {
    float s0 = 0.f, s1 = 0.f;
    for (int r = 0; r < S; ++r) {
      const float* t = cluster.map_shared_rank(tot, r);
      s0 += t[threadIdx.x * 2]; s1 += t[threadIdx.x * 2 + 1];
    }
    out[size_t(b) * ldo + 2 * cp] = s0; out[size_t(b) * ldo + 2 * cp + 1] = s1;
  }
  cluster.sync();  // keep the peers' shared memory alive until rank 0 has read it
}

// 16-bit [B,h,w,C] -> [B,2h,2w,C] with the values at the even positions and zeros elsewhere (operand of a stride-2 conv's dgrad)
__global__ void zero_stuff2x_kernel(const uint4* __restrict__ x, int h, int w, int C8, long total, uint4* __restrict__ out) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = int(i % C8);
  long r = i / C8;
  const int X = int(r % (2 * w)); r /= 2 * w;
  const int Y = int(r % (2 * h)); const long b = r / (2 * h);
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (!(X & 1) && !(Y & 1)) v = x[((b * h + (Y >> 1)) * w + (X >> 1)) * C8 + c];
  out[i] = v;
}

// fp32 [B,2h,2w,C] -> [B,h,w,C]: sum of each 2x2 block (backward of nearest-2x upsampling), stored or accumulated
__global__ void sum2x2_kernel(const float4* __restrict__ x, int h, int w, int C4, long total, float4* __restrict__ out, int accumulate) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = int(i % C4);
  long r = i / C4;
  const int X = int(r % w); r /= w;
  const int Y = int(r % h); const long b = r / h;
  const long base = ((b * 2 * h + 2 * Y) * 2 * w + 2 * X) * C4 + c;
  const float4 a = x[base], bq = x[base + C4], cq = x[base + long(2 * w) * C4], d = x[base + long(2 * w) * C4 + C4];
  float4 o = make_float4(a.x + bq.x + cq.x + d.x, a.y + bq.y + cq.y + d.y, a.z + bq.z + cq.z + d.z, a.w + bq.w + cq.w + d.w);
  if (accumulate) { const float4 old = out[i]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
  out[i] = o;
}

// dz[b,p,c] (16-bit NHWC) = scale * dout[b,c,p] * (out[b,c,p] > 0): ReLU backward of the projections' final pass + NCHW -> NHWC
__global__ void __launch_bounds__(256) relu_bwd_nchw_kernel(const float* __restrict__ dout, const float* __restrict__ outv, int C, int HW, float scale,
                                                            int fp16, uint16_t* __restrict__ dz) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) {
      const size_t i = (size_t(b) * C + c) * HW + p;
      v = outv[i] > 0.0f ? dout[i] * scale : 0.0f;
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    if (c < C && p < HW) dz[(size_t(b) * HW + p) * C + c] = cvt_16(tile[tx][r], fp16);
  }
}

// d_cond_emb[b,j] = scale * d_act[b,j] * silu'(emb[b,j] + cond_emb[b,j])   (time path: emb_act = silu(emb + cond_emb))
__global__ void temb_silu_bwd_kernel(const float* __restrict__ d_act, const float* __restrict__ emb, const float* __restrict__ cond_emb, long n,
                                     float scale, float* __restrict__ d_cond_emb) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d_cond_emb[i] = scale * d_act[i] * act_grad(emb[i] + cond_emb[i], ACT_SILU);
}

__global__ void scale_copy_kernel(const float* __restrict__ src, long n, float scale, float* __restrict__ dst) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * scale;
}

}  // namespace

const char* scale_copy_f32(const float* src, long n, float scale, float* dst, cudaStream_t st) {
  launch_k(scale_copy_kernel, dim3(unsigned((n + 255) / 256)), dim3(256), 0, st, src, n, scale, dst);
  return cudaGetLastError() == cudaSuccess ? nullptr : "scale_copy_f32 launch failed";
}

// ------------------------------------------------------------------------------------------------ launchers
static inline int gn_bwd_slab_pix(int HW) { return HW >= 16384 ? 64 : 16; }  // >= 256 CTAs per image batch at every UNet level
int groupnorm_bwd_slabs(int HW) { const int sp = gn_bwd_slab_pix(HW); return (HW + sp - 1) / sp; }

const char* groupnorm_bwd(const void* x0, int C0, const void* x1, int C1, int B, int HW, int in16, const float* stats, const float* gamma,
                          const float* beta, float eps, int act, const void* dy16, int fp16, float* partial, float* coef, float* chan,
                          const float* extra, void* out16, float* dx0, int acc0, float* dx1, int acc1, float* dgamma, float* dbeta,
                          float affine_scale, cudaStream_t st) {
  const int C = C0 + C1;
  if (C % 64 != 0 || C0 % 2 != 0 || C > 2 * kGnThreads * kGnKMax) return "groupnorm_bwd: unsupported channel count";
  if ((dgamma || dbeta) && !chan) return "groupnorm_bwd: affine gradients need the per-channel scratch";
  const int sp = gn_bwd_slab_pix(HW), slabs = (HW + sp - 1) / sp;
  const dim3 grid(slabs, B);
  const GnBwdGeo geo = gn_bwd_geo(C);
  const int threads = geo.TX * geo.P;
  const uint16_t* dy = static_cast<const uint16_t*>(dy16);
#define GN_BWD_K(KERNEL, ...)                                                                     \
  switch (geo.K) {                                                                                  \
    case 1: if (in16) launch_k(KERNEL<true, 1>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); else launch_k(KERNEL<false, 1>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); break; \
    case 2: if (in16) launch_k(KERNEL<true, 2>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); else launch_k(KERNEL<false, 2>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); break; \
    case 3: if (in16) launch_k(KERNEL<true, 3>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); else launch_k(KERNEL<false, 3>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); break; \
    case 4: if (in16) launch_k(KERNEL<true, 4>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); else launch_k(KERNEL<false, 4>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); break; \
    default: if (in16) launch_k(KERNEL<true, 5>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); else launch_k(KERNEL<false, 5>, dim3(grid), dim3(threads), 0, st, __VA_ARGS__); break; \
  }
  GN_BWD_K(gn_bwd_reduce_kernel, x0, C0, x1, C1, HW, sp, stats, gamma, beta, eps, act, dy, fp16, partial)
  launch_k(gn_bwd_finalize_kernel, dim3(dim3(32, B)), dim3(128), 0, st, partial, slabs, C, HW, gamma, coef, chan);
  if (dgamma || dbeta) launch_k(gn_bwd_affine_kernel, dim3((C + 127) / 128), dim3(128), 0, st, chan, B, C, affine_scale, dgamma, dbeta);
  if (out16 || dx0 || dx1) {
    uint16_t* o16 = static_cast<uint16_t*>(out16);
    GN_BWD_K(gn_bwd_apply_kernel, x0, C0, x1, C1, HW, sp, stats, gamma, beta, eps, act, dy, coef, extra, o16, dx0, acc0, dx1, acc1, fp16)
  }
#undef GN_BWD_K
  return cudaGetLastError() == cudaSuccess ? nullptr : "groupnorm_bwd launch failed";
}

const char* layernorm_bwd(const float* x, int M, int C, const float* gamma, float eps, const void* dy16, int fp16, float* dx, int accumulate,
                          cudaStream_t st, void* dx16) {
  if (C % 64 != 0 || C > 32 * kLnMaxPerLane) return "layernorm_bwd: C must be a multiple of 64, <= 1280";
  const uint16_t* dy = static_cast<const uint16_t*>(dy16);
  uint16_t* d16 = static_cast<uint16_t*>(dx16);
  const unsigned grid = (M + 7) / 8;
  switch (C / 32) {
    case 2: launch_k(ln_bwd_kernel<2>, dim3(grid), dim3(256), 0, st, x, M, C, gamma, eps, dy, fp16, dx, accumulate, d16); break;
    case 4: launch_k(ln_bwd_kernel<4>, dim3(grid), dim3(256), 0, st, x, M, C, gamma, eps, dy, fp16, dx, accumulate, d16); break;
    case 10: launch_k(ln_bwd_kernel<10>, dim3(grid), dim3(256), 0, st, x, M, C, gamma, eps, dy, fp16, dx, accumulate, d16); break;
    case 20: launch_k(ln_bwd_kernel<20>, dim3(grid), dim3(256), 0, st, x, M, C, gamma, eps, dy, fp16, dx, accumulate, d16); break;
    case 40: launch_k(ln_bwd_kernel<40>, dim3(grid), dim3(256), 0, st, x, M, C, gamma, eps, dy, fp16, dx, accumulate, d16); break;
    default: return "layernorm_bwd: C must be 64, 128, 320, 640 or 1280";
  }
  return cudaGetLastError() == cudaSuccess ? nullptr : "layernorm_bwd launch failed";
}

const char* geglu_fwd(const void* raw16, long M, int H, void* out16, int fp16, cudaStream_t st) {
  if (H % 2 != 0) return "geglu: H must be even";
  const long n = M * H / 2;
  launch_k(geglu_fwd_kernel, dim3(unsigned((n + 255) / 256)), dim3(256), 0, st, static_cast<const uint16_t*>(raw16), M, H, fp16, static_cast<uint16_t*>(out16));
  return cudaGetLastError() == cudaSuccess ? nullptr : "geglu_fwd launch failed";
}
const char* geglu_bwd(const void* raw16, const void* dout16, long M, int H, void* draw16, int fp16, cudaStream_t st) {
  if (H % 2 != 0) return "geglu: H must be even";
  const long n = M * H / 2;
  launch_k(geglu_bwd_kernel, dim3(unsigned((n + 255) / 256)), dim3(256), 0, st, static_cast<const uint16_t*>(raw16), static_cast<const uint16_t*>(dout16), M, H, fp16,
                                                            static_cast<uint16_t*>(draw16));
  return cudaGetLastError() == cudaSuccess ? nullptr : "geglu_bwd launch failed";
}

const char* colsum_per_image(const void* x16, int B, int HW, int C, int fp16, float* out, int ldo, cudaStream_t st) {
  if (C % 2 != 0) return "colsum_per_image: C must be even";
  int S = 8;  // cluster size (portable maximum): fewer slabs for small pixel counts
  while (S > 1 && HW < 64 * S) S >>= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((C / 2 + 31) / 32, B, S);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = S;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, colsum_img_kernel, static_cast<const uint16_t*>(x16), HW, C, fp16, out, ldo) != cudaSuccess)
    return "colsum_per_image launch failed";
  return cudaGetLastError() == cudaSuccess ? nullptr : "colsum_per_image launch failed";
}

const char* zero_stuff2x(const void* x16, int B, int h, int w, int C, void* out16, cudaStream_t st) {
  if (C % 8 != 0) return "zero_stuff2x: C must be a multiple of 8";
  const long total = long(B) * 2 * h * 2 * w * (C / 8);
  launch_k(zero_stuff2x_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), 0, st, static_cast<const uint4*>(x16), h, w, C / 8, total, static_cast<uint4*>(out16));
  return cudaGetLastError() == cudaSuccess ? nullptr : "zero_stuff2x launch failed";
}

const char* sum2x2(const float* x, int B, int h, int w, int C, float* out, int accumulate, cudaStream_t st) {
  if (C % 4 != 0) return "sum2x2: C must be a multiple of 4";
  const long total = long(B) * h * w * (C / 4);
  launch_k(sum2x2_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), 0, st, reinterpret_cast<const float4*>(x), h, w, C / 4, total, reinterpret_cast<float4*>(out),
                                                             accumulate);
  return cudaGetLastError() == cudaSuccess ? nullptr : "sum2x2 launch failed";
}

const char* relu_bwd_nchw_to_nhwc16(const float* dout, const float* out, int B, int C, int HW, float scale, void* dz16, int fp16, cudaStream_t st) {
  launch_k(relu_bwd_nchw_kernel, dim3(dim3((HW + 31) / 32, (C + 31) / 32, B)), dim3(256), 0, st, dout, out, C, HW, scale, fp16, static_cast<uint16_t*>(dz16));
  return cudaGetLastError() == cudaSuccess ? nullptr : "relu_bwd_nchw_to_nhwc16 launch failed";
}

const char* temb_silu_bwd(const float* d_act, const float* emb, const float* cond_emb, long n, float scale, float* d_cond_emb, cudaStream_t st) {
  launch_k(temb_silu_bwd_kernel, dim3(unsigned((n + 255) / 256)), dim3(256), 0, st, d_act, emb, cond_emb, n, scale, d_cond_emb);
  return cudaGetLastError() == cudaSuccess ? nullptr : "temb_silu_bwd launch failed";
}

}  // namespace madm

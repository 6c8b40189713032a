// Weight packing: fp32 PyTorch parameter layouts -> bf16 K-major GEMM operands consumed by gemm_tc.cu through TMA.
// The LoRA update of the active adapter is folded here, once per adapter switch: W' = W + (alpha/r) * B @ A in fp32,
// rounded to bf16 once, so the forward path runs the plain projection with no extra skinny GEMMs
// (reference: peft LoRA Linear on to_q/to_k/to_v/to_out.0, modeling/meta_arch/mtmadise.py:115-147).
#include "cvt.cuh"
#include "kernels.h"

namespace madm {

// out[n, tap*Cpad + c] = w[n, c, tap]  (w is [N, C, taps] contiguous = [N,C,kh,kw]); zero for c >= C and k >= taps*Cpad
__global__ void pack_conv_kernel(const float* __restrict__ w, int N, int C, int taps, int Cpad, int Kpad, int ldo, int fp16,
                                 uint16_t* __restrict__ out, const float* __restrict__ row_scale) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = long(N) * Kpad;
  if (i >= total) return;
  const int k = int(i % Kpad);
  const int n = int(i / Kpad);
  float v = 0.f;
  if (k < taps * Cpad) {
    const int tap = k / Cpad, c = k % Cpad;
    if (c < C) v = w[(size_t(n) * C + c) * taps + tap];
  }
  if (row_scale) v *= row_scale[n];  // eval-mode BatchNorm scale of output channel n folded into the weight (fp32, before rounding)
  out[size_t(n) * ldo + k] = cvt_16(v, fp16);
}

const char* pack_conv_weight(const float* w, int N, int C, int taps, int Cpad, int Kpad, int ldo, void* out, int fp16, cudaStream_t st,
                             const float* row_scale) {
  const long total = long(N) * Kpad;
  pack_conv_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(w, N, C, taps, Cpad, Kpad, ldo, fp16, reinterpret_cast<uint16_t*>(out),
                                                                  row_scale);
  return cudaGetLastError() == cudaSuccess ? nullptr : "pack_conv_weight launch failed";
}

// ---- dgrad operands (first building block of SURVEY §8 row f-3): the input gradient of a stride-1 conv / a linear is the same implicit
// GEMM run on the output gradient with the weight's in / out roles swapped (and the filter taps mirrored):
//   conv:   dX[b,y,x,ci] = sum_{ky,kx,co} dY[b, y-(ky-1), x-(kx-1), co] * W[co,ci,ky,kx]
//           -> out[ci, tap'*CoPad + co] = W[co, ci, taps-1-tap']     (tap' = the forward kernel's tap order, so seg_3x3 offsets apply as is)
//   linear: dX = dY (W + s B A)   -> out[k, n] = W'[n, k]
__global__ void pack_conv_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int CoPad, int Kpad, int ldo, int fp16,
                                       uint16_t* __restrict__ out) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = long(Cin) * Kpad;
  if (i >= total) return;
  const int k = int(i % Kpad);
  const int ci = int(i / Kpad);
  float v = 0.f;
  if (k < taps * CoPad) {
    const int tap = k / CoPad, co = k % CoPad;
    if (co < Cout) v = w[(size_t(co) * Cin + ci) * taps + (taps - 1 - tap)];
  }
  out[size_t(ci) * ldo + k] = cvt_16(v, fp16);
}

const char* pack_conv_dgrad_weight(const float* w, int Cout, int Cin, int taps, int CoPad, int Kpad, int ldo, void* out, int fp16, cudaStream_t st) {
  if (taps != 1 && taps != 9) return "pack_conv_dgrad: 1x1 or 3x3 (stride 1) filters only";
  const long total = long(Cin) * Kpad;
  pack_conv_dgrad_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(w, Cout, Cin, taps, CoPad, Kpad, ldo, fp16, reinterpret_cast<uint16_t*>(out));
  return cudaGetLastError() == cudaSuccess ? nullptr : "pack_conv_dgrad_weight launch failed";
}

// out[k, n] = 16-bit( w[n,k] + scale * sum_j lb[n,j] * la[j,k] )   (the transpose of pack_linear_kernel's result)
__global__ void pack_linear_dgrad_kernel(const float* __restrict__ w, int N, int K, const float* __restrict__ la, const float* __restrict__ lb,
                                         int r, float scale, int ldo, int fp16, uint16_t* __restrict__ out) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= long(N) * K) return;
  const int n = int(i % N);  // n fastest: coalesced writes of the transposed matrix
  const int k = int(i / N);
  float v = w[size_t(n) * K + k];
  if (la != nullptr) {
    float acc = 0.f;
    for (int j = 0; j < r; ++j) acc += lb[size_t(n) * r + j] * la[size_t(j) * K + k];
    v += scale * acc;
  }
  out[size_t(k) * ldo + n] = cvt_16(v, fp16);
}

const char* pack_linear_dgrad_weight(const float* w, int N, int K, const float* la, const float* lb, int r, float scale, int ldo, void* out,
                                     int fp16, cudaStream_t st) {
  const long total = long(N) * K;
  pack_linear_dgrad_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(w, N, K, la, lb, r, scale, ldo, fp16, reinterpret_cast<uint16_t*>(out));
  return cudaGetLastError() == cudaSuccess ? nullptr : "pack_linear_dgrad_weight launch failed";
}

// Many linears in one launch (the 128 LoRA-wrapped projections are re-folded after every optimizer step and adapter switch of a training
// step: one launch per 48 of them instead of one each).  32 x 32 tiles: w is read along k, the LoRA factors of the tile go through shared
// memory, the result is written along k (forward operand) or, transposed through the tile, along n (input-gradient operand).
__global__ void __launch_bounds__(256) pack_lora_multi_kernel(const __grid_constant__ LoraPackTable t) {
  const LoraPackEntry e = t.e[blockIdx.z];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  if (n0 >= e.N || k0 >= e.K) return;
  __shared__ float sw[32][33];
  __shared__ float sb[32][17];
  __shared__ float sa[16][33];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  for (int i = ty; i < 32; i += 8) sw[i][tx] = (n0 + i < e.N && k0 + tx < e.K) ? e.w[size_t(n0 + i) * e.K + k0 + tx] : 0.f;
  const bool lora = e.la != nullptr;
  if (lora) {
    for (int i = tid; i < 32 * 16; i += 256) {
      const int r = i >> 4, j = i & 15;
      sb[r][j] = (n0 + r < e.N && j < e.r) ? e.lb[size_t(n0 + r) * e.r + j] : 0.f;
      const int jj = i >> 5, c = i & 31;
      sa[jj][c] = (jj < e.r && k0 + c < e.K) ? e.la[size_t(jj) * e.K + k0 + c] : 0.f;
    }
  }
  __syncthreads();
  if (lora) {
    for (int i = ty; i < 32; i += 8) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = fmaf(sb[i][j], sa[j][tx], acc);
      sw[i][tx] = fmaf(t.scale, acc, sw[i][tx]);
    }
    __syncthreads();
  }
  uint16_t* out = reinterpret_cast<uint16_t*>(e.out);
  if (!t.transpose) {
    for (int i = ty; i < 32; i += 8)
      if (n0 + i < e.N && k0 + tx < e.K) out[size_t(n0 + i) * e.ldo + k0 + tx] = cvt_16(sw[i][tx], t.fp16);
  } else {
    for (int i = ty; i < 32; i += 8)
      if (k0 + i < e.K && n0 + tx < e.N) out[size_t(k0 + i) * e.ldo + n0 + tx] = cvt_16(sw[tx][i], t.fp16);
  }
}

const char* pack_lora_multi(const LoraPackEntry* entries, int n, float scale, int transpose, int fp16, cudaStream_t st) {
  for (int first = 0; first < n; first += kLoraPackMax) {
    LoraPackTable t;
    t.n = n - first < kLoraPackMax ? n - first : kLoraPackMax;
    t.scale = scale; t.transpose = transpose; t.fp16 = fp16;
    int maxN = 0, maxK = 0;
    for (int i = 0; i < t.n; ++i) {
      t.e[i] = entries[first + i];
      if (t.e[i].la && t.e[i].r > 16) return "pack_lora_multi: LoRA rank must be <= 16";
      if (t.e[i].N > maxN) maxN = t.e[i].N;
      if (t.e[i].K > maxK) maxK = t.e[i].K;
    }
    pack_lora_multi_kernel<<<dim3((maxK + 31) / 32, (maxN + 31) / 32, t.n), dim3(32, 8), 0, st>>>(t);
    if (cudaGetLastError() != cudaSuccess) return "pack_lora_multi launch failed";
  }
  return nullptr;
}

// out[n, k] = bf16( w[n,k] + scale * sum_j lb[n,j] * la[j,k] )
__global__ void pack_linear_kernel(const float* __restrict__ w, int N, int K, const float* __restrict__ la,
                                   const float* __restrict__ lb, int r, float scale, int ldo, int fp16, uint16_t* __restrict__ out) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= long(N) * K) return;
  const int k = int(i % K);
  const int n = int(i / K);
  float v = w[i];
  if (la != nullptr) {
    float acc = 0.f;
    for (int j = 0; j < r; ++j) acc += lb[size_t(n) * r + j] * la[size_t(j) * K + k];
    v += scale * acc;
  }
  out[size_t(n) * ldo + k] = cvt_16(v, fp16);
}

const char* pack_linear_weight(const float* w, int N, int K, const float* la, const float* lb, int r, float scale, int ldo, void* out,
                               int fp16, cudaStream_t st) {
  const long total = long(N) * K;
  pack_linear_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(w, N, K, la, lb, r, scale, ldo, fp16, reinterpret_cast<uint16_t*>(out));
  return cudaGetLastError() == cudaSuccess ? nullptr : "pack_linear_weight launch failed";
}

// GEGLU proj weight W[2*C4, K]: value rows [0,C4), gate rows [C4, 2*C4).  Packed row 128*t + j (j<64) = value row 64*t + j,
// packed row 128*t + 64 + j = gate row C4 + 64*t + j.
__global__ void pack_geglu_kernel(const float* __restrict__ w, const float* __restrict__ bias, int C4, int K, int fp16,
                                  uint16_t* __restrict__ out, float* __restrict__ out_bias) {
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total = long(2 * C4) * K;
  if (i >= total) return;
  const int k = int(i % K);
  const int pr = int(i / K);
  const int t = pr / 128, j = pr % 128;
  const int src = (j < 64) ? (64 * t + j) : (C4 + 64 * t + (j - 64));
  out[i] = cvt_16(w[size_t(src) * K + k], fp16);
  if (k == 0 && out_bias) out_bias[pr] = bias ? bias[src] : 0.f;
}

const char* pack_geglu_weight(const float* w, const float* bias, int C4, int K, void* out, float* out_bias, int fp16, cudaStream_t st) {
  if (C4 % 64 != 0) return "pack_geglu_weight: 4C must be a multiple of 64";
  const long total = long(2 * C4) * K;
  pack_geglu_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(w, bias, C4, K, fp16, reinterpret_cast<uint16_t*>(out), out_bias);
  return cudaGetLastError() == cudaSuccess ? nullptr : "pack_geglu_weight launch failed";
}

// latents = scale * (Wq[0:4,:] @ (conv_out(x) ) + bq[0:4]) ; conv_out(x) = Wout * x + bout
//   => W'[o, tap*C + c] = scale * sum_j Wq[o,j] * Wout[j,c,tap],  b'[o] = scale * (sum_j Wq[o,j]*bout[j] + bq[o]);  rows 4..15 zero.
__global__ void pack_vae_head_kernel(const float* __restrict__ w_out, const float* __restrict__ b_out, const float* __restrict__ w_q,
                                     const float* __restrict__ b_q, float scale, int C, int fp16, uint16_t* __restrict__ out,
                                     float* __restrict__ out_bias) {
  const int K = 9 * C;
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= long(16) * K) return;
  const int k = int(i % K);
  const int o = int(i / K);
  float v = 0.f;
  if (o < 4) {
    const int tap = k / C, c = k % C;
    for (int j = 0; j < 8; ++j) v += w_q[o * 8 + j] * w_out[(size_t(j) * C + c) * 9 + tap];
    v *= scale;
  }
  out[i] = cvt_16(v, fp16);
  if (k == 0) {
    float b = 0.f;
    if (o < 4) {
      for (int j = 0; j < 8; ++j) b += w_q[o * 8 + j] * b_out[j];
      b = scale * (b + b_q[o]);
    }
    out_bias[o] = b;
  }
}

const char* pack_vae_latent_head(const float* w_out, const float* b_out, const float* w_q, const float* b_q, float scale, int C,
                                 void* out, float* out_bias, int fp16, cudaStream_t st) {
  const long total = long(16) * 9 * C;
  pack_vae_head_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(w_out, b_out, w_q, b_q, scale, C, fp16,
                                                                      reinterpret_cast<uint16_t*>(out), out_bias);
  return cudaGetLastError() == cudaSuccess ? nullptr : "pack_vae_latent_head launch failed";
}

}  // namespace madm

// Weight gradients of the LoRA training step's trainable set: the contraction runs over the PIXEL axis,
//   dW[n, k] = alpha * sum_m dY[m, n] * X[m, k]
// with both operands stored pixel-major (NHWC rows), i.e. a "TN" GEMM.  Used for the LoRA factors (skinny: k = rank 16), the GN-bottleneck
// projections' 1x1 convs and, with an implicit im2col of X (zero padding = bounds check on the shifted pixel), their 3x3 conv.
// These are small next to the dgrad GEMMs (LoRA + projections only: the base UNet weights are frozen), so the kernel is a plain
// warp-level tensor-core kernel (wmma 16x16x16, fp32 accumulate) with a deterministic split over M: every split writes its own fp32
// partial tile, a second kernel reduces them in fixed order, scales, and scatters into the parameter's PyTorch layout.
#include "kernels.h"
#include "launch.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mma.h>

namespace madm {

namespace {

using namespace nvcuda;

template <typename T> __device__ __forceinline__ T from_float_t(float v);
template <> __device__ __forceinline__ __half from_float_t<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float_t<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

constexpr int WG_TN = 64;   // output rows (columns of dY) per CTA
constexpr int WG_MC = 64;   // pixels per smem chunk
constexpr int WG_PAD = 8;

struct WgradParams {
  const uint16_t* a; int lda;
  const uint16_t* b; int ldb;
  int M, N, K, taps;
  int H, W;           // conv mode (taps == 9): pixel grid of one image
  int m_per_split;    // multiple of WG_MC
  float* part;        // [splits][taps][N][K]
};

template <typename T, int TK>
__global__ void __launch_bounds__(128) wgrad_kernel(const WgradParams p) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  __shared__ __align__(32) T As[WG_MC][WG_TN + WG_PAD];
  __shared__ __align__(32) T Bs[WG_MC][TK + WG_PAD];
  const int tilesK = p.K / TK;
  const int tap = blockIdx.x / tilesK, k0 = (blockIdx.x % tilesK) * TK;
  const int n0 = blockIdx.y * WG_TN;
  const int split = blockIdx.z;
  const int warp = threadIdx.x >> 5;
  constexpr int FR = TK == 64 ? 2 : 1;  // 16x16 fragments per warp along each axis
  const int wr = TK == 64 ? (warp >> 1) * 32 : warp * 16;
  const int wc = TK == 64 ? (warp & 1) * 32 : 0;
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[FR][FR];
#pragma unroll
  for (int i = 0; i < FR; ++i)
#pragma unroll
    for (int j = 0; j < FR; ++j) wmma::fill_fragment(acc[i][j], 0.0f);
  int dy = 0, dx = 0;
  if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
  const int m_begin = split * p.m_per_split, m_end = min(p.M, m_begin + p.m_per_split);
  for (int mc = m_begin; mc < m_end; mc += WG_MC) {
    // ---- stage the chunk: 16-byte loads, rows past M / outside the image are zero
    for (int i = threadIdx.x; i < WG_MC * (WG_TN / 8); i += 128) {
      const int r = i / (WG_TN / 8), c8 = (i % (WG_TN / 8)) * 8;
      const int m = mc + r;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (m < m_end) v = *reinterpret_cast<const uint4*>(p.a + size_t(m) * p.lda + n0 + c8);
      *reinterpret_cast<uint4*>(&As[r][c8]) = v;
    }
    for (int i = threadIdx.x; i < WG_MC * (TK / 8); i += 128) {
      const int r = i / (TK / 8), c8 = (i % (TK / 8)) * 8;
      const int m = mc + r;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (m < m_end) {
        long src = m;
        bool ok = true;
        if (p.taps == 9) {
          const int pix = p.H * p.W;
          const int img = m / pix, rem = m - img * pix, y = rem / p.W + dy, x = rem % p.W + dx;
          ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
          src = (long(img) * p.H + y) * p.W + x;
        }
        if (ok) v = *reinterpret_cast<const uint4*>(p.b + size_t(src) * p.ldb + k0 + c8);
      }
      *reinterpret_cast<uint4*>(&Bs[r][c8]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WG_MC; kk += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, T, wmma::col_major> fa[FR];   // dY^T: element (n, m) at As[m][n]
      wmma::fragment<wmma::matrix_b, 16, 16, 16, T, wmma::row_major> fb[FR];   // X: element (m, k) at Bs[m][k]
#pragma unroll
      for (int i = 0; i < FR; ++i) wmma::load_matrix_sync(fa[i], &As[kk][wr + 16 * i], WG_TN + WG_PAD);
#pragma unroll
      for (int j = 0; j < FR; ++j) wmma::load_matrix_sync(fb[j], &Bs[kk][wc + 16 * j], TK + WG_PAD);
#pragma unroll
      for (int i = 0; i < FR; ++i)
#pragma unroll
        for (int j = 0; j < FR; ++j) wmma::mma_sync(acc[i][j], fa[i], fb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dst = p.part + ((size_t(split) * p.taps + tap) * p.N + n0) * p.K + k0;
#pragma unroll
  for (int i = 0; i < FR; ++i)
#pragma unroll
    for (int j = 0; j < FR; ++j) wmma::store_matrix_sync(dst + size_t(wr + 16 * i) * p.K + wc + 16 * j, acc[i][j], p.K, wmma::mem_row_major);
}

// out[layout(n, k, tap)] = alpha * sum_s part[s][tap][n][k]
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int taps, int N, int K, float alpha, float* __restrict__ out,
                                    long so_n, long so_k, long so_tap) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const long i = long(blockIdx.x) * blockDim.x + threadIdx.x;
  const long per = long(taps) * N * K;
  if (i >= per) return;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += part[size_t(sp) * per + i];
  const int k = int(i % K);
  const int n = int((i / K) % N);
  const int tap = int(i / (long(K) * N));
  out[n * so_n + k * so_k + tap * so_tap] = alpha * s;
}


// ------------------------------------------------------------------------------------------------ LoRA factor gradients, fused
// Both factor gradients of one LoRA-wrapped linear y = (W + s B A) x in ONE kernel (+ one reduce):
//   U = X A^T  [M,16]     dB = s dY^T U  [N,16]            V = dY B  [M,16]     dA = s V^T X  [16,K]
// The path used to be two skinny tcgen05 GEMMs, two weight-gradient kernels and two reduces per linear: 768 launches of 5-10 us per
// training pass, 6.2 of its 32 ms.  Here a CTA owns a range of 64-row chunks of the pixel axis: per chunk it streams 64-column pieces
// of X and dY through shared memory twice -- first for the two skinny products (warps 0-3: U, warps 4-7: V, 16-bit results kept in
// shared memory), then for the two outer products, whose [N,16] / [K,16] fp32 accumulators live in registers across all chunks
// (KP / NP pieces per warp, compile-time; four pieces = 256 columns are staged per round so a 1280-wide layer needs 5 + 5 rounds).
// Every CTA writes its partial sums; lora_grad_reduce_kernel adds them in fixed order.
struct LoraGradParams {
  const uint16_t* x; int ldx;     // [M, K]
  const uint16_t* dy; int ldy;    // [M, N]
  const uint16_t* a16;            // [16, K] (lora_A, 16-bit)
  const uint16_t* bt16;           // [16, N] (lora_B^T, 16-bit)
  int M, chunks_per_cta;
  int groups;                     // gridDim.y: the outer-product rounds are dealt round-robin to `groups` CTAs per chunk range (small M: few chunk ranges)
  float* part_b;                  // [ctas][N][16]
  float* part_a;                  // [ctas][K][16]
};

constexpr int LG_W = 256;          // columns of X / dY staged per round (four 64-column pieces)
constexpr int LG_LD = LG_W + 8;
constexpr size_t LG_SMEM = size_t(2) * 64 * LG_LD * 2 + size_t(2) * 64 * 24 * 2 + size_t(8) * 16 * 20 * 4 + size_t(2) * 16 * LG_LD * 2;

template <typename T, int KP, int NP>
__global__ void __launch_bounds__(256, 1) lora_grad_kernel(const LoraGradParams p) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  constexpr int K = 64 * KP, N = 64 * NP, PMAX = KP > NP ? KP : NP, ROUNDS = (PMAX + 3) / 4;
  extern __shared__ __align__(128) unsigned char lg_smem[];
  T (*Xs)[LG_LD] = reinterpret_cast<T (*)[LG_LD]>(lg_smem);
  T (*Ys)[LG_LD] = reinterpret_cast<T (*)[LG_LD]>(lg_smem + size_t(64) * LG_LD * 2);
  T (*UVs)[64][24] = reinterpret_cast<T (*)[64][24]>(lg_smem + size_t(2) * 64 * LG_LD * 2);       // [0] = U, [1] = V (16-bit)
  float (*Fs)[16][20] = reinterpret_cast<float (*)[16][20]>(lg_smem + size_t(2) * 64 * LG_LD * 2 + size_t(2) * 64 * 24 * 2);  // per-warp fp32 staging
  T (*Fac)[16][LG_LD] = reinterpret_cast<T (*)[16][LG_LD]>(lg_smem + size_t(2) * 64 * LG_LD * 2 + size_t(2) * 64 * 24 * 2 + size_t(8) * 16 * 20 * 4);  // [0] = A, [1] = B^T: 256 columns
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = warp >> 2, t = warp & 3;  // role 0: U / dB (outer product with dY), role 1: V / dA (outer product with X)
  constexpr int PO = 4 * ROUNDS;             // outer-product pieces a warp owns (N or K up to 1280: 20)
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[PO];
#pragma unroll
  for (int i = 0; i < PO; ++i) wmma::fill_fragment(acc[i], 0.0f);
  const int c_begin = blockIdx.x * p.chunks_per_cta;
  const int n_chunks = (p.M + 63) / 64;
  const int c_end = min(n_chunks, c_begin + p.chunks_per_cta);

  auto stage = [&](int m0, int rd, bool with_factors) {  // columns [256 rd, 256 rd + 256) of X and dY (64 rows); rows past M / columns past K, N are zero
    if (with_factors) {  // the same columns of A [16,K] and B^T [16,N] (fragment loads from global memory are scattered 2-byte accesses)
      for (int i = threadIdx.x; i < 2 * 16 * (LG_W / 8); i += 256) {
        const int which = i / (16 * (LG_W / 8)), j = i % (16 * (LG_W / 8));
        const int r = j / (LG_W / 8), c8 = (j % (LG_W / 8)) * 8, col = rd * LG_W + c8;
        const int ld = which == 0 ? K : N;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (col < ld) v = *reinterpret_cast<const uint4*>((which == 0 ? p.a16 : p.bt16) + size_t(r) * ld + col);
        *reinterpret_cast<uint4*>(&Fac[which][r][c8]) = v;
      }
    }
    for (int i = threadIdx.x; i < 64 * (LG_W / 8); i += 256) {
      const int r = i / (LG_W / 8), c8 = (i % (LG_W / 8)) * 8;
      const int m = m0 + r, col = rd * LG_W + c8;
      uint4 vx = make_uint4(0u, 0u, 0u, 0u), vy = vx;
      if (m < p.M) {
        if (col < K) vx = *reinterpret_cast<const uint4*>(p.x + size_t(m) * p.ldx + col);
        if (col < N) vy = *reinterpret_cast<const uint4*>(p.dy + size_t(m) * p.ldy + col);
      }
      *reinterpret_cast<uint4*>(&Xs[r][c8]) = vx;
      *reinterpret_cast<uint4*>(&Ys[r][c8]) = vy;
    }
  };

  for (int ch = c_begin; ch < c_end; ++ch) {
    const int m0 = ch * 64;
    // ---- phase 1: U = X A^T (warps 0-3), V = dY B (warps 4-7); warp t owns rows [16t, 16t+16)
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> sk;
    wmma::fill_fragment(sk, 0.0f);
#pragma unroll 1
    for (int rd = 0; rd < ROUNDS; ++rd) {
      __syncthreads();
      stage(m0, rd, true);
      __syncthreads();
      const T* src = role == 0 ? &Xs[16 * t][0] : &Ys[16 * t][0];
      const int ldf = role == 0 ? K : N;
      const T* fac = &Fac[role][0][0];  // [16][256 columns of K or N]: element (k, r) at fac[r * LG_LD + k]
      const int kend = min(LG_W, ldf - rd * LG_W);
      for (int kk = 0; kk < kend; kk += 16) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, T, wmma::row_major> fa;
        wmma::fragment<wmma::matrix_b, 16, 16, 16, T, wmma::col_major> fb;
        wmma::load_matrix_sync(fa, src + kk, LG_LD);
        wmma::load_matrix_sync(fb, fac + kk, LG_LD);
        wmma::mma_sync(sk, fa, fb, sk);
      }
    }
    wmma::store_matrix_sync(&Fs[warp][0][0], sk, 20, wmma::mem_row_major);
    __syncwarp();
    for (int i = lane; i < 256; i += 32) UVs[role][16 * t + (i >> 4)][i & 15] = from_float_t<T>(Fs[warp][i >> 4][i & 15]);
    // ---- phase 2: dB += dY^T U (warps 0-3), dA^T += X^T V (warps 4-7); warp t owns output rows [64 pc + 16 t, +16) of every piece pc
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      if (rd % p.groups != int(blockIdx.y)) continue;  // (uniform per CTA) another CTA of this chunk range owns this round's output pieces
      __syncthreads();  // (also publishes U / V on the first pass)
      stage(m0, rd, false);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pc = 4 * rd + j;
        if (pc < (role == 0 ? NP : KP)) {
          const T* src = role == 0 ? &Ys[0][64 * j + 16 * t] : &Xs[0][64 * j + 16 * t];  // transposed operand: element (n, m) at src[m * LG_LD + n]
#pragma unroll
          for (int kk = 0; kk < 64; kk += 16) {
            wmma::fragment<wmma::matrix_a, 16, 16, 16, T, wmma::col_major> fa;
            wmma::fragment<wmma::matrix_b, 16, 16, 16, T, wmma::row_major> fb;
            wmma::load_matrix_sync(fa, src + kk * LG_LD, LG_LD);
            wmma::load_matrix_sync(fb, &UVs[role][kk][0], 24);
            wmma::mma_sync(acc[pc], fa, fb, acc[pc]);
          }
        }
      }
    }
  }
  float* dst = (role == 0 ? p.part_b + size_t(blockIdx.x) * N * 16 : p.part_a + size_t(blockIdx.x) * K * 16) + size_t(16 * t) * 16;
#pragma unroll
  for (int pc = 0; pc < PO; ++pc)
    if (pc < (role == 0 ? NP : KP) && (pc / 4) % p.groups == int(blockIdx.y)) wmma::store_matrix_sync(dst + size_t(pc) * 64 * 16, acc[pc], 16, wmma::mem_row_major);
}

// gB[n*16 + r] = alpha * sum_c part_b[c][n][r]   ;   gA[r*K + k] = alpha * sum_c part_a[c][k][r]
__global__ void lora_grad_reduce_kernel(const float* __restrict__ part_b, const float* __restrict__ part_a, int ctas, int N, int K, float alpha,
                                        float* __restrict__ gB, float* __restrict__ gA) {
  pdl_trigger();  // programmatic dependent launch (launch.cuh): no global access before pdl_wait()
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * 16) {
    if (!gB) return;
    float s = 0.f;
    for (int c = 0; c < ctas; ++c) s += part_b[size_t(c) * N * 16 + i];
    gB[i] = alpha * s;
  } else if (i < (N + K) * 16) {
    if (!gA) return;
    const int j = i - N * 16, k = j >> 4, r = j & 15;
    float s = 0.f;
    for (int c = 0; c < ctas; ++c) s += part_a[size_t(c) * K * 16 + j];
    gA[size_t(r) * K + k] = alpha * s;
  }
}

template <typename T, int KP, int NP>
const char* lora_grad_launch_kn(const LoraGradParams& p, int ctas, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(lora_grad_kernel<T, KP, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LG_SMEM)) != cudaSuccess)
      return "lora_grads: cudaFuncSetAttribute failed";
    attr = true;
  }
  launch_k(lora_grad_kernel<T, KP, NP>, dim3(dim3(ctas, p.groups)), dim3(256), LG_SMEM, st, p);
  return nullptr;
}
template <typename T, int KP>
const char* lora_grad_launch_n(const LoraGradParams& p, int NPn, int ctas, cudaStream_t st) {
  switch (NPn) {
    case 5: return lora_grad_launch_kn<T, KP, 5>(p, ctas, st);
    case 10: return lora_grad_launch_kn<T, KP, 10>(p, ctas, st);
    case 20: return lora_grad_launch_kn<T, KP, 20>(p, ctas, st);
  }
  return "lora_grads: N must be 320, 640 or 1280";
}
template <typename T>
const char* lora_grad_launch(const LoraGradParams& p, int KPn, int NPn, int ctas, cudaStream_t st) {
  switch (KPn) {
    case 5: return lora_grad_launch_n<T, 5>(p, NPn, ctas, st);
    case 10: return lora_grad_launch_n<T, 10>(p, NPn, ctas, st);
    case 12: return lora_grad_launch_n<T, 12>(p, NPn, ctas, st);
    case 20: return lora_grad_launch_n<T, 20>(p, NPn, ctas, st);
  }
  return "lora_grads: K must be 320, 640, 768 or 1280";
}

}  // namespace

int wgrad_splits(int M, int N, int K, int taps) {
  const int tk = K % 64 == 0 ? 64 : 16;
  const long tiles = long(taps) * (K / tk) * (N / WG_TN);
  const int chunks = (M + WG_MC - 1) / WG_MC;
  long s = (296 + tiles - 1) / tiles;
  if (s > 32) s = 32;
  if (s > chunks) s = chunks;
  return s < 1 ? 1 : int(s);
}
size_t wgrad_scratch_floats(int M, int N, int K, int taps) { return size_t(wgrad_splits(M, N, K, taps)) * taps * N * K; }

const char* wgrad(const void* dy16, int lda, const void* x16, int ldb, int M, int N, int K, int taps, int Bimg, int H, int W, float alpha,
                  float* out, long so_n, long so_k, long so_tap, float* scratch, int fp16, cudaStream_t st) {
  if (N % WG_TN != 0 || !(K % 64 == 0 || K == 16) || (taps != 1 && taps != 9)) return "wgrad: N must be a multiple of 64, K a multiple of 64 or 16, taps 1 or 9";
  if (lda % 8 != 0 || ldb % 8 != 0 || (reinterpret_cast<uintptr_t>(dy16) & 15) || (reinterpret_cast<uintptr_t>(x16) & 15))
    return "wgrad: operands must be 16-byte aligned with pitches that are multiples of 8";
  if (taps == 9 && long(Bimg) * H * W != M) return "wgrad: conv mode needs M = B*H*W";
  WgradParams p;
  p.a = static_cast<const uint16_t*>(dy16); p.lda = lda; p.b = static_cast<const uint16_t*>(x16); p.ldb = ldb;
  p.M = M; p.N = N; p.K = K; p.taps = taps; p.H = H; p.W = W;
  const int splits = wgrad_splits(M, N, K, taps);
  const int chunks = (M + WG_MC - 1) / WG_MC;
  p.m_per_split = ((chunks + splits - 1) / splits) * WG_MC;
  p.part = scratch;
  const int tk = K % 64 == 0 ? 64 : 16;
  const dim3 grid(unsigned(taps * (K / tk)), unsigned(N / WG_TN), unsigned(splits));
  if (fp16) {
    if (tk == 64) launch_k(wgrad_kernel<__half, 64>, dim3(grid), dim3(128), 0, st, p); else launch_k(wgrad_kernel<__half, 16>, dim3(grid), dim3(128), 0, st, p);
  } else {
    if (tk == 64) launch_k(wgrad_kernel<__nv_bfloat16, 64>, dim3(grid), dim3(128), 0, st, p); else launch_k(wgrad_kernel<__nv_bfloat16, 16>, dim3(grid), dim3(128), 0, st, p);
  }
  const long per = long(taps) * N * K;
  launch_k(wgrad_reduce_kernel, dim3(unsigned((per + 255) / 256)), dim3(256), 0, st, scratch, splits, taps, N, K, alpha, out, so_n, so_k, so_tap);
  return cudaGetLastError() == cudaSuccess ? nullptr : "wgrad launch failed";
}


// ---- fused LoRA factor gradients (rank 16): gB [N,16] = alpha dY^T (X A^T), gA [16,K] = alpha (dY B)^T X; either output may be null
static int lora_grad_ctas(int M) {
  const int chunks = (M + 63) / 64;
  return chunks < 148 ? chunks : 148;
}
bool lora_grads_supported(int N, int K) { return (N == 320 || N == 640 || N == 1280) && (K == 320 || K == 640 || K == 768 || K == 1280); }
size_t lora_grads_scratch_floats(int M, int N, int K) { return size_t(lora_grad_ctas(M)) * 16 * (size_t(N) + K); }
const char* lora_grads(const void* x16, int ldx, const void* dy16, int ldy, const void* a16, const void* bt16, int M, int N, int K, float alpha, float* gA,
                       float* gB, float* scratch, int fp16, cudaStream_t st) {
  if (!lora_grads_supported(N, K)) return "lora_grads: unsupported N / K";
  if (ldx % 8 != 0 || ldy % 8 != 0 || ((reinterpret_cast<uintptr_t>(x16) | reinterpret_cast<uintptr_t>(dy16) | reinterpret_cast<uintptr_t>(a16) |
                                         reinterpret_cast<uintptr_t>(bt16)) & 31))
    return "lora_grads: operands must be 32-byte aligned with pitches that are multiples of 8";
  LoraGradParams p;
  p.x = static_cast<const uint16_t*>(x16); p.ldx = ldx; p.dy = static_cast<const uint16_t*>(dy16); p.ldy = ldy;
  p.a16 = static_cast<const uint16_t*>(a16); p.bt16 = static_cast<const uint16_t*>(bt16);
  p.M = M;
  const int ctas = lora_grad_ctas(M), chunks = (M + 63) / 64;
  p.chunks_per_cta = (chunks + ctas - 1) / ctas;
  p.part_b = scratch; p.part_a = scratch + size_t(ctas) * N * 16;
  const int used = (chunks + p.chunks_per_cta - 1) / p.chunks_per_cta;  // CTAs that own at least one chunk (the others would write zeros)
  {  // few chunk ranges (16x16 / 8x8 levels, the 77-token context): split the outer-product rounds over up to 5 CTAs per range
    const int rounds = ((K > N ? K : N) / 64 + 3) / 4;
    int g = 148 / used;
    if (g > rounds) g = rounds;
    p.groups = g < 1 ? 1 : g;
  }
  if (const char* e = fp16 ? lora_grad_launch<__half>(p, K / 64, N / 64, used, st) : lora_grad_launch<__nv_bfloat16>(p, K / 64, N / 64, used, st)) return e;
  launch_k(lora_grad_reduce_kernel, dim3(((N + K) * 16 + 255) / 256), dim3(256), 0, st, p.part_b, p.part_a, used, N, K, alpha, gB, gA);
  return cudaGetLastError() == cudaSuccess ? nullptr : "lora_grads launch failed";
}

}  // namespace madm

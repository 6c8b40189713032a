// Weight gradients of the LoRA training step's trainable set: the contraction runs over the PIXEL axis,
//   dW[n, k] = alpha * sum_m dY[m, n] * X[m, k]
// with both operands stored pixel-major (NHWC rows), i.e. a "TN" GEMM.  Used for the LoRA factors (skinny: k = rank 16), the GN-bottleneck
// projections' 1x1 convs and, with an implicit im2col of X (zero padding = bounds check on the shifted pixel), their 3x3 conv.
// These are small next to the dgrad GEMMs (LoRA + projections only: the base UNet weights are frozen), so the kernel is a plain
// warp-level tensor-core kernel (wmma 16x16x16, fp32 accumulate) with a deterministic split over M: every split writes its own fp32
// partial tile, a second kernel reduces them in fixed order, scales, and scatters into the parameter's PyTorch layout.
#include "kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mma.h>

namespace madm {

namespace {

using namespace nvcuda;

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
    if (tk == 64) wgrad_kernel<__half, 64><<<grid, 128, 0, st>>>(p); else wgrad_kernel<__half, 16><<<grid, 128, 0, st>>>(p);
  } else {
    if (tk == 64) wgrad_kernel<__nv_bfloat16, 64><<<grid, 128, 0, st>>>(p); else wgrad_kernel<__nv_bfloat16, 16><<<grid, 128, 0, st>>>(p);
  }
  const long per = long(taps) * N * K;
  wgrad_reduce_kernel<<<unsigned((per + 255) / 256), 256, 0, st>>>(scratch, splits, taps, N, K, alpha, out, so_n, so_k, so_tap);
  return cudaGetLastError() == cudaSuccess ? nullptr : "wgrad launch failed";
}

}  // namespace madm

// Micro-benchmark: how fast can ONE warp per scheduler stream `p = 2^(s*c - m)` -> 16-bit pack, as the attention softmax does?
// Variants: A = ex2 + pack in natural order (ptxas interleaves the F2FP two MUFUs behind), B = ex2 only (sum), C = packs forced
// a whole 16-element block behind their MUFUs through a real (free) data dependency, D = f16x2 ex2 (no F2FP after the MUFU).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_stream mufu_stream.cu ; run: ./mufu_stream
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int V>
__global__ void k(const float* __restrict__ in, uint32_t* __restrict__ out, long long* __restrict__ clk, int iters, float sl, float ms) {
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = in[(threadIdx.x * 32 + i) & 1023];
  uint32_t acc = 0;
  float facc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if constexpr (V == 0) {
      float e[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) e[i] = ex2(fmaf(s[i], sl, ms));
#pragma unroll
      for (int i = 0; i < 32; i += 2) acc ^= pack(e[i], e[i + 1]);
    } else if constexpr (V == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) facc += ex2(fmaf(s[i], sl, ms));
    } else if constexpr (V == 2) {
      float e0[16], e1[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) e0[i] = ex2(fmaf(s[i], sl, ms));
#pragma unroll
      for (int i = 0; i < 16; ++i) e1[i] = ex2(fmaf(s[16 + i], sl, ms));
      // packs of block 0 depend (for free: + 0 * x, x finite) on the LAST ex2 of block 1; block 1's on a value of the next iteration's inputs
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(fmaf(e1[15], 0.0f, e0[i]), e0[i + 1]);
#pragma unroll
      for (int i = 0; i < 16; i += 2) acc ^= pack(e1[i], e1[i + 1]);
    } else if constexpr (V == 3) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) acc ^= ex2h2(pack(fmaf(s[i], sl, ms), fmaf(s[i + 1], sl, ms)));
    } else if constexpr (V == 4) {  // all 32 MUFUs first, then all packs behind one dependency on the last
      float e[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) e[i] = ex2(fmaf(s[i], sl, ms));
#pragma unroll
      for (int i = 0; i < 32; i += 2) acc ^= pack(fmaf(e[31], 0.0f, e[i]), e[i + 1]);
    }
    ms += 1e-7f;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(facc);
}

int main() {
  float* in; uint32_t* out; long long* clk;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = -float(i % 37) * 0.3f;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  const int iters = 2000;
  const char* names[5] = {"A ex2+pack natural order", "B ex2 only", "C packs one block behind (fma dep)", "D f16x2 ex2", "E all packs behind the last ex2"};
  for (int threads : {128, 256, 512}) {
    for (int v = 0; v < 5; ++v) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (v) {
          case 0: k<0><<<148, threads>>>(in, out, clk, iters, 0.2f, -0.1f); break;
          case 1: k<1><<<148, threads>>>(in, out, clk, iters, 0.2f, -0.1f); break;
          case 2: k<2><<<148, threads>>>(in, out, clk, iters, 0.2f, -0.1f); break;
          case 3: k<3><<<148, threads>>>(in, out, clk, iters, 0.2f, -0.1f); break;
          case 4: k<4><<<148, threads>>>(in, out, clk, iters, 0.2f, -0.1f); break;
        }
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      printf("%d warps/SMSP  %-40s %7.1f clk per 32 elements per warp = %5.2f clk per ex2 (per SMSP: %5.2f)\n", threads / 128, names[v], double(c) / iters,
             double(c) / iters / 32, double(c) / iters / 32 / (threads / 128));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

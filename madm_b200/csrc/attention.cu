// Fused flash-style attention for the SD-1.4 UNet transformer blocks (self: n in {4096,1024,256,64}; cross: 77 keys),
// 8 heads of d in {40, 80, 160}.  One CTA = 64 query rows of one (image, head); 4 warps x 16 rows.  K/V tiles of 64 keys
// are double-buffered in shared memory with cp.async; S = QK^T and O += PV run on the warp-level tensor-core path
// (mma.sync m16n8k16 bf16, fp32 accumulate) with an online softmax in registers (exp2, fp32 running max/sum).
// Round-1 implementation: the tcgen05/TMEM version (S and O accumulators in TMEM) is the planned replacement.
#include "cvt.cuh"
#include "kernels.h"

namespace madm {

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
template <bool FP16>
__device__ __forceinline__ void mma_16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (FP16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool FP16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  return pack2_16(a, b, FP16 ? 1 : 0);
}

template <int D>
struct AttnCfg {
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int RS = DP * 2 + 16;  // smem row pitch in bytes (conflict-free ldmatrix)
  static constexpr int TILE = 64 * RS;
  static constexpr int SMEM = 5 * TILE;   // Q + 2xK + 2xV
  static constexpr int CHUNKS = D * 2 / 16;
};

template <int D>
__device__ __forceinline__ void load_tile(uint8_t* smem, const uint16_t* g, int ld, int row0, int nrows_total, int tid) {
  using C = AttnCfg<D>;
  for (int i = tid; i < 64 * C::CHUNKS; i += 128) {
    const int r = i / C::CHUNKS;
    const int ch = i - r * C::CHUNKS;
    const int gr = row0 + r;
    const bool ok = gr < nrows_total;
    const uint16_t* src = g + size_t(ok ? gr : 0) * ld + ch * 8;
    cp_async16(s_u32(smem + r * C::RS + ch * 16), src, ok ? 16 : 0);
  }
}

template <int D, bool FP16>
__global__ void __launch_bounds__(128) flash_attn_kernel(const uint16_t* __restrict__ Q, int ldq, const uint16_t* __restrict__ K,
                                                         int ldk, const uint16_t* __restrict__ V, int ldv,
                                                         uint16_t* __restrict__ O, int ldo, int Nq, int Nk, long q_bs, long kv_bs,
                                                         long o_bs, float scale_log2) {
  using C = AttnCfg<D>;
  constexpr int DP = C::DP;
  constexpr int KS = DP / 16;  // k-steps over head dim for QK^T
  constexpr int NT = DP / 8;   // n-tiles over head dim for PV
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + C::TILE;
  uint8_t* sV = smem + 3 * C::TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
  const uint16_t* Qg = Q + size_t(b) * q_bs + h * D;
  const uint16_t* Kg = K + size_t(b) * kv_bs + h * D;
  const uint16_t* Vg = V + size_t(b) * kv_bs + h * D;

  // zero the padded head-dim columns once (cp.async never writes them)
  if (DP > D) {
    for (int i = tid; i < 5 * 64; i += 128) {
      uint8_t* rowp = smem + size_t(i) * C::RS + D * 2;
      for (int j = 0; j < (DP - D) * 2; j += 4) *reinterpret_cast<uint32_t*>(rowp + j) = 0u;
    }
  }
  __syncthreads();

  load_tile<D>(sQ, Qg, ldq, q0, Nq, tid);
  load_tile<D>(sK, Kg, ldk, 0, Nk, tid);
  load_tile<D>(sV, Vg, ldv, 0, Nk, tid);
  cp_async_commit();

  const int ntiles = (Nk + 63) / 64;
  uint32_t qf[KS][4];
  float o[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int kt = 0; kt < ntiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ntiles) {
      load_tile<D>(sK + (buf ^ 1) * C::TILE, Kg, ldk, (kt + 1) * 64, Nk, tid);
      load_tile<D>(sV + (buf ^ 1) * C::TILE, Vg, ldv, (kt + 1) * 64, Nk, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
      const uint32_t qbase = s_u32(sQ) + (warp * 16 + (lane & 15)) * C::RS + (lane >> 4) * 16;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) ldsm_x4(qbase + ks * 32, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    const uint32_t kb = s_u32(sK + buf * C::TILE);
    const uint32_t vb = s_u32(sV + buf * C::TILE);

    // ---- S = Q K^T : 16 x 64 per warp
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
        uint32_t b0, b1, b2, b3;
        const uint32_t addr = kb + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * C::RS + ks * 32 + ((lane >> 3) & 1) * 16;
        ldsm_x4(addr, b0, b1, b2, b3);
        mma_16<FP16>(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_16<FP16>(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    // ---- online softmax (rows lane/4 and lane/4+8): max on the raw scores, then p = 2^(s*c - m*c) as one FFMA + MUFU
    if (kt == ntiles - 1 && (Nk & 63) != 0) {  // only the last, ragged key tile needs masking
      const int key_base = kt * 64 + 2 * (lane & 3);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (key_base + i * 8 + (e & 1) >= Nk) s[i][e] = -INFINITY;
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f}, ms[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = ex2((m_run[r] - mx[r]) * scale_log2);
      m_run[r] = mx[r];
      ms[r] = -mx[r] * scale_log2;
    }
    uint32_t pf[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = ex2(fmaf(s[i][0], scale_log2, ms[0])), p1 = ex2(fmaf(s[i][1], scale_log2, ms[0]));
      const float p2 = ex2(fmaf(s[i][2], scale_log2, ms[1])), p3 = ex2(fmaf(s[i][3], scale_log2, ms[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[i][0] = pack2<FP16>(p0, p1);
      pf[i][1] = pack2<FP16>(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
      const uint32_t a0 = pf[2 * kk][0], a1 = pf[2 * kk][1], a2 = pf[2 * kk + 1][0], a3 = pf[2 * kk + 1][1];
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {  // pairs of 8-wide head-dim tiles
        uint32_t b0, b1, b2, b3;
        const uint32_t addr = vb + (kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3)) * C::RS + np * 32 + (lane >> 4) * 16;
        ldsm_x4_t(addr, b0, b1, b2, b3);
        mma_16<FP16>(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_16<FP16>(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }

  // ---- normalise and store
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int r0 = q0 + warp * 16 + (lane >> 2);
  uint16_t* Og = O + size_t(b) * o_bs + h * D;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int col = i * 8 + 2 * (lane & 3);
    if (col < D) {
      if (r0 < Nq) *reinterpret_cast<uint32_t*>(Og + size_t(r0) * ldo + col) = pack2<FP16>(o[i][0] * inv0, o[i][1] * inv0);
      if (r0 + 8 < Nq) *reinterpret_cast<uint32_t*>(Og + size_t(r0 + 8) * ldo + col) = pack2<FP16>(o[i][2] * inv1, o[i][3] * inv1);
    }
  }
}

template <int D, bool FP16>
static const char* launch_attn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int B,
                               int heads, int Nq, int Nk, long q_bs, long kv_bs, long o_bs, float scale, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(flash_attn_kernel<D, FP16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<D>::SMEM) != cudaSuccess)
      return "attention: cudaFuncSetAttribute failed";
    attr = true;
  }
  dim3 grid((Nq + 63) / 64, heads, B);
  flash_attn_kernel<D, FP16><<<grid, 128, AttnCfg<D>::SMEM, st>>>(
      reinterpret_cast<const uint16_t*>(q), ldq, reinterpret_cast<const uint16_t*>(k), ldk,
      reinterpret_cast<const uint16_t*>(v), ldv, reinterpret_cast<uint16_t*>(o), ldo, Nq, Nk, q_bs, kv_bs, o_bs,
      scale * 1.4426950408889634f);
  return cudaGetLastError() == cudaSuccess ? nullptr : "attention: launch failed";
}

const char* flash_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int B,
                            int heads, int d, int Nq, int Nk, long q_bs, long kv_bs, long o_bs, float scale, int fp16, cudaStream_t st) {
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 2) return "attention: row pitches must be multiples of 8 elements";
  if (Nk < 1 || Nq < 1) return "attention: empty problem";
  switch (d) {
    case 40: return fp16 ? launch_attn<40, true>(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, q_bs, kv_bs, o_bs, scale, st) : launch_attn<40, false>(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, q_bs, kv_bs, o_bs, scale, st);
    case 80: return fp16 ? launch_attn<80, true>(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, q_bs, kv_bs, o_bs, scale, st) : launch_attn<80, false>(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, q_bs, kv_bs, o_bs, scale, st);
    case 160: return fp16 ? launch_attn<160, true>(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, q_bs, kv_bs, o_bs, scale, st) : launch_attn<160, false>(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, q_bs, kv_bs, o_bs, scale, st);
  }
  return "attention: unsupported head dim (40, 80, 160)";
}

}  // namespace madm

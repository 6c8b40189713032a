// tcgen05/TMEM implicit-GEMM for sm_100a: every dense contraction of the MADM feature-extraction path
// (ResBlock conv3x3/1x1, up/down-sample convs, attention q/k/v/out projections with the LoRA update already
// folded into the packed weight, GEGLU feed-forward, VAE convs, feature projections) runs through this kernel.
//
//   D[M,N] = act( alpha * sum_seg sum_tap A_seg(tap)[M,C] * W[N, k-slice]^T + bias[N] + rowbias[img(m),N] + residual[M,N] )
//
// A is never materialised as an im2col matrix: each 64-channel K chunk of each filter tap is one 4-D TMA box over the
// NHWC bf16 activation tensor, shifted by (dx,dy); out-of-bounds rows/columns are zero-filled by TMA, which is exactly
// the convolution's zero padding.  W tiles are 2-D TMA boxes over the packed K-major weight matrix.  Both land in
// 128B-swizzled shared memory and are consumed directly by tcgen05.mma (UMMA 128xBNx16, fp32 accumulators in TMEM).
// Warp roles: warp0 = TMA producer, warp1 = TMEM allocator + single-thread MMA issuer, warps2-5 = epilogue
// (tcgen05.ld -> registers -> fused bias / time-embedding row bias / fp32 residual / SiLU / GEGLU -> fp32 and/or bf16 stores).
// Two CTAs are resident per SM (<=113 KB smem, <=256 TMEM columns each) so one CTA's epilogue overlaps the other's mainloop.
#include "gemm_tc.h"
#include "cvt.cuh"
#include "ptx.cuh"

#include <mutex>
#include <stdio.h>

namespace madm {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int kThreads = 192;
static constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

struct GemmParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  int nseg;
  int kchunks[2];
  int cpt[2];
  int8_t dx[2][9];
  int8_t dy[2][9];
  int boff[2][9];
  int M, N;
  int W, pix;  // output grid width and pixels per image
  int box_w, box_h, box_b;
  const float* bias;
  const float* rowbias;
  int rows_per_img;
  int ld_rowbias;
  const float* residual;
  int ldr;
  float* out_f32;
  int ldo32;
  __nv_bfloat16* out_bf16;
  int ldo16;
  int act;
  float alpha;
  int vec_ok;
  int n_tiles;
  int fp16;
};

template <int BN>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // keep <= ~110 KB so two CTAs fit in one SM's 227 KB
  static constexpr int STAGES = (110 * 1024 / STAGE_BYTES) > 6 ? 6 : (110 * 1024 / STAGE_BYTES);
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr size_t SMEM = size_t(STAGES) * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }
__device__ __forceinline__ float gelu_erf_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }


template <int BN>
__global__ void __launch_bounds__(kThreads) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + C::STAGES * A_STAGE_BYTES;
  const uint32_t sBar = sB + C::STAGES * C::B_STAGE_BYTES;  // full[STAGES], empty[STAGES], tmem_full, tmem_addr
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (C::STAGES + s); };
  const uint32_t tmem_full_bar = sBar + 8u * (2 * C::STAGES);
  const uint32_t tmem_slot = sBar + 8u * (2 * C::STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x % p.n_tiles;
  const int tile_m = blockIdx.x / p.n_tiles;
  const int m0 = tile_m * BM;
  const int n0 = tile_n * BN;
  const int total_chunks = p.kchunks[0] + (p.nseg > 1 ? p.kchunks[1] : 0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA[0]);
    if (p.nseg > 1) prefetch_tmap(&p.tmA[1]);
    prefetch_tmap(&p.tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int x0, y0, b0;
      if (p.pix >= BM) {
        b0 = m0 / p.pix;
        const int rem = m0 - b0 * p.pix;
        y0 = rem / p.W;
        x0 = rem - y0 * p.W;
      } else {
        b0 = m0 / p.pix;
        y0 = 0;
        x0 = 0;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < total_chunks; ++kc) {
        const int seg = (kc < p.kchunks[0]) ? 0 : 1;
        const int lk = seg ? kc - p.kchunks[0] : kc;
        const int tap = lk / p.cpt[seg];
        const int cc = lk - tap * p.cpt[seg];
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
        tma_load_4d(sA + stage * A_STAGE_BYTES, &p.tmA[seg], full_bar(stage), cc * BK, x0 + p.dx[seg][tap],
                    y0 + p.dy[seg][tap], b0 + p.boff[seg][tap]);
        tma_load_2d(sB + stage * C::B_STAGE_BYTES, &p.tmB, full_bar(stage), kc * BK, n0);
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_16(BM, BN, p.fp16);
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = 0; kc < total_chunks; ++kc) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(sA + stage * A_STAGE_BYTES);
        const uint64_t bdesc = make_smem_desc_sw128(sB + stage * C::B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 32 B (16 bf16) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
          umma_bf16_ss(tmem_base, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kc | k) != 0);
        }
        umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
    const float* rb = nullptr;
    if (p.rowbias != nullptr && row_ok) rb = p.rowbias + size_t(m / p.rows_per_img) * p.ld_rowbias;
    const float alpha = p.alpha;
    uint32_t r[32];

    if (p.act == ACT_GEGLU) {
      // weight rows are tile-interleaved: columns [0,64) of this tile = h, [64,128) = gate, for output cols [64*tile_n, +64)
      if constexpr (BN == 128) {
        uint32_t g[32];
        const int on0 = tile_n * 64;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          __syncwarp();
          tmem_ld32(taddr + half * 32, r);
          tmem_ld32(taddr + 64 + half * 32, g);
          tmem_ld_wait();
          if (row_ok) {
            __nv_bfloat16* o16 = p.out_bf16 + size_t(m) * p.ldo16 + on0 + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float v[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                float h = __uint_as_float(r[j + t]) * alpha;
                float gt = __uint_as_float(g[j + t]) * alpha;
                if (p.bias) {
                  h += __ldg(p.bias + n0 + half * 32 + j + t);
                  gt += __ldg(p.bias + n0 + 64 + half * 32 + j + t);
                }
                v[t] = h * gelu_erf_f(gt);
              }
              uint4 pk;
              pk.x = pack2_16(v[0], v[1], p.fp16);
              pk.y = pack2_16(v[2], v[3], p.fp16);
              pk.z = pack2_16(v[4], v[5], p.fp16);
              pk.w = pack2_16(v[6], v[7], p.fp16);
              *reinterpret_cast<uint4*>(o16 + j) = pk;
            }
          }
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        __syncwarp();
        if constexpr (BN % 32 != 0) {
          tmem_ld16(taddr + c, r);
        } else {
          tmem_ld32(taddr + c, r);
        }
        tmem_ld_wait();
        constexpr int CW = (BN % 32 != 0) ? 16 : 32;
        const int nb = n0 + c;
        if (!row_ok) {
          // nothing to store for padded rows; fall through to the warp-converged loop head
        } else if (p.vec_ok && nb + CW <= p.N) {
#pragma unroll
          for (int j = 0; j < CW; j += 8) {
            const int n = nb + j;
            float v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[j + t]) * alpha;
            if (p.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (rb) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(rb + n + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (p.residual) {
              const float* rp = p.residual + size_t(m) * p.ldr + n;
              const float4 b0 = *reinterpret_cast<const float4*>(rp);
              const float4 b1 = *reinterpret_cast<const float4*>(rp + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (p.act == ACT_SILU) {
#pragma unroll
              for (int t = 0; t < 8; ++t) v[t] = silu_f(v[t]);
            } else if (p.act == ACT_RELU) {
#pragma unroll
              for (int t = 0; t < 8; ++t) v[t] = fmaxf(v[t], 0.0f);
            }
            if (p.out_f32) {
              float* op = p.out_f32 + size_t(m) * p.ldo32 + n;
              *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
            if (p.out_bf16) {
              uint4 pk;
              pk.x = pack2_16(v[0], v[1], p.fp16);
              pk.y = pack2_16(v[2], v[3], p.fp16);
              pk.z = pack2_16(v[4], v[5], p.fp16);
              pk.w = pack2_16(v[6], v[7], p.fp16);
              *reinterpret_cast<uint4*>(p.out_bf16 + size_t(m) * p.ldo16 + n) = pk;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const int n = nb + j;
            if (n >= p.N) continue;
            float v = __uint_as_float(r[j]) * alpha;
            if (p.bias) v += __ldg(p.bias + n);
            if (rb) v += __ldg(rb + n);
            if (p.residual) v += p.residual[size_t(m) * p.ldr + n];
            if (p.act == ACT_SILU) v = silu_f(v);
            else if (p.act == ACT_RELU) v = fmaxf(v, 0.0f);
            if (p.out_f32) p.out_f32[size_t(m) * p.ldo32 + n] = v;
            if (p.out_bf16) reinterpret_cast<uint16_t*>(p.out_bf16)[size_t(m) * p.ldo16 + n] = cvt_16(v, p.fp16);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static const char* encode_map(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                              const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return "cuTensorMapEncodeTiled unavailable (no CUDA driver)";
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static thread_local char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dims %llu/%llu)", int(r), rank,
             (unsigned long long)dims[0], (unsigned long long)dims[1]);
    return buf;
  }
  return nullptr;
}

static int pick_bn(const GemmDesc& d) {
  if (d.act == ACT_GEGLU) return 128;
  const int N = d.N;
  if (N <= 16) return 16;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N % 128 == 0) return 128;
  if (N % 160 == 0) return 160;
  if (N % 192 == 0) return 192;
  if (N % 64 == 0 && N < 512) return 64;
  return 128;
}

const char* gemm_prepare(const GemmDesc& d, GemmLaunch* out) {
  GemmLaunch& L = *out;
  L.d = d;
  if (d.nseg < 1 || d.nseg > 2) return "gemm: nseg must be 1 or 2";
  if (d.M <= 0 || d.N <= 0) return "gemm: empty problem";
  if (!d.out_f32 && !d.out_bf16) return "gemm: no output";
  if (d.act == ACT_GEGLU && (!d.out_bf16 || d.out_f32 || d.residual || d.rowbias)) return "gemm: GEGLU epilogue writes bf16 only";
  if (d.act == ACT_GEGLU && d.N % 64 != 0) return "gemm: GEGLU needs N % 64 == 0";
  L.bn = d.bn ? d.bn : pick_bn(d);
  const GemmASeg& s0 = d.seg[0];
  const int pix = s0.H * s0.W;
  if (d.M % pix != 0 && !(s0.H == 1)) return "gemm: M must be a multiple of H*W";
  // M-tile box over (W, H, B)
  if (s0.H == 1 && s0.Bt == 1) {  // plain [M,K] matrix: rows past M are zero-filled by TMA and masked in the epilogue
    L.box_w = BM; L.box_h = 1; L.box_b = 1;
  } else if (s0.W >= BM) {
    if (s0.W % BM != 0) return "gemm: W must be a multiple of 128 when W >= 128";
    L.box_w = BM; L.box_h = 1; L.box_b = 1;
  } else {
    if (BM % s0.W != 0) return "gemm: W must divide 128";
    L.box_w = s0.W;
    int rows = BM / s0.W;
    if (s0.H >= rows) {
      if (s0.H % rows != 0) return "gemm: H must be a multiple of 128/W";
      L.box_h = rows; L.box_b = 1;
    } else {
      if (rows % s0.H != 0) return "gemm: H*W must divide 128";
      L.box_h = s0.H; L.box_b = rows / s0.H;
    }
  }
  int ktot = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const GemmASeg& sg = d.seg[s];
    if (sg.C % BK != 0 || sg.C <= 0) return "gemm: segment channels must be a positive multiple of 64";
    if (sg.ntaps < 1 || sg.ntaps > 9) return "gemm: ntaps out of range";
    if (s > 0 && (sg.H != s0.H || sg.W != s0.W)) return "gemm: segments must share the output grid";
    const int ld = sg.ld ? sg.ld : sg.C;
    if (ld % 8 != 0) return "gemm: channel pitch must be a multiple of 8";
    if ((reinterpret_cast<uintptr_t>(sg.ptr) & 15) != 0) return "gemm: A pointer must be 16B aligned";
    L.cpt[s] = sg.C / BK;
    L.kchunks[s] = sg.ntaps * L.cpt[s];
    ktot += sg.ntaps * sg.C;
    cuuint64_t dims[4] = {cuuint64_t(sg.C), cuuint64_t(sg.W), cuuint64_t(sg.H), cuuint64_t(sg.Bt)};
    cuuint64_t strides[3] = {cuuint64_t(ld) * 2, cuuint64_t(sg.W) * ld * 2, cuuint64_t(sg.H) * sg.W * ld * 2};
    cuuint32_t box[4] = {cuuint32_t(BK), cuuint32_t(L.box_w), cuuint32_t(L.box_h), cuuint32_t(L.box_b)};
    if (const char* e = encode_map(&L.tmA[s], sg.ptr, 4, dims, strides, box)) return e;
  }
  {
    const int Nw = d.Nw ? d.Nw : d.N;
    if ((reinterpret_cast<uintptr_t>(d.w) & 15) != 0) return "gemm: W pointer must be 16B aligned";
    cuuint64_t dims[2] = {cuuint64_t(ktot), cuuint64_t(Nw)};
    const int ldw = d.ldw ? d.ldw : ktot;
    if (ldw % 8 != 0 || ldw < ktot) return "gemm: weight pitch must be >= Ktot and a multiple of 8";
    cuuint64_t strides[1] = {cuuint64_t(ldw) * 2};
    cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(L.bn)};
    if (const char* e = encode_map(&L.tmB, d.w, 2, dims, strides, box)) return e;
  }
  const int Ncols = (d.act == ACT_GEGLU) ? 2 * d.N : d.N;
  const int n_tiles = (Ncols + L.bn - 1) / L.bn;
  const int m_tiles = (d.M + BM - 1) / BM;
  L.grid = dim3(unsigned(n_tiles) * unsigned(m_tiles));
  switch (L.bn) {
    case 16: L.smem = Cfg<16>::SMEM; break;
    case 32: L.smem = Cfg<32>::SMEM; break;
    case 64: L.smem = Cfg<64>::SMEM; break;
    case 128: L.smem = Cfg<128>::SMEM; break;
    case 160: L.smem = Cfg<160>::SMEM; break;
    case 192: L.smem = Cfg<192>::SMEM; break;
    case 256: L.smem = Cfg<256>::SMEM; break;
    default: return "gemm: unsupported N tile";
  }
  return nullptr;
}

template <int BN>
static const char* launch_bn(const GemmLaunch& L, const GemmParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg<BN>::SMEM)) != cudaSuccess)
      return "gemm: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    attr_set = true;
  }
  gemm_tc_kernel<BN><<<L.grid, kThreads, Cfg<BN>::SMEM, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? nullptr : "gemm: kernel launch failed";
}

const char* gemm_launch(const GemmLaunch& L, cudaStream_t stream) {
  const GemmDesc& d = L.d;
  GemmParams p;
  p.tmA[0] = L.tmA[0];
  p.tmA[1] = L.tmA[d.nseg > 1 ? 1 : 0];
  p.tmB = L.tmB;
  p.nseg = d.nseg;
  for (int s = 0; s < 2; ++s) {
    p.kchunks[s] = s < d.nseg ? L.kchunks[s] : 0;
    p.cpt[s] = s < d.nseg ? L.cpt[s] : 1;
    for (int t = 0; t < 9; ++t) {
      p.dx[s][t] = d.seg[s].dx[t];
      p.dy[s][t] = d.seg[s].dy[t];
      p.boff[s][t] = d.seg[s].b_off[t];
    }
  }
  p.M = d.M;
  p.N = (d.act == ACT_GEGLU) ? 2 * d.N : d.N;
  p.W = d.seg[0].W;
  p.pix = d.seg[0].H * d.seg[0].W;
  p.box_w = L.box_w; p.box_h = L.box_h; p.box_b = L.box_b;
  p.bias = d.bias;
  p.rowbias = d.rowbias;
  p.rows_per_img = d.rows_per_img > 0 ? d.rows_per_img : 1;
  p.ld_rowbias = d.ld_rowbias ? d.ld_rowbias : d.N;
  p.residual = d.residual; p.ldr = d.ldr;
  p.out_f32 = d.out_f32; p.ldo32 = d.ldo32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d.out_bf16); p.ldo16 = d.ldo16;
  p.act = d.act;
  p.alpha = d.alpha;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vec_ok = (d.N % 8 == 0) && (!d.bias || al16(d.bias)) && (!d.rowbias || (al16(d.rowbias) && p.ld_rowbias % 4 == 0)) &&
             (!d.residual || (al16(d.residual) && d.ldr % 4 == 0)) && (!d.out_f32 || (al16(d.out_f32) && d.ldo32 % 4 == 0)) &&
             (!d.out_bf16 || (al16(d.out_bf16) && d.ldo16 % 8 == 0));
  if (d.act == ACT_GEGLU && !p.vec_ok) return "gemm: GEGLU epilogue needs 16B-aligned outputs";
  p.n_tiles = (p.N + L.bn - 1) / L.bn;
  p.fp16 = d.fp16;
  switch (L.bn) {
    case 16: return launch_bn<16>(L, p, stream);
    case 32: return launch_bn<32>(L, p, stream);
    case 64: return launch_bn<64>(L, p, stream);
    case 128: return launch_bn<128>(L, p, stream);
    case 160: return launch_bn<160>(L, p, stream);
    case 192: return launch_bn<192>(L, p, stream);
    case 256: return launch_bn<256>(L, p, stream);
  }
  return "gemm: unsupported N tile";
}

}  // namespace madm

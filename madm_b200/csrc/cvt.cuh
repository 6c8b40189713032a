// fp32 -> 16-bit operand conversion shared by all kernels.  The compute dtype of the GEMM operands is a runtime
// choice (DT_FP16 or DT_BF16, uniform per launch): tcgen05.mma kind::f16 runs both at the same rate.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace madm {

enum DType : int { DT_BF16 = 0, DT_FP16 = 1 };

__device__ __forceinline__ uint32_t pack2_16(float a, float b, int fp16) {
  if (fp16) {
    __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
  }
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint16_t cvt_16(float a, int fp16) {
  if (fp16) {
    __half t = __float2half_rn(a);
    return *reinterpret_cast<uint16_t*>(&t);
  }
  __nv_bfloat16 t = __float2bfloat16_rn(a);
  return *reinterpret_cast<uint16_t*>(&t);
}
__device__ __forceinline__ uint2 pack4_16(float a, float b, float c, float d, int fp16) {
  uint2 r;
  r.x = pack2_16(a, b, fp16);
  r.y = pack2_16(c, d, fp16);
  return r;
}

}  // namespace madm

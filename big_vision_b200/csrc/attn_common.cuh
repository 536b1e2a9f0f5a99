// Pieces shared by the attention kernels (attention.cu: whole-key-range forward and resident
// backward for N <= 256; attention_stream.cu: key-block streaming forward and backward for any N).
#pragma once
#include "common.cuh"
#include "host_utils.h"

namespace bv {
namespace attn {

constexpr int DH = 64;
constexpr int TQ = 128;
constexpr int TILE_BYTES = TQ * DH * 2;       // 16 KB: 128 rows x 128 B
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
}

template <int R> __device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
}
template <int R> __device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
}
__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= 0 on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5,
// degree-4 polynomial for 2^f (rel. error < 5e-5, far inside the bf16 rounding of P), exponent
// patched in with integer arithmetic.  B200's MUFU.EX2 sustains ~8 lanes/clk/SM, which makes the
// exponentials the bound of the softmax; splitting them between MUFU and this path doubles the rate.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;     // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float pl = fmaf(f, 0.0096181291f, 0.0555041087f);
  pl = fmaf(pl, f, 0.2402265070f);
  pl = fmaf(pl, f, 0.6931471806f);
  pl = fmaf(pl, f, 1.0f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(t) << 23));
}



inline int make_tmap_bnd(CUtensorMap* m, const void* ptr, int cols, int64_t N, int64_t B, int64_t ld,
                  int64_t bs, uint32_t box_rows) {
  uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
  uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(bs) * 2};
  uint32_t box[3] = {64, box_rows, 1};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, strides, box, true);
}


}  // namespace attn
}  // namespace bv

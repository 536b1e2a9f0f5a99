// Key-block streaming attention forward for any sequence length (head dim 64, no mask).
// Reference: flax.linen.MultiHeadDotProductAttention as called at models/vit.py:93-98 and :176-178
// (q scaled by 1/sqrt(dh), softmax over keys, weights times v).  This is the kernel behind
// config 5 (ViT-L/14@336: 576 keys per image) where the key range no longer fits in TMEM; it also
// serves the short sequences.
//
// Work unit: a 128-query tile of one (image, head) against successive 128-key BLOCKS.
//   * two softmax warpgroups, each owning WHOLE blocks (thread = one query row, its 128 scores in
//     registers): no cross-warp exchange of row maxima or sums at all.  Block g of the CTA's stream
//     goes to warpgroup g & 1, so the two warpgroups ping-pong and one of them is always in its
//     exponential phase while the tensor core produces the other one's scores;
//   * every block is normalised by ITS OWN row maximum: P_g = exp(s - m_g) (bf16, through shared
//     memory), O_g = P_g V_g is a fresh TMEM accumulator, and the epilogue warpgroup folds the blocks
//     together in registers, acc = acc * 2^(M - M') + O_g * 2^(m_g - M'), so nothing in TMEM is ever
//     rescaled and the softmax warps keep no state between blocks;
//   * K/V blocks stream through a 3-stage TMA ring, Q is double-buffered across tiles, the S and
//     P.V products are issued by two dedicated converged warps (elect.sync, see attention.cu).
// TMEM: S_w [128 x 128] fp32 at columns w*128; O buffers [128 x 64] at 256 + (2w + b)*64  (512 total).
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"
#include "attn_common.cuh"

#include <stdlib.h>

namespace bv {
namespace {

using namespace attn;

constexpr int F2_THREADS = 512;
constexpr int BK = 128;                         // keys per block
// Shared memory: Q (2 x 16 KB) | K/V ring (NST x 32 KB) | [P: 2 x 32 KB, only when P goes through shared
// memory] | output staging 16 KB | stats 4 KB | barriers.  With P in tensor memory (TSP) the 64 KB of
// the P buffers buy two more ring stages.  (The ncu source view of the 3-stage build shows 11 % of the
// stall samples on the softmax warps' s_full wait; the deeper ring changed the 576-key time from 1.433
// to 1.421 ms, so K/V latency was not what they were waiting for -- the scores of a warpgroup's next
// block can only be issued once it has started the current one.)
template <bool TSP> struct F2L {
  static constexpr int NST = TSP ? 5 : 3;
  static constexpr int Q_OFF = 0;
  static constexpr int KV_OFF = 2 * TILE_BYTES;
  static constexpr int P_OFF = KV_OFF + NST * 2 * TILE_BYTES;
  static constexpr int O_OFF = P_OFF + (TSP ? 0 : 2 * 2 * TILE_BYTES);
  static constexpr int ST_OFF = O_OFF + TILE_BYTES;               // stats [w][buf][m|l][128] fp32 = 4 KB
  static constexpr int BAR_OFF = ST_OFF + 2 * 2 * 2 * 128 * 4;
  static constexpr int SMEM = BAR_OFF + 512 + 1024;
};
static_assert(F2L<true>::SMEM <= 232448 && F2L<false>::SMEM <= 232448, "shared memory budget");
constexpr int F2_DEFAULT_VAR = 108;              // softmax variant (see softmax_block); BV_ATTN_SM overrides

struct Fwd2Dev {
  int tiles;          // B * H * QT
  int H, QT, Nq, Nk, NB;   // NB = key blocks per tile
  float scale_log2;   // scale * log2(e)
  float* lse;         // [B, H, Nq] or null
};

// two exponentials per MUFU operation: exp2 of an fp16 pair (the arguments are <= 0; fp16 keeps them to
// 2^-11 relative, i.e. the result to <= 0.27 % for p >= 2^-16 and <= 0.07 % for p >= 1/16 -- tighter
// than the bf16 rounding of P it replaces).  lo -> lower half = lower column.
__device__ __forceinline__ uint32_t ex2_f16x2(float lo, float hi) {
  uint32_t h, e;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(h));
  return e;
}
__device__ __forceinline__ float f16x2_sum(uint32_t v) {
  float a, b;
  asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}" : "=f"(a), "=f"(b) : "r"(v));
  return a + b;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));   // three-input max (sm_100)
  return d;
}

// One block of scores for one softmax warpgroup thread: load, (mask), max, exponentials, pack.
// FULL: all 128 columns are real keys (no masking, compile-time trip counts).
// VAR (tuning switch, BV_ATTN_SM): bit 0 = four independent max / sum chains instead of one serial
// chain of 128 (each warp scheduler holds only two softmax warps, so a 128-deep dependent chain is
// pure latency); bit 1 = three-input max; bits 2.. = exponential split: 0 alternate 8-column units
// between MUFU and the FMA-pipe polynomial, 1 = one unit in four on the polynomial, 2 = all MUFU,
// 3 = all MUFU as fp16 pairs (ex2.approx.f16x2; P is then an fp16 operand of the P.V product).
template <bool FULL, int VAR, bool TSP>
__device__ __forceinline__ void softmax_block(uint32_t s_addr, uint32_t p_row, uint32_t sw, int valid,
                                              float scale_log2, uint32_t s_empty_bar, uint32_t p_empty_bar,
                                              uint32_t p_empty_parity, float& mxs_out, float& sum_out) {
  constexpr bool ILP4 = (VAR & 1) != 0, MAX3 = (VAR & 2) != 0;
  constexpr int SPLIT = VAR >> 2;
  const int lane = threadIdx.x & 31;
  // units of 8 columns that take part in the P.V product: whole 16-key MMA steps
  const int nunits = FULL ? 16 : ((valid + 15) >> 4) * 2;
#define UNIT_ON(u) (FULL || (u) < nunits)
  // All 16 units are loaded UNCONDITIONALLY (columns past `valid` hold stale accumulator data and
  // are never used): a register that is only conditionally written inside the persistent loop stays
  // live around the whole loop, and 128 such registers per variant do not fit.
  uint32_t sv[16][8];
#pragma unroll
  for (int u = 0; u < 16; ++u) tmem_ld_x8(s_addr + u * 8, sv[u]);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(s_empty_bar);
  float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float sma[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    if (UNIT_ON(u)) {
      if (!FULL) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (u * 8 + j >= valid) sv[u][j] = 0xff800000u;     // -inf: key columns past Nk
      }
      float& m = mxa[ILP4 ? (u & 3) : 0];
      if (MAX3) {
        m = max3(m, __uint_as_float(sv[u][0]), __uint_as_float(sv[u][1]));
        const float a = max3(__uint_as_float(sv[u][2]), __uint_as_float(sv[u][3]), __uint_as_float(sv[u][4]));
        const float b = max3(__uint_as_float(sv[u][5]), __uint_as_float(sv[u][6]), __uint_as_float(sv[u][7]));
        m = max3(m, a, b);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) m = fmaxf(m, __uint_as_float(sv[u][j]));
      }
    }
  }
  const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
  const float mxs = mx * scale_log2;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    if (UNIT_ON(u)) {
      if (SPLIT == 3) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t e2 = ex2_f16x2(fmaf(__uint_as_float(sv[u][2 * jj]), scale_log2, -mxs),
                                        fmaf(__uint_as_float(sv[u][2 * jj + 1]), scale_log2, -mxs));
          sma[ILP4 ? jj : 0] += f16x2_sum(e2);
          sv[u][jj] = e2;
        }
        continue;
      }
      float e[8];
      // exp2(scale * s - scale * max), split between MUFU and the FMA-pipe polynomial
      const bool poly = SPLIT == 0 ? (u & 1) : SPLIT == 1 ? ((u & 3) == 3) : false;
      if (poly) {
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = ex2_poly(fmaf(__uint_as_float(sv[u][j]), scale_log2, -mxs));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = ex2_mufu(fmaf(__uint_as_float(sv[u][j]), scale_log2, -mxs));
      }
      if (ILP4) {
        sma[0] += e[0] + e[4]; sma[1] += e[1] + e[5]; sma[2] += e[2] + e[6]; sma[3] += e[3] + e[7];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) sma[0] += e[j];
      }
      sv[u][0] = pack_bf16(e[0], e[1]); sv[u][1] = pack_bf16(e[2], e[3]);
      sv[u][2] = pack_bf16(e[4], e[5]); sv[u][3] = pack_bf16(e[6], e[7]);
    }
  }
  const float sum = (sma[0] + sma[1]) + (sma[2] + sma[3]);
  // the P buffer is free once the P.V product of this warpgroup's previous block has retired
  mbar_wait(p_empty_bar, p_empty_parity);
  if (TSP) {
    // P stays in tensor memory (operand A of the P.V product): unit u = packed columns 4u..4u+3 of the
    // P region that follows this warpgroup's S region; `p_row` carries that TMEM address here
    tc_fence_after();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (FULL || g * 4 < nunits) {
        uint32_t w16[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          w16[q * 4 + 0] = sv[g * 4 + q][0]; w16[q * 4 + 1] = sv[g * 4 + q][1];
          w16[q * 4 + 2] = sv[g * 4 + q][2]; w16[q * 4 + 3] = sv[g * 4 + q][3];
        }
        tmem_st_32x32b_x16(p_row + g * 16, w16);
      }
    }
    tmem_st_wait();
    tc_fence_before();
  } else {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (UNIT_ON(u)) {
        const uint32_t k = static_cast<uint32_t>(u);
        const uint32_t addr = p_row + (k >> 3) * TILE_BYTES + (((k & 7u) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(sv[u][0]),
                     "r"(sv[u][1]), "r"(sv[u][2]), "r"(sv[u][3]) : "memory");
      }
    }
  }
#undef UNIT_ON
  mxs_out = mxs;
  sum_out = sum;
}

// TSP: P stays in tensor memory (tcgen05.st, A operand of the P.V product read from TMEM): no 32 KB
// store + 32 KB MMA read of P per block through shared memory.  TMEM then holds S_w | P_w at w * 192
// (128 + 64 columns) and ONE O accumulator per warpgroup at 384 + w * 64.
template <int VAR, bool TSP>
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd_stream_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                       const Fwd2Dev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  using LY = F2L<TSP>;
  constexpr int F2_NST = LY::NST, F2_Q_OFF = LY::Q_OFF, F2_KV_OFF = LY::KV_OFF, F2_P_OFF = LY::P_OFF,
                F2_O_OFF = LY::O_OFF, F2_ST_OFF = LY::ST_OFF, F2_BAR_OFF = LY::BAR_OFF;

  const uint32_t bar = base + F2_BAR_OFF;
  auto q_full = [&](int s) { return bar + 8u * s; };                    // 0,1
  auto q_empty = [&](int s) { return bar + 8u * (2 + s); };             // 2,3
  auto kv_full = [&](int s) { return bar + 8u * (4 + s); };             // 4..8
  auto kv_empty = [&](int s) { return bar + 8u * (9 + s); };            // 9..13
  auto s_full = [&](int w) { return bar + 8u * (14 + w); };             // 14,15
  auto s_empty = [&](int w) { return bar + 8u * (16 + w); };            // 16,17
  auto p_full = [&](int w) { return bar + 8u * (18 + w); };             // 18,19
  auto p_empty = [&](int w) { return bar + 8u * (20 + w); };            // 20,21
  auto o_full = [&](int w, int b) { return bar + 8u * (22 + 2 * w + b); };    // 22..25
  auto o_empty = [&](int w, int b) { return bar + 8u * (26 + 2 * w + b); };   // 26..29
  auto st_full = [&](int w, int b) { return bar + 8u * (30 + 2 * w + b); };   // 30..33
  const uint32_t tmem_slot = bar + 8u * 34;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + F2_BAR_OFF + 8 * 34);
  float* stats = reinterpret_cast<float*>(base_ptr + F2_ST_OFF);   // [(w*2+b)*2 + {0: m, 1: l}][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    for (int s = 0; s < 2; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    for (int s = 0; s < F2_NST; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int w = 0; w < 2; ++w) {
      mbar_init(s_full(w), 1);  mbar_init(s_empty(w), 4);
      mbar_init(p_full(w), 4);  mbar_init(p_empty(w), 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(o_full(w, b), 1); mbar_init(o_empty(w, b), 4); mbar_init(st_full(w, b), 4);
      }
    }
    fence_barrier_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // TMEM columns and the O / statistics buffer of warpgroup w's n-th block
  auto s_col = [](uint32_t w) { return TSP ? w * 192u : w * 128u; };
  auto p_col = [](uint32_t w) { return w * 192u + 128u; };                       // TSP only
  auto o_col = [](uint32_t w, uint32_t ob) { return TSP ? 384u + w * 64u : 256u + (2u * w + ob) * 64u; };
  auto OB = [](uint32_t n) { return TSP ? 0u : (n & 1u); };                       // buffer index
  auto OPH = [](uint32_t n) { return TSP ? (n & 1u) : ((n >> 1) & 1u); };         // parity of this use

  const int my_tiles = (p.tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  const int NB = p.NB;

  if (warp >= 8 && warp < 12) {
    reg_dec<40>();
    if (warp == 8) {
      // ================= TMA producer =================
      if (lane == 0) {
        uint32_t g = 0;
        for (int i = 0; i < my_tiles; ++i) {
          const int tile = blockIdx.x + i * gridDim.x;
          const int qt = tile % p.QT;
          const int bh = tile / p.QT;
          const int h = bh % p.H, b = bh / p.H;
          const int qs = i & 1;
          mbar_wait(q_empty(qs), ((static_cast<uint32_t>(i) >> 1) & 1u) ^ 1u);
          mbar_expect_tx(q_full(qs), TILE_BYTES);
          tma_load_3d(base + F2_Q_OFF + qs * TILE_BYTES, &tmQ, q_full(qs), h * DH, qt * TQ, b);
          for (int j = 0; j < NB; ++j, ++g) {
            const uint32_t st = g % F2_NST;
            mbar_wait(kv_empty(st), ((g / F2_NST) & 1u) ^ 1u);
            const uint32_t k_s = base + F2_KV_OFF + st * 2 * TILE_BYTES;
            mbar_expect_tx(kv_full(st), 2 * TILE_BYTES);
            tma_load_3d(k_s, &tmK, kv_full(st), h * DH, j * BK, b);
            tma_load_3d(k_s + TILE_BYTES, &tmV, kv_full(st), h * DH, j * BK, b);
          }
        }
      }
    } else if (warp == 9) {
      // ================= S = Q K^T issuer (converged warp, one elected lane issues) =================
      const uint32_t idesc_s = umma_idesc_bf16(128, BK, 0, 0);   // both operands K-major
      uint32_t g = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int qs = i & 1;
        mbar_wait(q_full(qs), (static_cast<uint32_t>(i) >> 1) & 1u);
        const uint64_t dq = umma_smem_desc_sw128(base + F2_Q_OFF + qs * TILE_BYTES, 16, 1024);
        for (int j = 0; j < NB; ++j, ++g) {
          const uint32_t w = g & 1u, st = g % F2_NST;
          mbar_wait(kv_full(st), (g / F2_NST) & 1u);
          mbar_wait(s_empty(w), ((g >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint64_t dk = umma_smem_desc_sw128(base + F2_KV_OFF + st * 2 * TILE_BYTES, 16, 1024);
          const uint32_t d = tmem_base + s_col(w);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(d, dq + k * 2, dk + k * 2, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(s_full(w));
          }
          __syncwarp();
        }
      }
    } else if (warp == 10) {
      // ================= O_g = P_g V_g issuer =================
      // V is MN-major; P is bf16, or fp16 when the exponentials come out of MUFU as fp16 pairs
      // (instruction descriptor bits [7,10) = A format: 0 = f16, 1 = bf16)
      const uint32_t idesc_o = (VAR >> 2) == 3 ? (umma_idesc_bf16(128, DH, 0, 1) & ~(7u << 7))
                                               : umma_idesc_bf16(128, DH, 0, 1);
      uint32_t g = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int qs = i & 1;
        for (int j = 0; j < NB; ++j, ++g) {
          const uint32_t w = g & 1u, st = g % F2_NST, n = g >> 1, ob = OB(n);
          int valid = p.Nk - j * BK;
          if (valid > BK) valid = BK;
          const int ksteps = (valid + 15) >> 4;
          mbar_wait(p_full(w), n & 1u);
          mbar_wait(kv_full(st), (g / F2_NST) & 1u);
          mbar_wait(o_empty(w, ob), OPH(n) ^ 1u);
          tc_fence_after();
          const uint64_t dpd = umma_smem_desc_sw128(base + F2_P_OFF + w * 2 * TILE_BYTES, 16, 1024);
          const uint64_t dvd = umma_smem_desc_sw128(base + F2_KV_OFF + st * 2 * TILE_BYTES + TILE_BYTES, 8192, 1024);
          const uint32_t d = tmem_base + o_col(w, ob);
          const uint32_t pt = tmem_base + p_col(w);
          if (elect_one()) {
            for (int kk = 0; kk < ksteps; ++kk) {
              if (TSP) umma_bf16_ts(d, pt + kk * 8, dvd + kk * 128, idesc_o, kk > 0 ? 1u : 0u);
              else umma_bf16_ss(d, dpd + (kk >> 2) * (TILE_BYTES / 16) + (kk & 3) * 2, dvd + kk * 128, idesc_o,
                                kk > 0 ? 1u : 0u);
            }
            umma_commit(o_full(w, ob));
            umma_commit(p_empty(w));
            umma_commit(kv_empty(st));       // K_g was consumed by S_g before softmax could finish P_g
            if (j == NB - 1) umma_commit(q_empty(qs));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 12) {
    // ================= epilogue warpgroup: fold the blocks of a tile, normalise, store =================
    reg_dec<120>();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - 384;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t o_s = base + F2_O_OFF;
    const int pQT = pin_reg(p.QT), pH = pin_reg(p.H), pNq = pin_reg(p.Nq);
    float* __restrict__ p_lse = pin_reg(p.lse);
    uint32_t g = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int qt = tile % pQT;
      const int bh = tile / pQT;
      const int h = bh % pH, b = bh / pH;
      const bool active = quarter * 32 < pNq - qt * TQ;     // any real query row in this warp?
      float acc[DH];
#pragma unroll
      for (int c = 0; c < DH; ++c) acc[c] = 0.f;
      float M = -INFINITY, Lsum = 0.f;
      for (int j = 0; j < NB; ++j, ++g) {
        const uint32_t w = g & 1u, n = g >> 1, ob = OB(n);
        mbar_wait(o_full(w, ob), OPH(n));
        mbar_wait(st_full(w, ob), OPH(n));
        tc_fence_after();
        if (active) {
          const float* st = stats + ((2 * w + ob) * 2) * 128;
          const float mb = st[row], lb = st[128 + row];
          const float Mn = fmaxf(M, mb);
          const float fa = ex2_mufu(M - Mn), fb = ex2_mufu(mb - Mn);
          const uint32_t o_addr = tmem_base + lane_addr + o_col(w, ob);
#pragma unroll
          for (int c = 0; c < DH / 16; ++c) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(o_addr + c * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k)
              acc[c * 16 + k] = fmaf(acc[c * 16 + k], fa, __uint_as_float(ov[k]) * fb);
          }
          Lsum = fmaf(Lsum, fa, lb * fb);
          M = Mn;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty(w, ob));
      }
      const float inv = 1.0f / Lsum;
      const int qrow = qt * TQ + row;
      if (active && qrow < pNq && p_lse != nullptr)
        p_lse[static_cast<int64_t>(bh) * pNq + qrow] = (M + log2f(Lsum)) * LN2;
      if (etid == 0) tma_store_wait_read<0>();
      named_bar_sync(3, 128);
      if (active) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t addr = o_s + row * 128 + ((static_cast<uint32_t>(c) ^ sw) << 4);
          float f[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = acc[c * 8 + k] * inv;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                       "r"(pack_bf16(f[0], f[1])), "r"(pack_bf16(f[2], f[3])),
                       "r"(pack_bf16(f[4], f[5])), "r"(pack_bf16(f[6], f[7])) : "memory");
        }
      }
      fence_proxy_async();
      named_bar_sync(3, 128);
      if (etid == 0) {
        tma_store_3d(&tmO, o_s, h * DH, qt * TQ, b);     // rows past Nq are clipped by the tensor map
        tma_store_commit();
      }
    }
    if (etid == 0) tma_store_wait<0>();
  } else {
    // ================= softmax warpgroups 0 and 1 =================
    reg_inc<176>();
    const int w = warp >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + s_col(w);
    const uint32_t p_row = TSP ? tmem_base + lane_addr + p_col(w) : base + F2_P_OFF + w * 2 * TILE_BYTES + row * 128;
    const float scale_log2 = pin_reg(p.scale_log2);
    const int pQT = pin_reg(p.QT), pNq = pin_reg(p.Nq), pNk = pin_reg(p.Nk);
    const uint32_t b_s_full = s_full(w), b_s_empty = s_empty(w), b_p_full = p_full(w), b_p_empty = p_empty(w);
    uint32_t g = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int qt = tile % pQT;
      const bool active = quarter * 32 < pNq - qt * TQ;
      for (int j = 0; j < NB; ++j, ++g) {
        if ((g & 1u) != static_cast<uint32_t>(w)) continue;
        const uint32_t n = g >> 1, ob = OB(n);
        int valid = pNk - j * BK;
        mbar_wait(b_s_full, n & 1u);
        tc_fence_after();
        if (!active) {
          // no real query row in this warp (tile past Nq): keep the barrier protocol going, touch no
          // data.  Its P rows stay stale; they only feed output rows the TMA store clips.
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(b_s_empty);
          mbar_wait(b_p_empty, (n & 1u) ^ 1u);
          mbar_wait(o_empty(w, ob), OPH(n) ^ 1u);
          __syncwarp();
          if (lane == 0) { mbar_arrive(b_p_full); mbar_arrive(st_full(w, ob)); }
          continue;
        }
        float mxs, sum;
        if (valid >= BK)
          softmax_block<true, VAR, TSP>(s_addr, p_row, sw, BK, scale_log2, b_s_empty, b_p_empty, (n & 1u) ^ 1u, mxs, sum);
        else
          softmax_block<false, VAR, TSP>(s_addr, p_row, sw, valid, scale_log2, b_s_empty, b_p_empty, (n & 1u) ^ 1u, mxs, sum);
        // statistics of this block for the epilogue; the slot is free once the epilogue has consumed
        // the previous block that used this (w, ob) buffer pair
        mbar_wait(o_empty(w, ob), OPH(n) ^ 1u);
        {
          float* st = stats + ((2 * w + ob) * 2) * 128;
          st[row] = mxs;
          st[128 + row] = sum;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(b_p_full); mbar_arrive(st_full(w, ob)); }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ============================================================================
// streaming backward
// ============================================================================
// Work unit: one 128-key tile kt of one (image, head) against ALL query tiles of that item.
//   per pair (kt, qt):  S = Q K^T, dP = dO V^T (TMEM) -> P = exp(scale S - lse), dS = scale P o (dP - delta)
//                       (registers -> bf16 shared tiles) -> dV += P^T dO, dK += dS^T Q, dQ_pair = dS K
//   dV / dK accumulate in TMEM over the query tiles and are written once per unit (bf16 TMA store, bias
//   gradients fused); dQ needs the sum over key tiles, which for long sequences cannot stay in TMEM
//   (5 x 64 columns at 576 queries): each pair's dQ tile goes out as an fp32 TMA REDUCE-ADD into a
//   caller-provided accumulation buffer (zeroed by the launcher), converted to bf16 afterwards.
//   delta = rowsum(O o dO) comes from a small pre-kernel, so the compute warps have no prologue.
// K/V tiles are double-buffered across units, Q/dO tiles stream through a 2-slot ring, S/dP of pair
// j+1 are issued before the gradient products of pair j (as in attention.cu's resident kernel).
constexpr int B2_THREADS = 512;
constexpr int B2_KV_OFF = 0;                             // 2 x (K | V)       64 KB
constexpr int B2_QD_OFF = 4 * TILE_BYTES;                // 2 x (Q | dO)      64 KB
constexpr int B2_P_OFF = 8 * TILE_BYTES;                 // P  [128 x 128]    32 KB
constexpr int B2_DS_OFF = 10 * TILE_BYTES;               // dS [128 x 128]    32 KB
constexpr int B2_STG_OFF = 12 * TILE_BYTES;              // staging           32 KB
constexpr int B2_BAR_OFF = 14 * TILE_BYTES;
constexpr int B2_SMEM = B2_BAR_OFF + 256 + 1024;

struct Bwd2Dev {
  int groups;          // B * H * KT
  int H, Nq, Nk, QT, KT;
  float scale, scale_log2;
  const float* lse;    // [B, H, Nq]
  const float* delta;  // [B, H, Nq]
  float* dk_colsum; float* dv_colsum;   // optional [H*64] bias gradients
  int variant;         // BV_BWD_VARIANT: bit 2 = alternate MUFU / polynomial exponentials (default: all MUFU)
};

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd_stream_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                       const __grid_constant__ CUtensorMap tmdQacc, const __grid_constant__ CUtensorMap tmdK,
                       const __grid_constant__ CUtensorMap tmdV, const Bwd2Dev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t p_s = base + B2_P_OFF, ds_s = base + B2_DS_OFF, stg_s = base + B2_STG_OFF;
  const uint32_t bar = base + B2_BAR_OFF;
  auto kv_full = [&](uint32_t s) { return bar + 8u * s; };              // 0,1
  auto kv_empty = [&](uint32_t s) { return bar + 8u * (2 + s); };       // 2,3
  auto qd_full = [&](uint32_t s) { return bar + 8u * (4 + s); };        // 4,5
  auto qd_empty = [&](uint32_t s) { return bar + 8u * (6 + s); };       // 6,7
  const uint32_t sdp_full = bar + 64, sdp_empty = bar + 72, pds_full = bar + 80, pds_empty = bar + 88;
  const uint32_t dkv_full = bar + 96, dkv_empty = bar + 104;
  auto dq_full = [&](uint32_t b) { return bar + 112u + 8u * b; };       // 112,120
  auto dq_empty = [&](uint32_t b) { return bar + 128u + 8u * b; };      // 128,136
  const uint32_t tmem_slot = bar + 144;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + B2_BAR_OFF + 144);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1);
      mbar_init(qd_full(s), 1); mbar_init(qd_empty(s), 1);
      mbar_init(dq_full(s), 1); mbar_init(dq_empty(s), 4);
    }
    mbar_init(sdp_full, 1);  mbar_init(sdp_empty, 8);
    mbar_init(pds_full, 8);  mbar_init(pds_empty, 1);
    mbar_init(dkv_full, 1);  mbar_init(dkv_empty, 4);
    fence_barrier_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  constexpr uint32_t S_COL = 0, DP_COL = 128, DV_COL = 256, DK_COL = 320, DQ_COL = 384;

  const int my_groups = (p.groups - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                        static_cast<int>(gridDim.x);
  const int QT = p.QT;

  if (warp >= 12) {
    // ---------------- gradient write-out warpgroup ----------------
    reg_dec<104>();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - 384;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const int pH = pin_reg(p.H), pKT = pin_reg(p.KT), pNk = pin_reg(p.Nk);
    // bf16 tile (dV / dK): TMEM -> bf16 -> swizzled staging -> TMA store, + fused bias gradient
    auto write_tile = [&](uint32_t tcol, const CUtensorMap* tm, int h, int r0, int b, float* colsum, int nvalid,
                          uint32_t release) {
      uint32_t a[64];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + tcol, *reinterpret_cast<uint32_t(*)[32]>(&a[0]));
      tmem_ld_32x32b_x32(tmem_base + lane_addr + tcol + 32, *reinterpret_cast<uint32_t(*)[32]>(&a[32]));
      tmem_ld_wait();
      if (release != 0u) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(release);
      }
      if (etid == 0) tma_store_wait_read<0>();
      named_bar_sync(3, 128);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t addr = stg_s + row * 128 + ((static_cast<uint32_t>(g) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                     "r"(pack_bf16(__uint_as_float(a[g * 8 + 0]), __uint_as_float(a[g * 8 + 1]))),
                     "r"(pack_bf16(__uint_as_float(a[g * 8 + 2]), __uint_as_float(a[g * 8 + 3]))),
                     "r"(pack_bf16(__uint_as_float(a[g * 8 + 4]), __uint_as_float(a[g * 8 + 5]))),
                     "r"(pack_bf16(__uint_as_float(a[g * 8 + 6]), __uint_as_float(a[g * 8 + 7]))) : "memory");
      }
      fence_proxy_async();
      named_bar_sync(3, 128);
      if (colsum != nullptr) {
        // column sums of the staged (bf16-rounded) tile over its valid rows: warp w sums rows
        // 32w..32w+31, lane l the column pair (2l, 2l+1)
        const int r_lo = (etid >> 5) * 32;
        const uint32_t cp = static_cast<uint32_t>(etid & 31);
        const uint32_t cbase = stg_s + (cp & 3) * 4;
        float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int r = r_lo + i;
          uint32_t wv;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv)
                       : "r"(cbase + r * 128 + (((cp >> 2) ^ static_cast<uint32_t>(r & 7)) << 4)));
          if (r0 + r >= nvalid) wv = 0u;
          s0[i & 1] += bf16_lo(wv);
          s1[i & 1] += bf16_hi(wv);
        }
        atomicAdd(colsum + h * DH + 2 * cp, s0[0] + s0[1]);
        atomicAdd(colsum + h * DH + 2 * cp + 1, s1[0] + s1[1]);
      }
      if (etid == 0) {
        tma_store_3d(tm, stg_s, h * DH, r0, b);
        tma_store_commit();
      }
    };
    uint32_t pc = 0;
    for (int gi = 0; gi < my_groups; ++gi) {
      const int group = blockIdx.x + gi * gridDim.x;
      const int kt = group % pKT, bh = group / pKT;
      const int h = bh % pH, b = bh / pH;
      for (int qt = 0; qt < QT; ++qt, ++pc) {
        // this pair's dQ contribution: fp32 tile -> two 128-byte-swizzled slabs -> TMA reduce-add
        const uint32_t buf = pc & 1u;
        mbar_wait(dq_full(buf), (pc >> 1) & 1u);
        tc_fence_after();
        uint32_t a[64];
        tmem_ld_32x32b_x32(tmem_base + lane_addr + DQ_COL + buf * DH, *reinterpret_cast<uint32_t(*)[32]>(&a[0]));
        tmem_ld_32x32b_x32(tmem_base + lane_addr + DQ_COL + buf * DH + 32, *reinterpret_cast<uint32_t(*)[32]>(&a[32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_empty(buf));
        if (etid == 0) tma_store_wait_read<0>();
        named_bar_sync(3, 128);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t addr = stg_s + sl * TILE_BYTES + row * 128 + ((static_cast<uint32_t>(c) ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a[sl * 32 + c * 4 + 0]),
                         "r"(a[sl * 32 + c * 4 + 1]), "r"(a[sl * 32 + c * 4 + 2]), "r"(a[sl * 32 + c * 4 + 3])
                         : "memory");
          }
        }
        fence_proxy_async();
        named_bar_sync(3, 128);
        if (etid == 0) {
          tma_reduce_add_3d(&tmdQacc, stg_s, h * DH, qt * TQ, b);
          tma_reduce_add_3d(&tmdQacc, stg_s + TILE_BYTES, h * DH + 32, qt * TQ, b);
          tma_store_commit();
        }
      }
      mbar_wait(dkv_full, static_cast<uint32_t>(gi) & 1u);
      tc_fence_after();
      write_tile(DV_COL, &tmdV, h, kt * TQ, b, p.dv_colsum, pNk, 0u);
      write_tile(DK_COL, &tmdK, h, kt * TQ, b, p.dk_colsum, pNk, dkv_empty);
    }
    if (etid == 0) tma_store_wait<0>();
  } else if (warp >= 8) {
    reg_dec<56>();
  }
  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t pc = 0;
      for (int gi = 0; gi < my_groups; ++gi) {
        const int group = blockIdx.x + gi * gridDim.x;
        const int kt = group % p.KT, bh = group / p.KT;
        const int h = bh % p.H, b = bh / p.H;
        const uint32_t ks = static_cast<uint32_t>(gi) & 1u;
        mbar_wait(kv_empty(ks), ((static_cast<uint32_t>(gi) >> 1) & 1u) ^ 1u);
        mbar_expect_tx(kv_full(ks), 2 * TILE_BYTES);
        tma_load_3d(base + B2_KV_OFF + ks * 2 * TILE_BYTES, &tmK, kv_full(ks), h * DH, kt * TQ, b);
        tma_load_3d(base + B2_KV_OFF + ks * 2 * TILE_BYTES + TILE_BYTES, &tmV, kv_full(ks), h * DH, kt * TQ, b);
        for (int qt = 0; qt < QT; ++qt, ++pc) {
          const uint32_t qs = pc & 1u;
          mbar_wait(qd_empty(qs), ((pc >> 1) & 1u) ^ 1u);
          mbar_expect_tx(qd_full(qs), 2 * TILE_BYTES);
          tma_load_3d(base + B2_QD_OFF + qs * 2 * TILE_BYTES, &tmQ, qd_full(qs), h * DH, qt * TQ, b);
          tma_load_3d(base + B2_QD_OFF + qs * 2 * TILE_BYTES + TILE_BYTES, &tmdO, qd_full(qs), h * DH, qt * TQ, b);
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer (whole warp converged, one elected lane issues) ----------------
    const uint32_t id_kk = umma_idesc_bf16(128, 128, 0, 0);   // S, dP
    const uint32_t id_mm = umma_idesc_bf16(128, DH, 1, 1);    // dV, dK : A^T (MN) x B (MN)
    const uint32_t id_km = umma_idesc_bf16(128, DH, 0, 1);    // dQ     : A (K)  x B (MN)
    const int total = my_groups * QT;
    // S = Q K^T and dP = dO V^T of pair pc (unit gi, query tile qt)
    auto issue_sdp = [&](uint32_t pc, uint32_t gi, int qt) {
      const uint32_t ks = gi & 1u, qs = pc & 1u;
      if (qt == 0) mbar_wait(kv_full(ks), (gi >> 1) & 1u);
      mbar_wait(qd_full(qs), (pc >> 1) & 1u);
      mbar_wait(sdp_empty, (pc & 1u) ^ 1u);
      tc_fence_after();
      uint32_t ka = base + B2_KV_OFF + ks * 2 * TILE_BYTES, qa = base + B2_QD_OFF + qs * 2 * TILE_BYTES;
      asm volatile("" : "+r"(ka), "+r"(qa));     // see attention.cu: keep the descriptors out of spill slots
      const uint64_t dq_k = umma_smem_desc_sw128(qa, 16, 1024), dk_k = umma_smem_desc_sw128(ka, 16, 1024);
      const uint64_t ddo_k = umma_smem_desc_sw128(qa + TILE_BYTES, 16, 1024);
      const uint64_t dv_k = umma_smem_desc_sw128(ka + TILE_BYTES, 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_bf16_ss(tmem_base + S_COL, dq_k + k * 2, dk_k + k * 2, id_kk, k > 0 ? 1u : 0u);
          umma_bf16_ss(tmem_base + DP_COL, ddo_k + k * 2, dv_k + k * 2, id_kk, k > 0 ? 1u : 0u);
        }
        umma_commit(sdp_full);
      }
      __syncwarp();
    };
    auto issue_grads = [&](uint32_t pc, uint32_t gi, int qt) {
      const uint32_t ks = gi & 1u, qs = pc & 1u, buf = pc & 1u;
      uint32_t ka = base + B2_KV_OFF + ks * 2 * TILE_BYTES, qa = base + B2_QD_OFF + qs * 2 * TILE_BYTES;
      uint32_t pa = p_s, dsa = ds_s;
      asm volatile("" : "+r"(ka), "+r"(qa), "+r"(pa), "+r"(dsa));
      mbar_wait(pds_full, pc & 1u);
      if (qt == 0) mbar_wait(dkv_empty, (gi & 1u) ^ 1u);
      mbar_wait(dq_empty(buf), ((pc >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint64_t dp_mn = umma_smem_desc_sw128(pa, TILE_BYTES, 1024);
      const uint64_t dds_mn = umma_smem_desc_sw128(dsa, TILE_BYTES, 1024);
      const uint64_t dds_k = umma_smem_desc_sw128(dsa, 16, 1024);
      const uint64_t ddo_mn = umma_smem_desc_sw128(qa + TILE_BYTES, 8192, 1024);
      const uint64_t dq_mn = umma_smem_desc_sw128(qa, 8192, 1024);
      const uint64_t dk_mn = umma_smem_desc_sw128(ka, 8192, 1024);
      const bool last_q = (qt == QT - 1);
      if (elect_one()) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint32_t accv = (qt > 0 || jj > 0) ? 1u : 0u;
          umma_bf16_ss(tmem_base + DV_COL, dp_mn + jj * 128, ddo_mn + jj * 128, id_mm, accv);
          umma_bf16_ss(tmem_base + DK_COL, dds_mn + jj * 128, dq_mn + jj * 128, id_mm, accv);
          umma_bf16_ss(tmem_base + DQ_COL + buf * DH, dds_k + (jj >> 2) * (TILE_BYTES / 16) + (jj & 3) * 2,
                       dk_mn + jj * 128, id_km, jj > 0 ? 1u : 0u);
        }
        umma_commit(pds_empty);
        umma_commit(qd_empty(qs));
        umma_commit(dq_full(buf));
        if (last_q) { umma_commit(dkv_full); umma_commit(kv_empty(ks)); }
      }
      __syncwarp();
    };
    if (total > 0) issue_sdp(0u, 0u, 0);
    uint32_t pc = 0;
    for (int gi = 0; gi < my_groups; ++gi) {
      for (int qt = 0; qt < QT; ++qt, ++pc) {
        if (static_cast<int>(pc) + 1 < total) {
          const bool wrap = (qt + 1 == QT);
          issue_sdp(pc + 1, static_cast<uint32_t>(wrap ? gi + 1 : gi), wrap ? 0 : qt + 1);
        }
        issue_grads(pc, static_cast<uint32_t>(gi), qt);
      }
    }
  } else if (warp < 8) {
    // ---------------- compute warps ----------------
    reg_inc<176>();
    const int quarter = warp & 3, hf = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const int pNq = pin_reg(p.Nq), pKT = pin_reg(p.KT);
    const float p_scale = pin_reg(p.scale), p_scale_log2 = pin_reg(p.scale_log2);
    const float* __restrict__ p_lse = pin_reg(p.lse);
    const float* __restrict__ p_delta = pin_reg(p.delta);
    const bool mixed_exp = (pin_reg(p.variant) & 4) != 0;
    uint32_t pc = 0;
    // row statistics (lse, delta) are fetched ONE PAIR AHEAD: the ncu source view of the first build
    // showed 7 % of all stall samples on the use of this load right before the S/dP wait (measured
    // effect on the kernel time: none outside the noise, the wait simply moved to sdp_full)
    auto load_stats = [&](int gi_, int qt_, float& l2_, float& dl_) {
      l2_ = INFINITY; dl_ = 0.f;
      if (gi_ >= my_groups) return;
      const int bh_ = (static_cast<int>(blockIdx.x) + gi_ * static_cast<int>(gridDim.x)) / pKT;
      const int qrow_ = qt_ * TQ + row;
      // padded query rows: lse = +inf makes P exactly 0, their dO rows are TMA zero-fill
      if (qrow_ < pNq) {
        l2_ = __ldg(p_lse + static_cast<int64_t>(bh_) * pNq + qrow_);
        dl_ = __ldg(p_delta + static_cast<int64_t>(bh_) * pNq + qrow_);
      }
    };
    float l2n, dln;
    load_stats(0, 0, l2n, dln);
    const int pNk = pin_reg(p.Nk);
    for (int gi = 0; gi < my_groups; ++gi) {
      const int kt_cur = (static_cast<int>(blockIdx.x) + gi * static_cast<int>(gridDim.x)) % pKT;
      for (int qt = 0; qt < QT; ++qt, ++pc) {
        const uint32_t pp = pc & 1u;
        const float l2 = l2n * LOG2E, dl = dln;
        if (qt + 1 < QT) load_stats(gi, qt + 1, l2n, dln); else load_stats(gi + 1, 0, l2n, dln);
        mbar_wait(sdp_full, pp);
        tc_fence_after();
        float pe[64];
        {
          uint32_t t0[32], t1[32];
          tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + hf * 64, t0);
          tmem_ld_32x32b_x32(tmem_base + lane_addr + S_COL + hf * 64 + 32, t1);
          tmem_ld_wait();
          // no key masking needed: padded key columns only feed dV/dK rows that the TMA store clips
          // and a dQ product against zero-filled K rows; everything stays finite
          // live key columns of this thread's half in this key tile / any real query row in this warp:
          // P = dS = 0 elsewhere (576 tokens leave 64 rows / columns in the fifth tile)
#ifdef BV_NO_DEAD_SKIP
          const int ncl = 64;
          if (false) {
#else
          int ncl = pNk - kt_cur * TQ - hf * 64;
          ncl = ncl > 64 ? 64 : ncl;
          if (ncl <= 0 || (qt * TQ + quarter * 32 >= pNq)) {
#endif
#pragma unroll
            for (int j = 0; j < 64; ++j) pe[j] = 0.f;
          } else if (ncl < 64) {
#pragma unroll
            for (int u8 = 0; u8 < 8; ++u8) {
              if (u8 * 8 < ncl) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const int j = u8 * 8 + jj;
                  const float sj = __uint_as_float(j < 32 ? t0[j & 31] : t1[j & 31]);
                  pe[j] = ex2_mufu(fmaf(sj, p_scale_log2, -l2));
                }
              } else {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) pe[u8 * 8 + jj] = 0.f;
              }
            }
          } else if (!mixed_exp) {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const float sj = __uint_as_float(j < 32 ? t0[j & 31] : t1[j & 31]);
              pe[j] = ex2_mufu(fmaf(sj, p_scale_log2, -l2));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const float sj = __uint_as_float(j < 32 ? t0[j & 31] : t1[j & 31]);
              const float xa = fmaf(sj, p_scale_log2, -l2);
              pe[j] = (j & 2) ? ex2_poly(xa) : ex2_mufu(xa);
            }
          }
        }
        mbar_wait(pds_empty, pp ^ 1u);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t dv[32];
          tmem_ld_32x32b_x32(tmem_base + lane_addr + DP_COL + hf * 64 + c * 32, dv);
          tmem_ld_wait();
          if (c == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_empty);
          }
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) {
            const int u = c * 4 + uu;
            uint32_t pk[4], dk[4];
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) {
              const int j = u * 8 + j2 * 2;
              const float d0 = p_scale * pe[j] * (__uint_as_float(dv[(j & 31)]) - dl);
              const float d1 = p_scale * pe[j + 1] * (__uint_as_float(dv[(j & 31) + 1]) - dl);
              pk[j2] = pack_bf16(pe[j], pe[j + 1]);
              dk[j2] = pack_bf16(d0, d1);
            }
            const uint32_t off = hf * TILE_BYTES + row * 128 + ((static_cast<uint32_t>(u) ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_s + off), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_s + off), "r"(dk[0]), "r"(dk[1]), "r"(dk[2]), "r"(dk[3]) : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(pds_full);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b,h,t] = sum_j O[b,t,h*64+j] * dO[b,t,h*64+j]: eight lanes per (b, t, h) row of 64
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta,
                  int64_t B, int H, int N, int64_t ldo, int64_t bso, int64_t lddo, int64_t bsdo) {
  const int64_t total = B * N * H;                    // head rows
  const int chunk = threadIdx.x & 7;
  for (int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 3; r < total;
       r += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 3) {
    const int h = static_cast<int>(r % H);
    const int64_t bt = r / H;
    const int t = static_cast<int>(bt % N);
    const int64_t b = bt / N;
    const uint4 ao = ld_nc_na(reinterpret_cast<const uint4*>(o + b * bso + t * ldo + h * DH + chunk * 8));
    const uint4 ad = ld_nc_na(reinterpret_cast<const uint4*>(d_o + b * bsdo + t * lddo + h * DH + chunk * 8));
    float acc = bf16_lo(ao.x) * bf16_lo(ad.x) + bf16_hi(ao.x) * bf16_hi(ad.x);
    acc += bf16_lo(ao.y) * bf16_lo(ad.y) + bf16_hi(ao.y) * bf16_hi(ad.y);
    acc += bf16_lo(ao.z) * bf16_lo(ad.z) + bf16_hi(ao.z) * bf16_hi(ad.z);
    acc += bf16_lo(ao.w) * bf16_lo(ad.w) + bf16_hi(ao.w) * bf16_hi(ad.w);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (chunk == 0) delta[(b * H + h) * N + t] = acc;
  }
}

// dq[b,t,c] = bf16(acc[b,t,c]); optional colsum[c] += sum over (b,t) of the ROUNDED values (the bias
// gradient of the query projection, same definition as the fused column sums of dk / dv).
// Block = 256 threads: thread (rl, cg) walks rows rl, rl+R, ... of its row chunk for column group cg.
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, float* __restrict__ colsum,
                       int64_t rows_total, int N, int cols, int64_t lddq, int64_t bsdq, int rows_per_block) {
  const int groups = cols / 8;                        // 8-column groups
  const int rlanes = 256 / groups > 0 ? 256 / groups : 1;
  const int cg = threadIdx.x % groups, rl = threadIdx.x / groups;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < rlanes) {
    for (int64_t r = r0 + rl; r < r0 + rows_per_block && r < rows_total; r += rlanes) {
      const float4 a = *reinterpret_cast<const float4*>(acc + r * cols + cg * 8);
      const float4 c = *reinterpret_cast<const float4*>(acc + r * cols + cg * 8 + 4);
      uint4 q;
      q.x = pack_bf16(a.x, a.y); q.y = pack_bf16(a.z, a.w);
      q.z = pack_bf16(c.x, c.y); q.w = pack_bf16(c.z, c.w);
      const int64_t b = r / N;
      const int t = static_cast<int>(r % N);
      *reinterpret_cast<uint4*>(dq + b * bsdq + t * lddq + cg * 8) = q;
      cs[0] += bf16_lo(q.x); cs[1] += bf16_hi(q.x); cs[2] += bf16_lo(q.y); cs[3] += bf16_hi(q.y);
      cs[4] += bf16_lo(q.z); cs[5] += bf16_hi(q.z); cs[6] += bf16_lo(q.w); cs[7] += bf16_hi(q.w);
    }
    if (colsum != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(colsum + cg * 8 + k, cs[k]);
    }
  }
}

}  // namespace

int launch_attention_fwd_stream(const AttnArgs& a, cudaStream_t s) {
  Fwd2Dev p;
  p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk;
  p.QT = (a.Nq + TQ - 1) / TQ;
  p.NB = (a.Nk + BK - 1) / BK;
  p.tiles = static_cast<int>(a.B * a.H * p.QT);
  p.scale_log2 = a.scale * LOG2E;
  p.lse = a.lse;
  const int cols = a.H * DH;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = make_tmap_bnd(&tmQ, a.q, cols, a.Nq, a.B, a.ldq, a.bsq, TQ))) return rc;
  if ((rc = make_tmap_bnd(&tmK, a.k, cols, a.Nk, a.B, a.ldk, a.bsk, BK))) return rc;
  if ((rc = make_tmap_bnd(&tmV, a.v, cols, a.Nk, a.B, a.ldv, a.bsv, BK))) return rc;
  if ((rc = make_tmap_bnd(&tmO, a.o, cols, a.Nq, a.B, a.ldo, a.bso, TQ))) return rc;
  const int sms = num_sms();
  const int grid = p.tiles < sms ? p.tiles : sms;
  int var = F2_DEFAULT_VAR;
  { const char* e = getenv("BV_ATTN_SM"); if (e) var = atoi(e); }
  // BV_ATTN_SM + 100 selects the variant with P kept in tensor memory (TSP)
#define F2_LAUNCH(V, T)                                                                                         \
  case V + (T ? 100 : 0):                                                                                       \
    rc = check_cuda(cudaFuncSetAttribute(attn_fwd_stream_kernel<V, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         F2L<T>::SMEM), "cudaFuncSetAttribute(attn_fwd_stream)");                \
    if (rc) return rc;                                                                                          \
    attn_fwd_stream_kernel<V, T><<<grid, F2_THREADS, F2L<T>::SMEM, s>>>(tmQ, tmK, tmV, tmO, p);                 \
    break;
  switch (var) {
    F2_LAUNCH(0, false) F2_LAUNCH(5, false) F2_LAUNCH(8, false) F2_LAUNCH(9, false) F2_LAUNCH(12, false)
    F2_LAUNCH(8, true) F2_LAUNCH(9, true) F2_LAUNCH(12, true)
    default: set_error("BV_ATTN_SM=%d: unknown softmax variant", var); return BV_ERR_INVALID;
  }
#undef F2_LAUNCH
  return check_cuda(cudaGetLastError(), "attn_fwd_stream_kernel launch");
}

int launch_attention_bwd_stream(const AttnBwdArgs& g, cudaStream_t s) {
  const AttnArgs& a = g.f;
  if (g.dq_accum == nullptr || g.delta == nullptr) {
    set_error("bv_attention_bwd: sequences longer than 256 need the dq_accum [B,Nq,H*64] and delta "
              "[B,H,Nq] fp32 workspaces");
    return BV_ERR_INVALID;
  }
  if (a.H * DH > 2048 || (a.H * DH) % 8) { set_error("bv_attention_bwd: H*64 must be <= 2048"); return BV_ERR_INVALID; }
  Bwd2Dev p;
  p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk;
  p.QT = (a.Nq + TQ - 1) / TQ;
  p.KT = (a.Nk + TQ - 1) / TQ;
  p.groups = static_cast<int>(a.B * a.H * p.KT);
  p.scale = a.scale;
  p.scale_log2 = a.scale * LOG2E;
  p.lse = a.lse;
  p.delta = g.delta;
  p.dk_colsum = g.dk_colsum; p.dv_colsum = g.dv_colsum;
  { const char* e = getenv("BV_BWD_VARIANT"); p.variant = e ? atoi(e) : 0; }
  const int cols = a.H * DH;
  int rc;
  // delta = rowsum(O o dO)
  {
    const int64_t head_rows = a.B * a.Nq * a.H;
    int64_t blocks = (head_rows * 8 + 255) / 256;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    attn_delta_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
        reinterpret_cast<const bf16*>(a.o), reinterpret_cast<const bf16*>(g.d_o), g.delta, a.B, a.H, a.Nq,
        a.ldo, a.bso, g.lddo, g.bsdo);
    if ((rc = check_cuda(cudaGetLastError(), "attn_delta_kernel launch"))) return rc;
  }
  if ((rc = check_cuda(cudaMemsetAsync(g.dq_accum, 0, sizeof(float) * a.B * a.Nq * cols, s),
                       "cudaMemsetAsync(dq_accum)"))) return rc;
  CUtensorMap tmQ, tmK, tmV, tmdO, tmdQacc, tmdK, tmdV;
  if ((rc = make_tmap_bnd(&tmQ, a.q, cols, a.Nq, a.B, a.ldq, a.bsq, TQ))) return rc;
  if ((rc = make_tmap_bnd(&tmK, a.k, cols, a.Nk, a.B, a.ldk, a.bsk, TQ))) return rc;
  if ((rc = make_tmap_bnd(&tmV, a.v, cols, a.Nk, a.B, a.ldv, a.bsv, TQ))) return rc;
  if ((rc = make_tmap_bnd(&tmdO, g.d_o, cols, a.Nq, a.B, g.lddo, g.bsdo, TQ))) return rc;
  if ((rc = make_tmap_bnd(&tmdK, g.dk, cols, a.Nk, a.B, g.lddk, g.bsdk, TQ))) return rc;
  if ((rc = make_tmap_bnd(&tmdV, g.dv, cols, a.Nk, a.B, g.lddv, g.bsdv, TQ))) return rc;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(a.Nq), static_cast<uint64_t>(a.B)};
    uint64_t strides[2] = {static_cast<uint64_t>(cols) * 4, static_cast<uint64_t>(a.Nq) * cols * 4};
    uint32_t box[3] = {32, TQ, 1};
    if ((rc = make_tmap(&tmdQacc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g.dq_accum, dims, strides, box, true))) return rc;
  }
  rc = check_cuda(cudaFuncSetAttribute(attn_bwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       B2_SMEM), "cudaFuncSetAttribute(attn_bwd_stream)");
  if (rc) return rc;
  const int sms = num_sms();
  const int grid = p.groups < sms ? p.groups : sms;
  attn_bwd_stream_kernel<<<grid, B2_THREADS, B2_SMEM, s>>>(tmQ, tmK, tmV, tmdO, tmdQacc, tmdK, tmdV, p);
  if ((rc = check_cuda(cudaGetLastError(), "attn_bwd_stream_kernel launch"))) return rc;
  {
    const int64_t rows = a.B * a.Nq;
    const int rows_per_block = 256;
    const int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
    attn_dq_convert_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
        g.dq_accum, reinterpret_cast<bf16*>(g.dq), g.dq_colsum, rows, a.Nq, cols, g.lddq, g.bsdq, rows_per_block);
    rc = check_cuda(cudaGetLastError(), "attn_dq_convert_kernel launch");
  }
  return rc;
}

}  // namespace bv

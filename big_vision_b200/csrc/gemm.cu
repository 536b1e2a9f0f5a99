// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
//
// A and B are bf16 and each may be stored K-major (rows = M/N, K contiguous) or
// MN-major (rows = K, M/N contiguous); both go HBM -> smem by TMA (128B swizzle)
// and smem -> tensor core by UMMA descriptors, accumulating fp32 in TMEM.  This
// one kernel serves every dense contraction of the hot path:
//   forward  Y  = X  W      A = X  (K-major),  B = W  [K,N] (MN-major)    K1,K4,K6,K7,K8,K11
//   dgrad    dX = dY W^T    A = dY (K-major),  B = W  [K,N] (K-major)
//   wgrad    dW = X^T dY    A = X  (MN-major), B = dY (MN-major), split-K, fp32 reduce-add
// (reference call sites: flax Dense/DenseGeneral under models/vit.py:72-77,93-98,
//  176-178,212-214,261,272; models/mlp_mixer.py:35-37,72,82).
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issue,
// warps 2..5 = epilogue (TMEM -> regs -> swizzled smem -> TMA store / reduce).
// Two TMEM accumulators so the epilogue of tile i overlaps the mainloop of i+1.
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bv {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int OUT_BUF_BYTES = BM * 128;      // 128 rows x 128 B
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;   // 512 or 256 (power of two)
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES + 2 * OUT_BUF_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256 + 1024;  // + barriers + align slack
};

struct GemmDev {
  int M, N, K;
  int num_m_tiles, num_n_tiles, total_tiles;
  int kblocks_total, kblocks_per_split;
  int a_mn, b_mn;        // 1 = MN-major
  int epi;
  int reduce_out;        // 1 = TMA reduce-add into D (split-K / grad accumulation)
  float alpha;
  const float* bias;
  const bf16* aux;
  long long ldaux;
  int aux_row_mod;
};

template <int BN, bool OUT_F32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2,
            const GemmDev p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);

  const uint32_t out_buf = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = base + C::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + C::BAR_OFFSET + 8 * (2 * C::STAGES + 4));

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (p.epi == EPI_BIAS_GELU) tma_prefetch_desc(&tmD2);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto decode_tile = [&](int tile, int& m0, int& n0, int& kb0, int& kb1) {
    int n_tile = tile % p.num_n_tiles;
    int rest = tile / p.num_n_tiles;
    int m_tile = rest % p.num_m_tiles;
    int split = rest / p.num_m_tiles;
    m0 = m_tile * BM;
    n0 = n_tile * BN;
    kb0 = split * p.kblocks_per_split;
    kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
  };

  if (warp_idx == 0) {
    // ========================= TMA producer =========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int m0, n0, kb0, kb1;
        decode_tile(tile, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_s = base + stage * C::STAGE_BYTES;
          const uint32_t b_s = a_s + A_STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
          const int k0 = kb * BK;
          if (p.a_mn) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(a_s + j * 8192, &tmA, full_bar(stage), m0 + 64 * j, k0);
          } else {
            tma_load_2d(a_s, &tmA, full_bar(stage), k0, m0);
          }
          if (p.b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(b_s + j * 8192, &tmB, full_bar(stage), n0 + 64 * j, k0);
          } else {
            tma_load_2d(b_s, &tmB, full_bar(stage), k0, n0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ========================= MMA issuer =========================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int m0, n0, kb0, kb1;
        decode_tile(tile, m0, n0, kb0, kb1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_s = base + stage * C::STAGE_BYTES;
          const uint32_t b_s = a_s + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = umma_smem_desc_sw128(a_s + k * a_kstep, a_lbo, 1024u);
            const uint64_t bdesc = umma_smem_desc_sw128(b_s + k * b_kstep, b_lbo, 1024u);
            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));   // smem slot free once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));       // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ========================= epilogue =========================
    const int ep_tid = threadIdx.x - 64;
    const int lane_grp = warp_idx & 3;              // TMEM lanes this warp may access
    const int row = lane_grp * 32 + lane;           // row within the tile == TMEM lane
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const bool dual = (p.epi == EPI_BIAS_GELU);
    constexpr int SUB_PER_FLUSH = OUT_F32 ? 1 : 2;  // 32-col sub-chunks per 128B staging row
    constexpr int CH = OUT_F32 ? 32 : 64;           // columns per TMA store
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t flush = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int m0, n0, kb0, kb1;
      decode_tile(tile, m0, n0, kb0, kb1);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int grow = m0 + row;
      const bool row_ok = grow < p.M;
      const bf16* aux_row = nullptr;
      if (p.aux != nullptr && row_ok) {
        long long ar = p.aux_row_mod > 0 ? (grow % p.aux_row_mod) : grow;
        aux_row = p.aux + ar * p.ldaux;
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + acc * BN;

#pragma unroll 1
      for (int sc = 0; sc < BN / 32; ++sc) {
        const int sub = sc % SUB_PER_FLUSH;
        const uint32_t buf = dual ? out_buf : out_buf + (flush & 1u) * OUT_BUF_BYTES;
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + sc * 32, r);
        tmem_ld_wait();
        if (sc == BN / 32 - 1) {
          // accumulator fully drained into registers -> hand TMEM back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (sub == 0) {
          // make sure the TMA store that last used this staging buffer has read it
          if (ep_tid == 0) {
            if (dual) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
          }
          named_bar_sync(1, EPI_THREADS);
        }
        const int ncol0 = n0 + sc * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {       // 8 columns per group
          const int nc = ncol0 + g * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]) * p.alpha;
          const bool col_ok = nc < p.N;      // N % 8 == 0
          if (p.bias != nullptr && col_ok) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + nc));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + nc + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          float a[8];
          bool have_aux = false;
          if (aux_row != nullptr && col_ok) {
            const uint4 q = *reinterpret_cast<const uint4*>(aux_row + nc);
            a[0] = bf16_lo(q.x); a[1] = bf16_hi(q.x); a[2] = bf16_lo(q.y); a[3] = bf16_hi(q.y);
            a[4] = bf16_lo(q.z); a[5] = bf16_hi(q.z); a[6] = bf16_lo(q.w); a[7] = bf16_hi(q.w);
            have_aux = true;
          }
          float v2[8];
          if (p.epi == EPI_BIAS_GELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v2[i] = round_bf16(v[i]);
              v[i] = gelu_tanh(v2[i]);
            }
          } else if (p.epi == EPI_BIAS_RESID) {
            if (have_aux) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = (OUT_F32 ? v[i] : round_bf16(v[i])) + a[i];
            }
          } else if (p.epi == EPI_DGELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = have_aux ? v[i] * gelu_tanh_grad(a[i]) : 0.f;
          }
          if (OUT_F32) {
            const uint32_t p0 = static_cast<uint32_t>(g * 2), p1 = p0 + 1;
            const uint32_t a0 = buf + row * 128 + ((p0 ^ sw) << 4);
            const uint32_t a1 = buf + row * 128 + ((p1 ^ sw) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "f"(v[0]),
                         "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "f"(v[4]),
                         "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
          } else {
            const uint32_t piece = static_cast<uint32_t>(sub * 4 + g);
            const uint32_t a0 = buf + row * 128 + ((piece ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0),
                         "r"(pack_bf16(v[0], v[1])), "r"(pack_bf16(v[2], v[3])),
                         "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7])) : "memory");
            if (dual) {
              const uint32_t a2 = a0 + OUT_BUF_BYTES;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a2),
                           "r"(pack_bf16(v2[0], v2[1])), "r"(pack_bf16(v2[2], v2[3])),
                           "r"(pack_bf16(v2[4], v2[5])), "r"(pack_bf16(v2[6], v2[7])) : "memory");
            }
          }
        }
        if (sub == SUB_PER_FLUSH - 1) {
          fence_proxy_async();
          named_bar_sync(1, EPI_THREADS);
          if (ep_tid == 0) {
            const int c0 = n0 + (sc / SUB_PER_FLUSH) * CH;
            if (c0 < p.N) {
              if (p.reduce_out) tma_reduce_add_2d(&tmD, buf, c0, m0);
              else tma_store_2d(&tmD, buf, c0, m0);
              if (dual) tma_store_2d(&tmD2, buf + OUT_BUF_BYTES, c0, m0);
            }
            tma_store_commit();
          }
          ++flush;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (ep_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, bool OUT_F32>
int launch_cfg(const GemmArgs& g, cudaStream_t stream) {
  using C = Cfg<BN>;
  CUtensorMap tmA, tmB, tmD, tmD2;
  int rc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if (g.a_mn) rc = make_tmap_2d(&tmA, bf, g.A, g.M, g.K, g.lda * 2, 64, 64);
  else        rc = make_tmap_2d(&tmA, bf, g.A, g.K, g.M, g.lda * 2, 64, BM);
  if (rc) return rc;
  if (g.b_mn) rc = make_tmap_2d(&tmB, bf, g.B, g.N, g.K, g.ldb * 2, 64, 64);
  else        rc = make_tmap_2d(&tmB, bf, g.B, g.K, g.N, g.ldb * 2, 64, BN);
  if (rc) return rc;
  if (OUT_F32) rc = make_tmap_2d(&tmD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, g.D, g.N, g.M, g.ldd * 4, 32, BM);
  else         rc = make_tmap_2d(&tmD, bf, g.D, g.N, g.M, g.ldd * 2, 64, BM);
  if (rc) return rc;
  tmD2 = tmD;
  if (g.epi == EPI_BIAS_GELU) {
    rc = make_tmap_2d(&tmD2, bf, g.D2, g.N, g.M, g.ldd2 * 2, 64, BM);
    if (rc) return rc;
  }

  GemmDev p;
  p.M = (int)g.M; p.N = (int)g.N; p.K = (int)g.K;
  p.num_m_tiles = (int)((g.M + BM - 1) / BM);
  p.num_n_tiles = (int)((g.N + BN - 1) / BN);
  p.kblocks_total = (int)((g.K + BK - 1) / BK);
  int splits = g.splits;
  const int sms = num_sms();
  if (splits <= 0) {            // auto: fill the machine when the output grid is small
    splits = 1;
    if (g.reduce_out) {
      int tiles = p.num_m_tiles * p.num_n_tiles;
      if (tiles < sms) splits = sms / tiles;
    }
  }
  if (splits > p.kblocks_total) splits = p.kblocks_total;
  if (splits < 1) splits = 1;
  if (splits > 1 && !g.reduce_out) {
    set_error("bv_gemm: split-K requires reduce_out=1");
    return BV_ERR_INVALID;
  }
  p.kblocks_per_split = (p.kblocks_total + splits - 1) / splits;
  splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;
  p.total_tiles = p.num_m_tiles * p.num_n_tiles * splits;
  p.a_mn = g.a_mn; p.b_mn = g.b_mn; p.epi = g.epi; p.reduce_out = g.reduce_out;
  p.alpha = g.alpha;
  p.bias = g.bias;
  p.aux = reinterpret_cast<const bf16*>(g.aux);
  p.ldaux = g.ldaux;
  p.aux_row_mod = g.aux_row_mod;

  auto kern = gemm_kernel<BN, OUT_F32>;
  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES), "cudaFuncSetAttribute(gemm)");
    if (rc) return rc;
    attr_set = true;
  }
  int grid = p.total_tiles < sms ? p.total_tiles : sms;
  kern<<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, tmD, tmD2, p);
  return check_cuda(cudaGetLastError(), "gemm_kernel launch");
}

}  // namespace

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) { set_error("bv_gemm: empty problem"); return BV_ERR_INVALID; }
  if (g.N % 8 != 0) { set_error("bv_gemm: N=%lld must be a multiple of 8", (long long)g.N); return BV_ERR_INVALID; }
  if (g.M > 0x7fffffffLL || g.N > 0x7fffffffLL || g.K > 0x7fffffffLL) {
    set_error("bv_gemm: dimension exceeds int32"); return BV_ERR_INVALID;
  }
  if (g.epi < EPI_NONE || g.epi > EPI_DGELU) { set_error("bv_gemm: bad epilogue %d", g.epi); return BV_ERR_INVALID; }
  if ((g.epi == EPI_BIAS_RESID || g.epi == EPI_DGELU) && g.aux == nullptr) {
    set_error("bv_gemm: epilogue %d needs aux", g.epi); return BV_ERR_INVALID;
  }
  if (g.aux != nullptr && ((reinterpret_cast<uintptr_t>(g.aux) & 15) || (g.ldaux % 8))) {
    set_error("bv_gemm: aux must be 16B aligned with ldaux %% 8 == 0"); return BV_ERR_INVALID;
  }
  if (g.bias != nullptr && (reinterpret_cast<uintptr_t>(g.bias) & 15)) {
    set_error("bv_gemm: bias must be 16B aligned"); return BV_ERR_INVALID;
  }
  if (g.epi == EPI_BIAS_GELU && (g.out_dtype != DT_BF16 || g.D2 == nullptr || g.reduce_out)) {
    set_error("bv_gemm: BIAS_GELU needs bf16 output, D2 and no reduce"); return BV_ERR_INVALID;
  }
  const bool f32 = (g.out_dtype == DT_F32);
  if (!f32 && g.out_dtype != DT_BF16) { set_error("bv_gemm: bad out dtype"); return BV_ERR_INVALID; }
  int bn = g.block_n;
  if (bn == 0) {
    const long long m_tiles = (g.M + BM - 1) / BM;
    bn = (g.N % 256 == 0 || g.N >= 2048) ? 256 : 128;
    if (bn == 256 && m_tiles * ((g.N + 255) / 256) < num_sms() / 2 && !g.reduce_out) bn = 128;
  }
  if (bn == 256) return f32 ? launch_cfg<256, true>(g, stream) : launch_cfg<256, false>(g, stream);
  if (bn == 128) return f32 ? launch_cfg<128, true>(g, stream) : launch_cfg<128, false>(g, stream);
  set_error("bv_gemm: block_n must be 0, 128 or 256");
  return BV_ERR_INVALID;
}

}  // namespace bv

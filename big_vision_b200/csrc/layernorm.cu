// LayerNorm forward / backward (K3).  Reference: flax.linen.LayerNorm as called at
// models/vit.py:92,103,160,181 and models/mlp_mixer.py:48,53,79 -- eps = 1e-6,
// statistics in fp32 with the "fast variance" form var = max(E[x^2] - E[x]^2, 0).
//
// HBM-bound: one warp per row, 16-byte vector loads, the row stays in registers
// between the statistics pass and the normalise pass (one read + one write of
// [rows, d]).  The backward also folds in the residual-branch gradient and the
// column sums that are the bias gradients of the GEMMs upstream of the residual
// stream, so those need no extra pass over HBM.
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bv {
namespace {

__device__ __forceinline__ void load8(const void* base, int dtype, int64_t elem_off, float (&v)[8]) {
  if (dtype == DT_BF16) {
    const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + elem_off);
    v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
    v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
__device__ __forceinline__ void store8(void* base, int dtype, int64_t elem_off, const float (&v)[8]) {
  if (dtype == DT_BF16) {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]);
    q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + elem_off) = q;
  } else {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;

template <int NCH>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const void* __restrict__ x, int x_dt, const float* __restrict__ scale,
              const float* __restrict__ bias, void* __restrict__ y, int y_dt,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int d,
              float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nchunks = d >> 3;
  float v[NCH][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      load8(x, x_dt, row * d + c * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1 += v[i][j]; s2 += v[i][j] * v[i][j]; }
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float inv_d = 1.0f / static_cast<float>(d);
  const float mean = s1 * inv_d;
  const float var = fmaxf(s2 * inv_d - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
      float g[8], b[8], o[8];
      load8f(scale + c * 8, g);
      load8f(bias + c * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
      store8(y, y_dt, row * d + c * 8, o);
    }
  }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * scale
// dscale += sum_rows dy * xhat ; dbias += sum_rows dy ; dx_colsum += sum_rows dx
template <int NCH>
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_kernel(const void* __restrict__ dy, int dy_dt, const void* __restrict__ x, int x_dt,
              const float* __restrict__ scale, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const void* __restrict__ dres,
              void* __restrict__ dx, int dx_dt, float* __restrict__ dscale,
              float* __restrict__ dbias, float* __restrict__ dx_colsum, int64_t rows, int d) {
  extern __shared__ float red[];   // [3][d]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nchunks = d >> 3;
  const bool want_cs = dx_colsum != nullptr;
  for (int i = threadIdx.x; i < 3 * d; i += LN_THREADS) red[i] = 0.f;
  __syncthreads();

  float acc_g[NCH][8], acc_b[NCH][8], acc_c[NCH][8];
  float g[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc_g[i][j] = 0.f; acc_b[i][j] = 0.f; acc_c[i][j] = 0.f; g[i][j] = 0.f; }
    if (c < nchunks) load8f(scale + c * 8, g[i]);
  }
  const float inv_d = 1.0f / static_cast<float>(d);
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * LN_WARPS;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * LN_WARPS + warp; row < rows;
       row += warp_stride) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float xh[NCH][8], gy[NCH][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float xv[8], dv[8];
        load8(x, x_dt, row * d + c * 8, xv);
        load8(dy, dy_dt, row * d + c * 8, dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xv[j] - mean) * rstd;
          gy[i][j] = dv[j] * g[i][j];
          c1 += gy[i][j];
          c2 += gy[i][j] * xh[i][j];
          acc_g[i][j] += dv[j] * xh[i][j];
          acc_b[i][j] += dv[j];
        }
      }
    }
    c1 = warp_sum(c1) * inv_d;
    c2 = warp_sum(c2) * inv_d;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (gy[i][j] - c1 - xh[i][j] * c2);
        if (dres != nullptr) {
          float r[8];
          load8(dres, dx_dt, row * d + c * 8, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        if (dx_dt == DT_BF16) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = round_bf16(o[j]);
        }
        store8(dx, dx_dt, row * d + c * 8, o);
        if (want_cs) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc_c[i][j] += o[j];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&red[c * 8 + j], acc_g[i][j]);
        atomicAdd(&red[d + c * 8 + j], acc_b[i][j]);
        if (want_cs) atomicAdd(&red[2 * d + c * 8 + j], acc_c[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += LN_THREADS) {
    if (dscale) atomicAdd(dscale + i, red[i]);
    if (dbias) atomicAdd(dbias + i, red[d + i]);
    if (want_cs) atomicAdd(dx_colsum + i, red[2 * d + i]);
  }
}

int check_ln(int64_t rows, int d, const char* who) {
  if (rows < 0 || d <= 0 || d % 8 != 0 || d > 2048) {
    set_error("%s: need rows >= 0 and d %% 8 == 0, d <= 2048 (got rows=%lld d=%d)", who,
              (long long)rows, d);
    return BV_ERR_INVALID;
  }
  return BV_OK;
}

}  // namespace

int launch_layernorm_fwd(const void* x, int x_dt, const float* scale, const float* bias, void* y,
                         int y_dt, float* mean, float* rstd, int64_t rows, int d, float eps,
                         cudaStream_t s) {
  int rc = check_ln(rows, d, "bv_layernorm_fwd");
  if (rc) return rc;
  if (rows == 0) return BV_OK;
  const int nch = (d / 8 + 31) / 32;
  const unsigned grid = static_cast<unsigned>((rows + LN_WARPS - 1) / LN_WARPS);
#define LN_FWD_CASE(N)                                                                        \
  case N:                                                                                     \
    ln_fwd_kernel<N><<<grid, LN_THREADS, 0, s>>>(x, x_dt, scale, bias, y, y_dt, mean, rstd,   \
                                                 rows, d, eps);                               \
    break;
  switch (nch) {
    LN_FWD_CASE(1) LN_FWD_CASE(2) LN_FWD_CASE(3) LN_FWD_CASE(4)
    LN_FWD_CASE(5) LN_FWD_CASE(6) LN_FWD_CASE(7) LN_FWD_CASE(8)
    default: set_error("bv_layernorm_fwd: d too large"); return BV_ERR_INVALID;
  }
#undef LN_FWD_CASE
  return check_cuda(cudaGetLastError(), "ln_fwd_kernel launch");
}

int launch_layernorm_bwd(const void* dy, int dy_dt, const void* x, int x_dt, const float* scale,
                         const float* mean, const float* rstd, const void* dres, void* dx,
                         int dx_dt, float* dscale, float* dbias, float* dx_colsum, int64_t rows,
                         int d, cudaStream_t s) {
  int rc = check_ln(rows, d, "bv_layernorm_bwd");
  if (rc) return rc;
  if (rows == 0) return BV_OK;
  const int nch = (d / 8 + 31) / 32;
  int64_t blocks = (rows + LN_WARPS - 1) / LN_WARPS;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 2;   // 2 resident blocks/SM (register-bound)
  if (blocks > cap) blocks = cap;
  const size_t smem = 3 * static_cast<size_t>(d) * sizeof(float);
#define LN_BWD_CASE(N)                                                                         \
  case N:                                                                                      \
    ln_bwd_kernel<N><<<(unsigned)blocks, LN_THREADS, smem, s>>>(                               \
        dy, dy_dt, x, x_dt, scale, mean, rstd, dres, dx, dx_dt, dscale, dbias, dx_colsum, rows, d); \
    break;
  switch (nch) {
    LN_BWD_CASE(1) LN_BWD_CASE(2) LN_BWD_CASE(3) LN_BWD_CASE(4)
    LN_BWD_CASE(5) LN_BWD_CASE(6)
    default:
      set_error("bv_layernorm_bwd: d=%d > 1536 not supported yet", d);
      return BV_ERR_UNSUPPORTED;
  }
#undef LN_BWD_CASE
  return check_cuda(cudaGetLastError(), "ln_bwd_kernel launch");
}

}  // namespace bv

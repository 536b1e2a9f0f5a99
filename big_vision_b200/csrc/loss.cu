// Loss kernels.
//  * SigLIP pairwise sigmoid loss (K14): trainers/proj/image_text/siglip.py:291-306,
//    explicit per-device form trainers/proj/image_text/_deprecated_contrastive.py:117-141.
//      x_ij   = (zimg_i . ztxt_j) * exp(t') + b
//      loglik = log_sigmoid(+x_ij) on the positive diagonal, log_sigmoid(-x_ij) elsewhere
//      loss   = (1/B) sum_i sum_j -loglik_ij            (B = GLOBAL batch, siglip.py:306)
//    The dot products come from the tcgen05 GEMM; this kernel fuses scale+bias, the
//    loss reduction and d loss/d dot (written as the bf16 operand of the two gradient
//    GEMMs) plus the scalar gradients of t' and b in one pass over the [n, B] slab.
//  * sigmoid_xent / softmax_xent (K15): utils.py:236-243, 276-281.
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bv {
namespace {

__device__ __forceinline__ float log_sigmoid(float y) {
  // log sigma(y) = min(y, 0) - log1p(exp(-|y|))   (stable; matches jax.nn.log_sigmoid)
  return fminf(y, 0.f) - log1pf(__expf(-fabsf(y)));
}
__device__ __forceinline__ float sigmoid(float y) { return 1.f / (1.f + __expf(-y)); }

__device__ __forceinline__ void block_reduce3(float& a, float& b, float& c, float* sh) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) { sh[warp] = a; sh[32 + warp] = b; sh[64 + warp] = c; }
  __syncthreads();
  if (warp == 0) {
    a = lane < nw ? sh[lane] : 0.f;
    b = lane < nw ? sh[32 + lane] : 0.f;
    c = lane < nw ? sh[64 + lane] : 0.f;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  }
}

__global__ void __launch_bounds__(256)
siglip_loss_kernel(const float* __restrict__ dots, int64_t n, int64_t B, int64_t ld,
                   int64_t row_offset, const float* __restrict__ t_param,
                   const float* __restrict__ b_param, float inv_B, bf16* __restrict__ G,
                   int64_t ldg, float* __restrict__ loss, float* __restrict__ dt,
                   float* __restrict__ db, float* __restrict__ partials) {
  __shared__ float sh[96];
  const float t = __expf(t_param[0]);
  const float bias = b_param ? b_param[0] : 0.f;
  float l_acc = 0.f, t_acc = 0.f, b_acc = 0.f;
  const int64_t groups = B / 4;
  const int64_t total = n * groups;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = idx / groups;
    const int64_t j0 = (idx % groups) * 4;
    const float4 dv = *reinterpret_cast<const float4*>(dots + i * ld + j0);
    const float d[4] = {dv.x, dv.y, dv.z, dv.w};
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float sgn = (j0 + e == row_offset + i) ? 1.f : -1.f;
      const float x = d[e] * t + bias;
      l_acc -= log_sigmoid(sgn * x);
      const float gx = -sgn * sigmoid(-sgn * x) * inv_B;    // d loss / d x
      b_acc += gx;
      t_acc += gx * d[e] * t;                                // d x / d t' = dot * exp(t')
      g[e] = gx * t;                                         // d loss / d dot
    }
    uint2 q;
    q.x = pack_bf16(g[0], g[1]);
    q.y = pack_bf16(g[2], g[3]);
    *reinterpret_cast<uint2*>(G + i * ldg + j0) = q;
  }
  l_acc *= inv_B;
  block_reduce3(l_acc, t_acc, b_acc, sh);
  if (threadIdx.x == 0) {
    if (partials != nullptr) {
      // deterministic mode: one partial per block, summed in a fixed order by finish_sums_kernel
      partials[blockIdx.x] = l_acc;
      partials[gridDim.x + blockIdx.x] = t_acc;
      partials[2 * gridDim.x + blockIdx.x] = b_acc;
    } else {
      atomicAdd(loss, l_acc);
      if (dt) atomicAdd(dt, t_acc);
      if (db) atomicAdd(db, b_acc);
    }
  }
}

// Fixed-order finishing pass of the deterministic mode: out[k] += sum_i part[k*count + i].  One
// block; thread t sums elements t, t+256, ... in index order, then a fixed shared-memory tree.  The
// result depends only on (count, values), never on which block of the producer finished first --
// XLA's reductions are run-to-run deterministic and so is this path.
__global__ void __launch_bounds__(256)
finish_sums_kernel(const float* __restrict__ part, int64_t count, float* out0, float* out1,
                   float* out2) {
  __shared__ float sh[256];
  float* outs[3] = {out0, out1, out2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (outs[k] == nullptr) continue;          // uniform across the block
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < count; i += 256) acc += part[k * count + i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (static_cast<int>(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) outs[k][0] += sh[0];
    __syncthreads();
  }
}

// One direction of the softmax (CLIP) contrastive loss, _deprecated_contrastive.py:80-101:
//   x_ij = dots_ij * exp(t'),  loss_i = logsumexp_j x_ij - x_i,pos(i),  pos(i) = row_offset + i
//   loss += weight/global_B * sum_i loss_i ;  G_ij = weight/global_B * (softmax_ij - [j == pos]) * exp(t')
//   dt'  += sum_ij (G_ij / exp(t')) * x_ij ;  ncorrect += #[argmax_j x_ij == pos(i)]   (first max wins)
// One warp per row; per-row partials (loss, dt', correct) go to `rows_ws` [3, n] and are summed in a
// fixed order by finish_sums_kernel (deterministic like the other losses).
__global__ void __launch_bounds__(256)
softmax_contrastive_kernel(const float* __restrict__ dots, int64_t n, int64_t B, int64_t ld, int64_t row_offset,
                           const float* __restrict__ t_param, float scale, bf16* __restrict__ G, int64_t ldg,
                           float* __restrict__ rows_ws) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float t = __expf(t_param[0]);
  const float* d = dots + row * ld;
  const int64_t pos = row_offset + row;
  float mx = -INFINITY;
  int64_t arg = 0;
  for (int64_t c = lane; c < B; c += 32) {
    const float x = d[c] * t;
    if (x > mx) { mx = x; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const int64_t oarg = __shfl_xor_sync(0xffffffffu, arg, o);
    if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
  }
  float se = 0.f;
  for (int64_t c = lane; c < B; c += 32) se += __expf(d[c] * t - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  float dt_acc = 0.f;
  for (int64_t c = lane; c < B; c += 32) {
    const float x = d[c] * t;
    const float gx = (__expf(x - lse) - (c == pos ? 1.f : 0.f)) * scale;     // d loss / d x
    dt_acc += gx * x;
    G[row * ldg + c] = __float2bfloat16_rn(gx * t);                            // d loss / d dot
  }
  dt_acc = warp_sum(dt_acc);
  if (lane == 0) {
    const float xpos = (pos >= 0 && pos < B) ? d[pos] * t : 0.f;
    rows_ws[row] = (lse - xpos) * scale;
    rows_ws[n + row] = dt_acc;
    rows_ws[2 * n + row] = (arg == pos) ? 1.f : 0.f;
  }
}

// one warp per row
__global__ void __launch_bounds__(256)
sigmoid_xent_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                    float* __restrict__ loss, float* __restrict__ dlogits,
                    float* __restrict__ row_loss, int64_t n, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float inv_n = 1.f / static_cast<float>(n);
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float x = logits[row * C + c], y = labels[row * C + c];
    acc -= y * log_sigmoid(x) + (1.f - y) * log_sigmoid(-x);
    if (dlogits) dlogits[row * C + c] = (sigmoid(x) - y) * inv_n;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (row_loss != nullptr) row_loss[row] = acc * inv_n; else atomicAdd(loss, acc * inv_n);
  }
}

__global__ void __launch_bounds__(256)
softmax_xent_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                    float* __restrict__ loss, float* __restrict__ dlogits,
                    float* __restrict__ row_loss, int64_t n, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const float inv_n = 1.f / static_cast<float>(n);
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, logits[row * C + c]);
  mx = warp_max(mx);
  float se = 0.f, sy = 0.f, sxy = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float x = logits[row * C + c] - mx, y = labels[row * C + c];
    se += __expf(x); sy += y; sxy += y * x;
  }
  se = warp_sum(se); sy = warp_sum(sy); sxy = warp_sum(sxy);
  const float lse = logf(se);
  // -sum y (x - lse) = lse * sum(y) - sum(y x)
  if (lane == 0) {
    const float rl = (lse * sy - sxy) * inv_n;
    if (row_loss != nullptr) row_loss[row] = rl; else atomicAdd(loss, rl);
  }
  if (dlogits) {
    for (int c = lane; c < C; c += 32) {
      const float x = logits[row * C + c] - mx, y = labels[row * C + c];
      dlogits[row * C + c] = (__expf(x - lse) * sy - y) * inv_n;
    }
  }
}

}  // namespace

int launch_siglip_loss_ew(const float* dots, int64_t n, int64_t B, int64_t ld, int64_t row_offset,
                          const float* t_param, const float* b_param, int64_t global_B, void* G,
                          int64_t ldg, float* loss, float* dt, float* db, float* partials,
                          cudaStream_t s) {
  if (n <= 0 || B <= 0 || B % 4 || ld % 4 || ldg % 4 || global_B <= 0) {
    set_error("bv_siglip_loss: need n,B > 0 and B, ld, ldg multiples of 4");
    return BV_ERR_INVALID;
  }
  int64_t blocks = (n * (B / 4) + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > BV_LOSS_WS_FLOATS / 3) blocks = BV_LOSS_WS_FLOATS / 3;
  siglip_loss_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      dots, n, B, ld, row_offset, t_param, b_param, 1.0f / static_cast<float>(global_B),
      reinterpret_cast<bf16*>(G), ldg, loss, dt, db, partials);
  int rc = check_cuda(cudaGetLastError(), "siglip_loss_kernel launch");
  if (rc || partials == nullptr) return rc;
  finish_sums_kernel<<<1, 256, 0, s>>>(partials, blocks, loss, dt, db);
  return check_cuda(cudaGetLastError(), "finish_sums_kernel launch");
}

int launch_softmax_contrastive(const float* dots, int64_t n, int64_t B, int64_t ld, int64_t row_offset,
                               const float* t_param, int64_t global_B, float weight, void* G, int64_t ldg,
                               float* loss, float* dt, float* ncorrect, float* rows_ws, cudaStream_t s) {
  if (n <= 0 || B <= 0 || global_B <= 0 || rows_ws == nullptr) {
    set_error("bv_softmax_contrastive_loss: need n, B > 0 and the [3, n] workspace");
    return BV_ERR_INVALID;
  }
  softmax_contrastive_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(
      dots, n, B, ld, row_offset, t_param, weight / static_cast<float>(global_B), reinterpret_cast<bf16*>(G), ldg,
      rows_ws);
  int rc = check_cuda(cudaGetLastError(), "softmax_contrastive_kernel launch");
  if (rc) return rc;
  finish_sums_kernel<<<1, 256, 0, s>>>(rows_ws, n, loss, dt, ncorrect);
  return check_cuda(cudaGetLastError(), "finish_sums_kernel launch");
}

int launch_sigmoid_xent(const float* logits, const float* labels, float* loss, float* dlogits,
                        float* row_loss, int64_t n, int C, cudaStream_t s) {
  if (n <= 0) return BV_OK;
  sigmoid_xent_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(logits, labels, loss, dlogits,
                                                                        row_loss, n, C);
  int rc = check_cuda(cudaGetLastError(), "sigmoid_xent_kernel launch");
  if (rc || row_loss == nullptr) return rc;
  finish_sums_kernel<<<1, 256, 0, s>>>(row_loss, n, loss, nullptr, nullptr);
  return check_cuda(cudaGetLastError(), "finish_sums_kernel launch");
}
int launch_softmax_xent(const float* logits, const float* labels, float* loss, float* dlogits,
                        float* row_loss, int64_t n, int C, cudaStream_t s) {
  if (n <= 0) return BV_OK;
  softmax_xent_kernel<<<static_cast<unsigned>((n + 7) / 8), 256, 0, s>>>(logits, labels, loss, dlogits,
                                                                        row_loss, n, C);
  int rc = check_cuda(cudaGetLastError(), "softmax_xent_kernel launch");
  if (rc || row_loss == nullptr) return rc;
  finish_sums_kernel<<<1, 256, 0, s>>>(row_loss, n, loss, nullptr, nullptr);
  return check_cuda(cudaGetLastError(), "finish_sums_kernel launch");
}

}  // namespace bv

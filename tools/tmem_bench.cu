// Micro-benchmark (bring-up aid): TMEM -> register load throughput on sm_100a for different
// tcgen05.ld shapes and warp counts, plus MUFU.EX2 / fence.proxy.async costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bench tools/tmem_bench.cu && /tmp/tmem_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <int SHAPE>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr) {   // loads SHAPE columns, returns xor
  uint32_t acc = 0;
  if constexpr (SHAPE == 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
  } else if constexpr (SHAPE == 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  }
  return acc;
}

// mode 0: x8 loads with a wait after each; 1: x32 loads; 2: 13 x8 loads then ONE wait (as the
// attention kernel does); 3: MUFU ex2 x 104; 4: fence.proxy.async x 16
template <int MODE>
__global__ void bench(long long* out, uint32_t* sink, int iters, int nwarps_active) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  float facc = threadIdx.x * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int it = 0; it < iters; ++it) {
      if constexpr (MODE == 0) {
#pragma unroll
        for (int u = 0; u < 13; ++u) acc ^= ld_cols<8>(tbase + (warp >> 2) * 104 + u * 8);
      } else if constexpr (MODE == 1) {
#pragma unroll
        for (int u = 0; u < 3; ++u) acc ^= ld_cols<32>(tbase + (warp >> 2) * 104 + u * 32);
        acc ^= ld_cols<8>(tbase + (warp >> 2) * 104 + 96);
      } else if constexpr (MODE == 2) {
        uint32_t r[13][8];
#pragma unroll
        for (int u = 0; u < 13; ++u)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]), "=r"(r[u][4]), "=r"(r[u][5]),
                         "=r"(r[u][6]), "=r"(r[u][7])
                       : "r"(tbase + (warp >> 2) * 104 + u * 8));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 13; ++u)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc ^= r[u][i];
      } else if constexpr (MODE == 3) {
#pragma unroll
        for (int u = 0; u < 104; ++u) {
          float y;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(facc - (float)u));
          facc += y;
        }
      } else if constexpr (MODE == 4) {
#pragma unroll
        for (int u = 0; u < 16; ++u) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(facc);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int MODE>
void run(const char* name, int nwarps, double bytes_per_iter) {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 148 * 256 * 4);
  const int iters = 200;
  bench<MODE><<<148, 256>>>(d_out, d_sink, iters, nwarps);
  bench<MODE><<<148, 256>>>(d_out, d_sink, iters, nwarps);
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-44s warps=%d : %8.1f cycles/iter  %7.1f B/clk/SM  (%s)\n", name, nwarps, (double)cyc / iters,
         bytes_per_iter * nwarps / ((double)cyc / iters), cudaGetErrorString(e));
  cudaFree(d_out); cudaFree(d_sink);
}

int main() {
  const double b104 = 104.0 * 32 * 4;   // bytes one warp reads per iteration (104 columns x 32 lanes)
  for (int nw : {1, 4, 8}) {
    run<0>("13 x (ld.x8 + wait)", nw, b104);
    run<1>("3 x (ld.x32 + wait) + (ld.x8 + wait)", nw, b104);
    run<2>("13 x ld.x8, one wait", nw, b104);
  }
  for (int nw : {1, 4, 8}) run<3>("104 x MUFU.EX2 (dependent adds)", nw, 104.0 * 32);   // 'B' = lanes here
  for (int nw : {1, 8}) run<4>("16 x fence.proxy.async", nw, 16.0);
  return 0;
}

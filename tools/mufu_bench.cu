// Throughput of the special-function and FMA pipes on one SM-full of warps (results per SM per clock).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mufu_bench tools/mufu_bench.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i + 1); h[i] = 0xb800b800u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (MODE == 5) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3A000000;" : "+f"(a[i]));
      if (MODE == 6) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 7) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(h[i]));
      if (MODE == 8) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int lanes_per_op) {
  const int threads = 512, iters = 2000;
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * threads * 4); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, threads>>>(out, 10, cyc);
  k<MODE><<<148, threads>>>(out, iters, cyc);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double ops = (double)threads * iters * 8 * lanes_per_op;
  printf("%-28s %7.2f results/clk/SM  (%lld cycles)\n", name, ops / h[0], h[0]);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("tanh.approx.f32", 1);
  run<6>("tanh.approx.bf16x2", 2);
  run<8>("rcp.approx.ftz.f32", 1);
  run<4>("fma.rn.f32 (3 reg)", 1);
  run<5>("fma.rn.f32 (imm)", 1);
  run<7>("fma.rn.bf16x2", 2);
  return 0;
}

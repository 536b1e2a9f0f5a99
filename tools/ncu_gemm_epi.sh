#!/bin/bash
# ncu --set full captures of the gelu / dgelu epilogue GEMMs (image tower shapes)
mkdir -p gpurun_out
for ef in 1 3; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_kernel<\(int\)256, \(bool\)0, \(int\)${ef}," -s 14 -c 1 -o gpurun_out/prof_gemm_ef${ef} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ef${ef}.log 2>&1
  grep -E "WARNING|ERROR" gpurun_out/ncu_ef${ef}.log | head -3
done
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# ncu --set full capture of one launch per GEMM case of tools/gemm_shapes.py
mkdir -p gpurun_out
GEMM_SHAPES_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:gemm_kernel -c 8 -o gpurun_out/prof_gemm_shapes -f \
  python tools/gemm_shapes.py > gpurun_out/ncu_gemm_shapes.log 2>&1
grep -E "WARNING|ERROR|==PROF==" gpurun_out/ncu_gemm_shapes.log | head -12
ls -la gpurun_out/*.ncu-rep

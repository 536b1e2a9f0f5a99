#!/bin/bash
# Round summary captures (1 GPU): launch list with DRAM traffic for one step, plus --set full
# captures of the dominant GEMM (fc2 forward: bias+residual, K=3072) and the attention kernels.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -s 600 -c 545 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:gemm_kernel<\(int\)256, \(bool\)0, \(int\)0," -s 60 -c 1 -o gpurun_out/prof_gemm_final -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm_final.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 16 -c 1 \
  -o gpurun_out/prof_attn_fwd_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_final.csv

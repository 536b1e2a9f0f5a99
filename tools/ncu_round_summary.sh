#!/bin/bash
# Round summary captures (1 GPU).  (1) launch list of the bench command with per-launch duration
# and DRAM traffic; (2) --set full captures of the dominant GEMM (fc2 forward, K=3072, bias +
# residual epilogue), the gelu GEMM, the attention backward and the LayerNorm kernels.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -c 6000 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_final.log 2>&1
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:$2" -s $3 -c 1 -o gpurun_out/prof_$1 -f $B > gpurun_out/ncu_$1.log 2>&1
  grep -E "==ERROR==|No kernels" gpurun_out/ncu_$1.log | head -2
}
cap gemm_resid 'gemm_kernel<\(int\)256, \(bool\)0, \(int\)2,' 20
cap gemm_gelu 'gemm_kernel<\(int\)256, \(bool\)0, \(int\)1,' 20
cap attn_bwd 'attn_bwd_kernel' 30
cap ln_bwd 'ln_bwd_pipe_kernel' 40
cap ln_fwd 'ln_fwd_kernel' 40
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_final.csv

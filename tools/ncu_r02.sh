#!/bin/bash
# Round-2 evidence run (1 GPU): (1) launch list of a short bench with per-launch duration and DRAM
# traffic; (2) one `ncu --set full` capture per kernel of the hot path on the final code.
#   gpurun --timeout 2400 -- 'bash tools/ncu_r02.sh'
mkdir -p gpurun_out/ncu
B="python bench.py --steps 2 --warmup 3 --per-gpu-batch 256 --no-cpu-baseline --no-gpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -c 8000 --csv --log-file gpurun_out/ncu/launches_siglip_b16_n256.csv $B > gpurun_out/ncu/launches.log 2>&1
# the GEMM launches of the full-size default bench (1024 pairs): DRAM traffic per launch for roofline.traffic
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:gemm_kernel -s 933 -c 622 --csv --log-file gpurun_out/ncu/launches_siglip_b16_n1024_gemm.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu/launches_gemm.log 2>&1
cap() {  # name regex skip [command]
  local cmd="${4:-$B}"
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:$2" -s $3 -c 1 -o gpurun_out/ncu/$1 -f $cmd > gpurun_out/ncu/$1.log 2>&1
  grep -E "==ERROR==|No kernels" gpurun_out/ncu/$1.log | head -2
}
# skip counts land in the timed steps (warm-up launches come first)
cap gemm_plain   'gemm_kernel<\(int\)256, \(bool\)0, \(int\)0,' 60
cap gemm_gelu    'gemm_kernel<\(int\)256, \(bool\)0, \(int\)1,' 40
cap gemm_resid   'gemm_kernel<\(int\)256, \(bool\)0, \(int\)2,' 40
cap gemm_dgelu   'gemm_kernel<\(int\)256, \(bool\)0, \(int\)3,' 40
cap gemm_wgrad   'gemm_kernel<\(int\)256, \(bool\)1, \(int\)0,' 40
cap attn_fwd     'attn_fwd_kernel' 30
cap attn_bwd     'attn_bwd_kernel' 30
cap ln_fwd       'ln_fwd_stream_kernel' 40
cap ln_bwd       'ln_bwd_pipe_kernel' 40
cap siglip_loss  'siglip_loss_kernel' 3
cap adam         'adam_kernel' 4
cap patchify     'patchify_kernel' 3
cap embed_fwd    'embed_fwd_kernel' 3
cap embed_bwd    'embed_bwd_table_kernel' 3
cap colsum       'colsum' 10
cap l2norm       'l2norm_fwd_kernel' 3
cap sumsq        'sumsq_kernel' 3
A="python tools/attn_bench.py both"
BV_ATTN_FWD=stream BV_ATTN_BWD=stream cap attn_fwd_stream 'attn_fwd_stream_kernel' 40 "$A"
BV_ATTN_FWD=stream BV_ATTN_BWD=stream cap attn_bwd_stream 'attn_bwd_stream_kernel' 40 "$A"
cap top1 'top1' 0 "python -m pytest tests/test_eval_paths.py -q -m gpu"
cap retrieval 'retrieval' 0 "python -m pytest tests/test_eval_paths.py -q -m gpu"
ls gpurun_out/ncu/*.ncu-rep | wc -l
du -sh gpurun_out/ncu

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_attention_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -q 2>&1 | tail -8
for v in 9 109 108; do
  BV_ATTN_SM=$v BV_ATTN_FWD=stream BV_BENCH_SHAPES="512,16,576;1024,12,196;1024,12,64" timeout -s KILL 120 python tools/attn_bench.py fwd 2>&1 | tail -3
done
BV_ATTN_FWD=resident BV_BENCH_SHAPES="1024,12,196;1024,12,64" timeout -s KILL 120 python tools/attn_bench.py fwd 2>&1 | tail -2

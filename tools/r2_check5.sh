#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_attention_gpu.py -q 2>&1 | tail -4
BV_BENCH_SHAPES="512,16,576;1024,12,196;1024,12,64" BV_ATTN_FWD=stream BV_ATTN_BWD=stream timeout -s KILL 120 python tools/attn_bench.py both 2>&1 | tail -3
BV_ATTN_SM=9 BV_BENCH_SHAPES="512,16,576" BV_ATTN_FWD=stream timeout -s KILL 120 python tools/attn_bench.py fwd 2>&1 | tail -1
timeout -s KILL 900 python bench.py --workload siglip_l14_336 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --profile-calls \
  > gpurun_out/r02_bench_siglip_l14_336_b.json 2> gpurun_out/r02_bench_siglip_l14_336_b.err
cut -c1-220 gpurun_out/r02_bench_siglip_l14_336_b.json; grep "step \|attention\|(all)\|layernorm" gpurun_out/r02_bench_siglip_l14_336_b.err | head -8

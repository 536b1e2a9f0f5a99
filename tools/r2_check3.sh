#!/bin/bash
mkdir -p gpurun_out
tools/bin/mufu_bench | tee gpurun_out/r02_mufu_bench.log
python -m pytest tests/test_optax_gpu.py tests/test_precision_gpu.py tests/test_attention_gpu.py tests/test_kernels_gpu.py -q 2>&1 | tail -8
for cfg in "8 0" "0 4"; do set -- $cfg
  BV_ATTN_SM=$1 BV_BWD_VARIANT=$2 BV_ATTN_FWD=resident BV_ATTN_BWD=resident BV_BENCH_SHAPES="1024,12,196;1024,12,64" timeout -s KILL 120 python tools/attn_bench.py both 2>&1 | tail -2
  BV_ATTN_SM=$(( $1 + 1 )) BV_BWD_VARIANT=$2 BV_BENCH_SHAPES="512,16,576" timeout -s KILL 120 python tools/attn_bench.py both 2>&1 | tail -1
done
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --profile-calls > gpurun_out/r02_bench_siglip_b16_c4.json 2> gpurun_out/r02_bench_siglip_b16_c4.err
cut -c1-200 gpurun_out/r02_bench_siglip_b16_c4.json; grep "step \|attention\|(all)\|layernorm" gpurun_out/r02_bench_siglip_b16_c4.err | head -8

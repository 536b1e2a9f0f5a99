#!/bin/bash
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_attention_gpu.py -q 2>&1 | tail -6
for pk in 0 1; do BV_ATTN_PACK=$pk BV_BENCH_SHAPES="1024,12,64;2048,16,64" timeout -s KILL 120 python tools/attn_bench.py both 2>&1 | tail -2 | sed "s/$/ pack=$pk/"; done
for pk in 0 1; do BV_ATTN_PACK=$pk python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --profile-calls 2>&1 | grep "step \|attention\|\"value\"" | cut -c1-160 | sed "s/$/ pack=$pk/"; done

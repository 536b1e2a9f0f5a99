#!/bin/bash
# One gpurun call that answers "is the tree healthy and how fast is it": GPU tests, smoke, the
# default bench line, the per-call step breakdown and the three kernel timelines.
#   gpurun --timeout 900 -- 'bash tools/round_check.sh'
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile-calls > gpurun_out/check_bench.log 2>&1
grep "step \|attention\|(all)\|layernorm\|adam\|nccl" gpurun_out/check_bench.log
tail -1 gpurun_out/check_bench.log | cut -c1-260
python tools/gemm_shapes.py > gpurun_out/check_gemm_shapes.log 2>&1; cat gpurun_out/check_gemm_shapes.log
(python tools/attn_bwd_timeline.py 196 1024; python tools/attn_bwd_timeline.py 64 1024; python tools/attn_timeline.py 196 1024) \
  2>&1 | cut -c1-180 > gpurun_out/check_attn_timelines.log
grep -A1 "^N=" gpurun_out/check_attn_timelines.log | grep -v "^pair\|^tile\|^--"
python tools/ln_bench.py

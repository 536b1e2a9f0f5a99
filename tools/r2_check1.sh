#!/bin/bash
# Round-2 first GPU call: full GPU suite (no -x), the pipelined attention backward under a hard
# timeout, smoke, bench lines for configs 1-4 with the CPU port and the torch-GPU stand-in beside them.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader | head -2
python -c "import jax" 2>&1 | tail -1
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu.log; tail -4 gpurun_out/r02_pytest_gpu.log
BV_ATTN_BWD_PIPE=1 timeout -s KILL 240 python -m pytest tests/test_kernels_gpu.py -k "attention" -q 2>&1 | tail -15 > gpurun_out/r02_pytest_pipe.log; echo "pipe rc=$?"; tail -3 gpurun_out/r02_pytest_pipe.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 6 --warmup 3 --profile-calls > gpurun_out/r02_bench_siglip_b16.json 2> gpurun_out/r02_bench_siglip_b16.err
cut -c1-400 gpurun_out/r02_bench_siglip_b16.json; grep "step \|attention\|(all)\|layernorm\|adam" gpurun_out/r02_bench_siglip_b16.err | head
if grep -q passed gpurun_out/r02_pytest_pipe.log && ! grep -q failed gpurun_out/r02_pytest_pipe.log; then
  BV_ATTN_BWD_PIPE=1 timeout -s KILL 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline --profile-calls \
    > gpurun_out/r02_bench_pipe.json 2> gpurun_out/r02_bench_pipe.err
  cut -c1-200 gpurun_out/r02_bench_pipe.json; grep "step \|attention" gpurun_out/r02_bench_pipe.err | head -4
fi
for w in vit_b16_cls mixer_b16 vit_s16; do
  timeout -s KILL 600 python bench.py --workload $w --steps 10 --warmup 3 --profile-calls > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  cut -c1-300 gpurun_out/r02_bench_$w.json; tail -2 gpurun_out/r02_bench_$w.err | cut -c1-300
  grep "step \|(all)" gpurun_out/r02_bench_$w.err | head -3
done
python - <<'PY'
import json
for w in ["siglip_b16","vit_b16_cls","mixer_b16","vit_s16"]:
  try:
    d=json.loads(open(f"gpurun_out/r02_bench_{w}.json").read().strip().splitlines()[-1])
    g=d.get("gpu_baseline") or {}; c=d.get("cpu_baseline") or {}
    print(w, "ours", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "torch_gpu", g.get("value", g.get("unavailable")), "cpu", c.get("value"), "frac", round(d["roofline"]["frac"],3), "mfu", round(d["roofline"]["step_mfu"],3))
  except Exception as e: print(w, "ERR", e)
PY

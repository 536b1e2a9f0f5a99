#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 8 --warmup 3 --profile-calls > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
grep "step \|attention\|(all)\|layernorm" gpurun_out/r02_bench_final.err | head -7
for lib in "" "$PWD/big_vision_b200/libbv_b200_hint.so"; do
  BV_LIB_PATH=$lib python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=${lib##*/}', round(d['value'],1), 'pairs/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'MHz', round(d['roofline']['achieved'],1), 'TF/s gemm')"
done
BV_BENCH_SHAPES="1024,12,64;1024,12,196" timeout -s KILL 120 python tools/attn_bench.py both 2>&1 | tail -2
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
g, c = d.get("gpu_baseline") or {}, d.get("cpu_baseline") or {}
print("FINAL value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "torch_gpu", g.get("value"), "cpu", c.get("value"),
      "frac", round(d["roofline"]["frac"], 3), "mfu", round(d["roofline"]["step_mfu"], 3), "launches", d["gpu_launches"], d["clocks"])
PY

#!/bin/bash
# Final check of the round: the driver's own commands (pytest -x -m gpu, smoke, default bench) + config 5.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 8 --warmup 3 --profile-calls > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
grep "step \|attention\|(all)\|layernorm" gpurun_out/r02_bench_final.err | head -7
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
timeout -s KILL 900 python bench.py --workload siglip_l14_336 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --profile-calls \
  > gpurun_out/r02_bench_l14_final.json 2> gpurun_out/r02_bench_l14_final.err
grep "step \|attention\|(all)\|layernorm" gpurun_out/r02_bench_l14_final.err | head -7
python - <<'PY'
import json
for f in ["r02_bench_final", "r02_bench_l14_final"]:
  d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
  g, c = d.get("gpu_baseline") or {}, d.get("cpu_baseline") or {}
  print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "torch_gpu", g.get("value"), "cpu", c.get("value"),
        "frac", round(d["roofline"]["frac"], 3), "mfu", round(d["roofline"]["step_mfu"], 3), "launches", d["gpu_launches"], d["clocks"])
PY

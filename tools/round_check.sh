#!/bin/bash
# One gpurun call that answers "is the tree healthy and how fast is it": the whole GPU suite (no -x: one
# failure must not hide the rest), smoke, the default bench line with its per-call breakdown.
#   gpurun --timeout 1200 -- 'bash tools/round_check.sh'
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/round_check_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 8 --warmup 3 --profile-calls > gpurun_out/round_check_bench.json 2> gpurun_out/round_check_bench.err
grep "step \|attention\|(all)\|layernorm\|adam\|nccl" gpurun_out/round_check_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/round_check_bench.json").read().strip().splitlines()[-1])
g, c = d.get("gpu_baseline") or {}, d.get("cpu_baseline") or {}
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "torch_gpu", g.get("value"), "cpu", c.get("value"),
      "frac", round(d["roofline"]["frac"], 3), "mfu", round(d["roofline"]["step_mfu"], 3), "launches", d["gpu_launches"], d["clocks"])
PY

#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu2.log; tail -5 gpurun_out/r02_pytest_gpu2.log
cat gpurun_out/r02_precision.json 2>/dev/null
for v in 0 1 3 5 7 9 11; do
  BV_ATTN_SM=$v BV_ATTN_FWD=stream BV_BENCH_SHAPES="512,16,576;1024,12,196;1024,12,64" timeout -s KILL 120 python tools/attn_bench.py fwd 2>&1 | tail -3
done
timeout -s KILL 900 python bench.py --workload siglip_l14_336 --steps 3 --warmup 3 --no-cpu-baseline --profile-calls \
  > gpurun_out/r02_bench_siglip_l14_336.json 2> gpurun_out/r02_bench_siglip_l14_336.err
cut -c1-700 gpurun_out/r02_bench_siglip_l14_336.json; grep "step \|attention\|(all)\|layernorm\|adam\|cast\|patchify" gpurun_out/r02_bench_siglip_l14_336.err | head -12; tail -3 gpurun_out/r02_bench_siglip_l14_336.err | cut -c1-400
timeout -s KILL 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --input uint8 > gpurun_out/r02_bench_siglip_b16_u8.json 2>gpurun_out/r02_bench_siglip_b16_u8.err; cut -c1-300 gpurun_out/r02_bench_siglip_b16_u8.json; tail -2 gpurun_out/r02_bench_siglip_b16_u8.err
python - <<'PY'
import json
for w in ["siglip_l14_336", "siglip_b16_u8"]:
  try:
    d=json.loads(open(f"gpurun_out/r02_bench_{w}.json").read().strip().splitlines()[-1])
    g=d.get("gpu_baseline") or {}
    print(w, "ours", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "h2d", d["e2e"]["h2d_bytes_per_step"], "torch_gpu", g.get("value", g.get("unavailable")), "frac", round(d["roofline"]["frac"],3), "mfu", round(d["roofline"]["step_mfu"],3), "mem", d["config"].get("peak_mem_gib"))
  except Exception as e: print(w, "ERR", e)
PY

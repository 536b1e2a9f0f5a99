#!/bin/bash
# Final check of the round with the driver's own commands: pytest -x -m gpu, smoke, default bench, reference arm;
# plus the other BASELINE workloads (short) so that every bench line exists on the final tree.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 8 --warmup 3 --profile-calls > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
grep "step \|attention\|(all)\|layernorm" gpurun_out/r02_bench_final.err | head -7
for w in vit_b16_cls mixer_b16 vit_s16 siglip_l14_336; do
  st=10; [ $w = siglip_l14_336 ] && st=3
  timeout -s KILL 600 python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-gpu-baseline --profile-calls > gpurun_out/r02_bench_final_$w.json 2> gpurun_out/r02_bench_final_$w.err
  grep "profile-calls" gpurun_out/r02_bench_final_$w.err | head -12
done
python - <<'PY'
import json
for f in ["r02_bench_final", "r02_bench_final_vit_b16_cls", "r02_bench_final_mixer_b16", "r02_bench_final_vit_s16", "r02_bench_final_siglip_l14_336"]:
  try:
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
  except Exception as e:
    print(f, "ERR", e); continue
  g, c = d.get("gpu_baseline") or {}, d.get("cpu_baseline") or {}
  print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "torch_gpu", g.get("value"), "cpu", c.get("value"),
        "frac", round(d["roofline"]["frac"], 3), "mfu", round(d["roofline"]["step_mfu"], 3), "launches", d["gpu_launches"], d["clocks"]["sm_mhz"])
PY
# ncu --set full of the kernels written after the round's main capture (profiles/r02/ncu): the vectorised
# Mixer transposes and the max pool; digested on the box, text only comes back
mkdir -p gpurun_out/ncu2 /tmp/ncu2
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'transpose_tokens_kernel|untranspose_add_kernel|pool_max_bwd_kernel|pool_fwd_kernel' \
  -o /tmp/ncu2/late -f python tools/kernel_zoo.py > gpurun_out/ncu2/late.log 2>&1
ncu -i /tmp/ncu2/late.ncu-rep --page raw --csv > /tmp/ncu2/late_raw.csv 2>/dev/null
python tools/ncu_digest.py /tmp/ncu2/late_raw.csv > gpurun_out/ncu2/ncu_summary_late_kernels.md
gzip -c /tmp/ncu2/late_raw.csv > gpurun_out/ncu2/late_raw.csv.gz
tail -12 gpurun_out/ncu2/ncu_summary_late_kernels.md

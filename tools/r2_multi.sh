#!/bin/bash
# Multi-GPU check (gpurun --gpus N): 2-rank NCCL parity tests (both loss forms), then the bench under
# torchrun with the bucketed/overlapped gradient all-reduce and, for comparison, the single one.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -s KILL 600 python -m pytest tests/test_dist_gpu.py -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_dist_${N}gpu.log
run() {  # tag env...
  tag=$1; shift
  env "$@" timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29811 bench.py --gpus $N --steps ${STEPS:-8} --warmup 3 --profile-calls \
    > gpurun_out/r02_bench_${N}gpu_$tag.json 2> gpurun_out/r02_bench_${N}gpu_$tag.err
  tail -1 gpurun_out/r02_bench_${N}gpu_$tag.json | cut -c1-260
  grep "step \|nccl" gpurun_out/r02_bench_${N}gpu_$tag.err | head -6
}
run single BV_X=1
run overlap BV_GRAD_ALLREDUCE=overlap
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29812 bench.py --impl torch_gpu --gpus $N --steps 4 --warmup 3 > gpurun_out/r02_bench_${N}gpu_torch.json 2> gpurun_out/r02_bench_${N}gpu_torch.err
tail -1 gpurun_out/r02_bench_${N}gpu_torch.json | cut -c1-400
if [ "${2:-}" = "l14" ]; then
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29813 bench.py --workload siglip_l14_336 --gpus $N --steps 2 --warmup 3 \
    > gpurun_out/r02_bench_${N}gpu_l14.json 2> gpurun_out/r02_bench_${N}gpu_l14.err
  tail -1 gpurun_out/r02_bench_${N}gpu_l14.json | cut -c1-300; tail -2 gpurun_out/r02_bench_${N}gpu_l14.err | cut -c1-300
fi
python - <<PY
import json
for t in ["overlap", "single", "torch"]:
  try:
    d = json.loads(open(f"gpurun_out/r02_bench_${N}gpu_{t}.json").read().strip().splitlines()[-1])
    print(t, "n_gpus", d.get("n_gpus"), "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2))
  except Exception as e:
    print(t, "ERR", e)
PY

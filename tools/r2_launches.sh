#!/bin/bash
# ncu launch list of the DEFAULT bench command's workload (config 4 at 1024 pairs per GPU): per-launch
# duration and DRAM bytes of every kernel of 1 warm-up + 2 timed steps.  Digested on the box.
#   gpurun --timeout 900 -- 'bash tools/r2_launches.sh'
mkdir -p gpurun_out/launches /tmp/lx
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -c 2000 --csv --log-file /tmp/lx/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/launches/run.log 2>&1
tail -2 gpurun_out/launches/run.log | cut -c1-300
python tools/launch_summary.py /tmp/lx/launches.csv \
  "ncu launch list of \`bench.py --steps 2 --warmup 1\` (config 4, 1024 pairs): the first 2000 launches = parameter initialisation, the warm-up step and the timed steps, per-kernel totals" \
  > gpurun_out/launches/launch_summary_siglip_b16_n1024.md
gzip -c /tmp/lx/launches.csv > gpurun_out/launches/launches_siglip_b16_n1024.csv.gz
head -24 gpurun_out/launches/launch_summary_siglip_b16_n1024.md
ls -la gpurun_out/launches

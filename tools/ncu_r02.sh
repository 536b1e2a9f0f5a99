#!/bin/bash
# Round-2 evidence run (1 GPU).  (1) ONE `ncu --set full` pass over tools/kernel_zoo.py (every kernel of
# the hot path at bench-like shapes, second launch of each = warm), digested ON the box so that only
# text comes back (the reports themselves exceed the 64 MiB return limit); (2) launch lists of the real
# bench with per-launch duration and DRAM traffic.
#   gpurun --timeout 1800 -- 'bash tools/ncu_r02.sh'
mkdir -p gpurun_out/ncu /tmp/ncu
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -o /tmp/ncu/zoo -f python tools/kernel_zoo.py > gpurun_out/ncu/zoo.log 2>&1
tail -2 gpurun_out/ncu/zoo.log
ncu -i /tmp/ncu/zoo.ncu-rep --page raw --csv > /tmp/ncu/zoo_raw.csv 2>/dev/null
python tools/ncu_digest.py /tmp/ncu/zoo_raw.csv > gpurun_out/ncu/ncu_summary.md
ncu -i /tmp/ncu/zoo.ncu-rep --page details --csv 2>/dev/null | gzip > gpurun_out/ncu/zoo_details.csv.gz
gzip -c /tmp/ncu/zoo_raw.csv > gpurun_out/ncu/zoo_raw.csv.gz
# source-level stall attribution for the four attention kernels only (large otherwise)
for k in attn_fwd_stream_kernel attn_bwd_stream_kernel attn_fwd_kernel attn_bwd_kernel; do
  ncu -i /tmp/ncu/zoo.ncu-rep --page source --csv -k regex:$k 2>/dev/null | gzip > gpurun_out/ncu/source_$k.csv.gz
done
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -c 8000 --csv --log-file /tmp/ncu/launches_n256.csv $B --per-gpu-batch 256 > gpurun_out/ncu/launches.log 2>&1
gzip -c /tmp/ncu/launches_n256.csv > gpurun_out/ncu/launches_siglip_b16_n256.csv.gz
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:gemm_kernel -s 933 -c 622 --csv --log-file /tmp/ncu/launches_gemm.csv $B > gpurun_out/ncu/launches_gemm.log 2>&1
gzip -c /tmp/ncu/launches_gemm.csv > gpurun_out/ncu/launches_siglip_b16_n1024_gemm.csv.gz
du -sh gpurun_out/ncu; ls gpurun_out/ncu

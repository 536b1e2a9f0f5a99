python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --profile-calls > gpurun_out/bench_2gpu_prof.log 2>&1
grep "step \|nccl\|(all)\|attention_bwd" gpurun_out/bench_2gpu_prof.log; tail -1 gpurun_out/bench_2gpu_prof.log | cut -c1-220
nvidia-smi topo -m 2>&1 | head -8

#!/bin/bash
# round-2 GPU call H (gpurun --gpus 8): weak-scaling curve of bench.py at N = 1, 2, 4, 8 and BASELINE
# config 5 (32768^2 q4 n5) sharded over 8 GPUs, with the 1-GPU time of the same image beside it
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi -L > $OUT/h_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/h_bench_n1.json 2> $OUT/h_bench_n1.err
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n bench.py --gpus $n --steps 10 --warmup 3 > $OUT/h_bench_n${n}.json 2> $OUT/h_bench_n${n}.err
done
timeout 300 $TR --nproc-per-node 8 tools/run_configs.py --configs 5 --size 8192 --check > $OUT/h_cfg5_8k_8.txt 2>&1
timeout 600 $TR --nproc-per-node 8 tools/run_configs.py --configs 5 --size 32768 --reps 2 > $OUT/h_cfg5_32k_8.txt 2>&1
timeout 600 python tools/run_configs.py --configs 5 --size 32768 --reps 1 > $OUT/h_cfg5_32k_1.txt 2>&1
ls -la $OUT | tail -10

#!/bin/bash
# round-2 GPU call E (gpurun --gpus N): the C slab engine on real peers - P2P inside a process,
# CUDA IPC between processes - tests, weak-scaling bench, BASELINE config 5/6 checks
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
N=${1:-2}
nvidia-smi -L > $OUT/e_gpus_$N.txt
nvidia-smi topo -m >> $OUT/e_gpus_$N.txt 2>&1
if [ "$N" = "2" ]; then
timeout 1200 python -m pytest tests/test_gpu_slab_engine.py -m gpu -x -q -k "multi or ipc or sharded or processes" > $OUT/e_pytest_$N.log 2>&1; echo "rc=$?" >> $OUT/e_pytest_$N.log
fi
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/e_bench_n1_of_$N.json 2> $OUT/e_bench_n1_of_$N.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 $TR --nproc-per-node $n bench.py --gpus $n --steps 10 --warmup 3 > $OUT/e_bench_n${n}.json 2> $OUT/e_bench_n${n}.err
  fi
done
timeout 600 $TR --nproc-per-node $N tools/run_configs.py --configs 5 --size 8192 --check > $OUT/e_cfg5_8k_$N.txt 2>&1
timeout 600 $TR --nproc-per-node $N tools/run_configs.py --configs 6 --size 4096 --check > $OUT/e_cfg6_4k_$N.txt 2>&1
timeout 300 python tools/dropin_probe.py --flags 0 --gpus $N --reps 6 > $OUT/e_dropin_multi_$N.txt 2>&1
if [ "$N" = "8" ]; then
timeout 900 $TR --nproc-per-node 8 tools/run_configs.py --configs 5 --size 32768 --reps 2 > $OUT/e_cfg5_32k_8.txt 2>&1
timeout 900 $TR --nproc-per-node 1 tools/run_configs.py --configs 5 --size 32768 --reps 1 > $OUT/e_cfg5_32k_1.txt 2>&1
fi
ls -la $OUT | tail -12

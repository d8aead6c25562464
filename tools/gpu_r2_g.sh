#!/bin/bash
# round-2 GPU call G (--gpus 2): the multi-rank tests again (hardware work queues raised), the
# single-process multi-device drop-in call with banded host I/O
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_slab_engine.py -m gpu -x -q -k "multi or ipc or sharded or processes" > $OUT/g_pytest_2.log 2>&1; echo "rc=$?" >> $OUT/g_pytest_2.log
timeout 300 python tools/dropin_probe.py --flags 0 --gpus 2 --reps 6 > $OUT/g_dropin_multi_2.txt 2>&1
timeout 300 python tools/dropin_probe.py --flags 0 --reps 6 >> $OUT/g_dropin_multi_2.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/g_bench_n2.json 2> $OUT/g_bench_n2.err
ls -la $OUT | tail -6

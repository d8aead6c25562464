#!/bin/bash
# round-2 GPU call B: kernel tweaks (branch-free update, rebalance ILP, header prefetch, folded
# IDCT rounding, coalesced IDCT pass) + threaded row pipeline of the host entry points
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
L=jpeg-quantsmooth_b200/csrc
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/b_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/b_pytest.log
timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 > $OUT/b_tune.txt 2>&1
timeout 300 python tools/tune.py --flags 1 --variants 2:4 --steps 3 >> $OUT/b_tune.txt 2>&1
JPEGQS_B200_LIB=$L/libjpegqs_b200_idct1.so timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 >> $OUT/b_tune.txt 2>&1
JPEGQS_B200_LIB=$L/libjpegqs_b200_idct3.so timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 >> $OUT/b_tune.txt 2>&1
JPEGQS_B200_LIB=$L/libjpegqs_b200_phase.so timeout 300 python tools/phase_probe.py --flags 0 > $OUT/b_phase_q3.txt 2>&1
for t in 8 4 16; do
JPEGQS_IO_THREADS=$t timeout 300 python tools/dropin_probe.py --flags 0 >> $OUT/b_dropin.txt 2>&1
done
timeout 300 python tools/dropin_probe.py --flags 7 >> $OUT/b_dropin.txt 2>&1
timeout 300 python tools/e2e_probe.py > $OUT/b_e2e_probe.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/b_bench.json 2> $OUT/b_bench.err
ls -la $OUT | tail -12

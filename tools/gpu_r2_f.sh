#!/bin/bash
# round-2 GPU call F: kernel tweaks (3-input IDCT sums, rebalance sums in the update, tile
# look-ahead + L2 prefetch), swizzled coalesced IDCT variants, CLI wall time, full test suite
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
L=jpeg-quantsmooth_b200/csrc
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/f_pytest.log 2>&1; echo "rc=$?" >> $OUT/f_pytest.log
JPEGQS_B200_LIB=$L/libjpegqs_b200_phase.so timeout 300 python tools/phase_probe.py --flags 0 > $OUT/f_phase_q3.txt 2>&1
for v in "" _idctx2 _idctx3; do
  echo "== libjpegqs_b200$v" >> $OUT/f_tune.txt
  JPEGQS_B200_LIB=$L/libjpegqs_b200$v.so timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 >> $OUT/f_tune.txt 2>&1
done
timeout 300 python tools/tune.py --flags 1 --variants 2:4 --steps 3 >> $OUT/f_tune.txt 2>&1
timeout 600 bash tools/cli_walltime.sh > $OUT/f_cli_walltime.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/f_bench.json 2> $OUT/f_bench.err
ls -la $OUT | tail -8

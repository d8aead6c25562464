#!/bin/bash
# round-2 GPU call D: issue-rate microbenchmark, IDCT / LOW_QUALITY occupancy variants, lowq workload,
# drop-in probe q6, ncu of the shipped kernels
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
L=jpeg-quantsmooth_b200/csrc
timeout 120 tools/ubench_core > $OUT/d_ubench_core.txt 2>&1
for v in "" _idctp3 _idctp4; do
  echo "== libjpegqs_b200$v" >> $OUT/d_tune.txt
  JPEGQS_B200_LIB=$L/libjpegqs_b200$v.so timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 >> $OUT/d_tune.txt 2>&1
done
for v in "" _lowq1 _lowq6; do
  echo "== libjpegqs_b200$v" >> $OUT/d_lowq.txt
  JPEGQS_B200_LIB=$L/libjpegqs_b200$v.so timeout 300 python tools/run_configs.py --configs 7 >> $OUT/d_lowq.txt 2>&1
done
timeout 600 python tools/run_configs.py --configs 2,3,4 > $OUT/d_configs.txt 2>&1
timeout 300 python tools/dropin_probe.py --flags 0 > $OUT/d_dropin.txt 2>&1
timeout 300 python tools/dropin_probe.py --flags 1 >> $OUT/d_dropin.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:qs_smooth -s 2 -c 1 -o $OUT/r02_smooth_q3_v2 -f python tools/tune.py --flags 0 --variants 2:4 --steps 1 > $OUT/d_ncu1.log 2>&1
timeout 600 $NCU -k regex:qs_idct_pass -s 3 -c 2 -o $OUT/r02_idct_v2 -f python tools/tune.py --flags 0 --variants 2:4 --steps 1 > $OUT/d_ncu2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/d_bench.json 2> $OUT/d_bench.err
ls -la $OUT | tail -8

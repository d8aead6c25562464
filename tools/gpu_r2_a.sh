#!/bin/bash
# round-2 GPU call A: sanity, phase clocks, drop-in call "before", ncu captures of the kernels
# that had no evidence in round 1
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/a_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/a_pytest.log
PH=jpeg-quantsmooth_b200/csrc/libjpegqs_b200_phase.so
JPEGQS_B200_LIB=$PH timeout 300 python tools/phase_probe.py --flags 0 > $OUT/a_phase_q3.txt 2>&1
JPEGQS_B200_LIB=$PH timeout 300 python tools/phase_probe.py --flags 1 > $OUT/a_phase_q4.txt 2>&1
timeout 300 python tools/dropin_probe.py --flags 0 > $OUT/a_dropin_before.txt 2>&1
timeout 300 python tools/dropin_probe.py --flags 0 --contiguous >> $OUT/a_dropin_before.txt 2>&1
timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 > $OUT/a_tune_q3.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:qs_smooth -s 2 -c 1 -o $OUT/r02_smooth_q3_base -f python tools/tune.py --flags 0 --variants 2:4 --steps 1 > $OUT/a_ncu1.log 2>&1
timeout 600 $NCU -k regex:qs_smooth -s 2 -c 1 -o $OUT/r02_smooth_q4_base -f python tools/tune.py --flags 1 --variants 2:4 --steps 1 > $OUT/a_ncu2.log 2>&1
timeout 600 $NCU -k regex:qs_lowq -s 2 -c 1 -o $OUT/r02_lowq_base -f python tools/tune.py --flags 8 --variants 2:4 --steps 1 > $OUT/a_ncu3.log 2>&1
timeout 600 $NCU -k regex:qs_idct_pass -s 3 -c 2 -o $OUT/r02_idct_base -f python tools/tune.py --flags 0 --variants 2:4 --steps 1 > $OUT/a_ncu4.log 2>&1
ls -la $OUT | tail -20

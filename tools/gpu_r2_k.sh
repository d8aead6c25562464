#!/bin/bash
# round-2 GPU call K: mixed chunks (tuning key 8) A/B
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
L=jpeg-quantsmooth_b200/csrc
timeout 600 python -m pytest tests/test_gpu_edge.py -m gpu -x -q -k "tuning or libjpeg_facing" > $OUT/k_pytest.log 2>&1; echo "rc=$?" >> $OUT/k_pytest.log
timeout 300 python tools/tune.py --flags 0 --variants 2:4:4:1:0:1:0,2:4:4:1:0:1:1,2:4:4:1:0:1:0,2:4:4:1:0:1:1 --steps 3 > $OUT/k_tune.txt 2>&1
timeout 300 python tools/tune.py --flags 1 --variants 2:4:4:1:0:1:0,2:4:4:1:0:1:1 --steps 3 >> $OUT/k_tune.txt 2>&1
JPEGQS_B200_LIB=$L/libjpegqs_b200_phase.so timeout 300 python tools/phase_probe.py --flags 0 --merge 1 > $OUT/k_phase_merge1.txt 2>&1
JPEGQS_B200_LIB=$L/libjpegqs_b200_phase.so timeout 300 python tools/phase_probe.py --flags 0 --merge 0 > $OUT/k_phase_merge0.txt 2>&1
ls -la $OUT | tail -5

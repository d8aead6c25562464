#!/bin/bash
# round-2 GPU call C: slab engine + real-libjpeg + whole-image parity tests, phase clocks of the
# tweaked smoothing kernel, IDCT pass A/B (coalesced vs thread-private loads, occupancy)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
L=jpeg-quantsmooth_b200/csrc
timeout 1500 python -m pytest tests/test_gpu_slab_engine.py tests/test_real_libjpeg.py -m gpu -x -q > $OUT/c_pytest_new.log 2>&1; echo "rc=$?" >> $OUT/c_pytest_new.log
JPEGQS_B200_LIB=$L/libjpegqs_b200_phase.so timeout 300 python tools/phase_probe.py --flags 0 > $OUT/c_phase_q3.txt 2>&1
for v in "" _idct1 _idct3 _idctp1 _idctp2; do
  echo "== libjpegqs_b200$v" >> $OUT/c_tune.txt
  JPEGQS_B200_LIB=$L/libjpegqs_b200$v.so timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 >> $OUT/c_tune.txt 2>&1
done
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:qs_idct_pass -s 3 -c 2 -o $OUT/r02_idct_coalesced -f python tools/tune.py --flags 0 --variants 2:4 --steps 1 > $OUT/c_ncu1.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_slab_engine.py --deselect tests/test_real_libjpeg.py > $OUT/c_pytest_old.log 2>&1; echo "rc=$?" >> $OUT/c_pytest_old.log
ls -la $OUT | tail -8

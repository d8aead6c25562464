#!/bin/bash
# round-2 GPU call M: section-loop unroll factors, branch-free border word, rolled low-quality kernel (A/B, one box)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
L=jpeg-quantsmooth_b200/csrc
: > $OUT/m_tune.txt
for v in "" _blut _ub2 _ub4 _uh2 _uv2 _u212 _u222 ""; do
	echo "== libjpegqs_b200$v" >> $OUT/m_tune.txt
	JPEGQS_B200_LIB=$L/libjpegqs_b200$v.so timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 >> $OUT/m_tune.txt 2>&1
done
for v in "" _u222; do
	echo "== q4 libjpegqs_b200$v" >> $OUT/m_tune.txt
	JPEGQS_B200_LIB=$L/libjpegqs_b200$v.so timeout 300 python tools/tune.py --flags 1 --variants 2:4 --steps 2 >> $OUT/m_tune.txt 2>&1
done
for v in "" _lqr _lqr2 "" _lqr2; do
	echo "== lowq libjpegqs_b200$v" >> $OUT/m_tune.txt
	JPEGQS_B200_LIB=$L/libjpegqs_b200$v.so timeout 300 python tools/tune.py --flags 9 --variants 2:4 --steps 3 >> $OUT/m_tune.txt 2>&1
done
cat $OUT/m_tune.txt

#!/bin/bash
# round-2 GPU call Q (the last minute of the budget): jpegqs wall time with the threaded reader / writer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/q
E=jpeg-quantsmooth_b200/csrc/jpegqs; T=$(mktemp -d); I=tools/_in8k.jpg
{
JPEGQS_CODEC_TRACE=1 $E -v 1 -i 0 -q 3 $I $T/out.jpg 2>&1 | grep "wall time\|jpegcoef"
JPEGQS_SERIAL_DECODE=1 $E -v 1 -i 0 -q 3 $I $T/out_s.jpg 2>&1 | grep "wall time" | sed 's/^/serial decode: /'
cmp $T/out.jpg $T/out_s.jpg && echo "same file with the one-thread reader"
args=""; for k in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 17 18 19 20 21 22 23 24; do args="$args $I $T/o$k.jpg"; done
$E -v 1 -i 0 -q 3 --batch $args 2>&1 | grep "wall time\|^batch" | tail -3 | sed 's/^/batch of 24 (pipeline): /'
JPEGQS_NO_PIPELINE=1 $E -v 1 -i 0 -q 3 --batch $args 2>&1 | grep "wall time\|^batch" | tail -2 | sed 's/^/batch of 24 (one after the other): /'
cmp $T/out.jpg $T/o24.jpg && echo "batch output identical"
} > gpurun_out/q/q_cli.txt 2>&1
cat gpurun_out/q/q_cli.txt
rm -rf $T

#!/bin/bash
# wall-time breakdown of the jpegqs tool on an 8K file: each run is a fresh
# process, so CUDA start-up is in; with and without the start-up overlap (JPEGQS_NO_WARMUP=1)
cd "$(dirname "$0")/.."
T=$(mktemp -d)
python - "$T" <<'PY'
import sys, numpy as np
from PIL import Image
rng = np.random.RandomState(5)
h, w = 4320, 7680
y, x = np.mgrid[0:h, 0:w]
a = (128 + 60 * np.sin(x / 97.0) + 50 * np.cos(y / 131.0)).astype(np.float32)
img = np.stack([a + rng.randint(-12, 12, (h, w)), a[::-1] + rng.randint(-12, 12, (h, w)), a.T[:h, :w] if False else a[:, ::-1]], -1).clip(0, 255).astype(np.uint8)
Image.fromarray(img, "RGB").save(sys.argv[1] + "/in.jpg", quality=75)
PY
ls -la $T/in.jpg
E=jpeg-quantsmooth_b200/csrc/jpegqs
for mode in warm nowarm; do
  for i in 1 2 3; do
    if [ $mode = nowarm ]; then export JPEGQS_NO_WARMUP=1; else unset JPEGQS_NO_WARMUP; fi
    t0=$(date +%s.%N)
    $E -v 1 -i 8 -q 3 $T/in.jpg $T/out.jpg 2>&1 | grep "wall time\|quantsmooth" | tr '\n' ' '
    t1=$(date +%s.%N)
    echo " | $mode run $i: process total $(python3 -c "print(round($t1 - $t0, 3))") s"
  done
done
# one process, many files: start-up amortised; read / smooth / write of different files overlap
# (three-stage pipeline), JPEGQS_NO_PIPELINE=1 runs the pairs one after the other
args=""
for k in 1 2 3 4 5 6 7 8 9 10 11 12; do args="$args $T/in.jpg $T/o$k.jpg"; done
for mode in pipeline readers1 sequential; do
  unset JPEGQS_NO_PIPELINE JPEGQS_BATCH_READERS
  if [ $mode = sequential ]; then export JPEGQS_NO_PIPELINE=1; fi
  if [ $mode = readers1 ]; then export JPEGQS_BATCH_READERS=1; fi
  t0=$(date +%s.%N)
  $E -v 1 -i 0 -q 3 --batch $args 2>&1 | grep "wall time\|^batch" | tail -3 | sed "s/^/  batch ($mode), last 2 of 12 + steady state: /"
  t1=$(date +%s.%N)
  echo "batch of 12 files in one process ($mode): total $(python3 -c "print(round($t1 - $t0, 3))") s"
  cmp $T/out.jpg $T/o12.jpg && echo "batch output identical to the single-file run"
done
rm -rf $T

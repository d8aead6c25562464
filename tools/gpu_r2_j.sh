#!/bin/bash
# round-2 GPU call J: cell-based up-sampling kernel (q6), CLI batch mode, last full suite
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/j_pytest.log 2>&1; echo "rc=$?" >> $OUT/j_pytest.log
timeout 600 python tools/run_configs.py --configs 3 > $OUT/j_configs.txt 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:qs_upsample" -c 2 -o $OUT/r02_upsample_v2 -f python tools/tune.py --flags 7 --variants 2:4 --steps 1 > $OUT/j_ncu.log 2>&1
timeout 600 bash tools/cli_walltime.sh > $OUT/j_cli_walltime.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $OUT/j_smoke.log 2>&1
ls -la $OUT | tail -6

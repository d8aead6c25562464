#!/bin/bash
# round-2 GPU call N: GPU suite (parallel workers, one GPU), LOW_QUALITY kernel after the rolled loops /
# rescaling, CLI wall time with the faster codec and the batch pipeline
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/n; mkdir -p $OUT
nproc > $OUT/n_nproc.txt
timeout 600 python -m pytest tests -m gpu -q -n 6 --dist loadfile > $OUT/n_pytest.log 2>&1; echo "rc=$?" >> $OUT/n_pytest.log
tail -5 $OUT/n_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/n_smoke.log 2>&1
timeout 300 python tools/tune.py --flags 9 --variants 2:4 --steps 3 > $OUT/n_tune.txt 2>&1
timeout 300 python tools/tune.py --flags 11 --variants 2:4 --steps 3 >> $OUT/n_tune.txt 2>&1
timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 >> $OUT/n_tune.txt 2>&1
cat $OUT/n_tune.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:qs_lowq -s 2 -c 1 -o $OUT/r02_lowq_final -f python tools/tune.py --flags 9 --variants 2:4 --steps 1 > $OUT/n_ncu3.log 2>&1
timeout 600 bash tools/cli_walltime.sh > $OUT/n_cli_walltime.txt 2>&1
cat $OUT/n_cli_walltime.txt

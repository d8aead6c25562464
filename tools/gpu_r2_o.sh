#!/bin/bash
# round-2 GPU call O: CLI tests on the device (new file flavours, batch pipeline), CLI wall time incl. the
# steady-state batch rate, final bench lines and launch list, LOW_QUALITY summary
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/o; mkdir -p $OUT
timeout 400 python -m pytest tests/test_cli.py -m gpu -x -q > $OUT/o_pytest_cli.log 2>&1; echo "rc=$?" >> $OUT/o_pytest_cli.log
tail -3 $OUT/o_pytest_cli.log
timeout 300 python -m pytest tests/test_golden.py tests/test_gpu_edge.py -m gpu -x -q -k "golden or low_quality" > $OUT/o_pytest_lowq.log 2>&1; echo "rc=$?" >> $OUT/o_pytest_lowq.log
tail -3 $OUT/o_pytest_lowq.log
timeout 300 bash tools/cli_walltime.sh > $OUT/o_cli_walltime.txt 2>&1
cat $OUT/o_cli_walltime.txt
timeout 300 python bench.py > $OUT/o_bench.json 2> $OUT/o_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/o_bench_ref.json 2> $OUT/o_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/o_ncu_bench.log 2>&1
timeout 200 python tools/run_configs.py --configs 7 > $OUT/o_lowq.txt 2>&1
cat $OUT/o_bench.json $OUT/o_bench_ref.json $OUT/o_lowq.txt

#!/bin/bash
# round-2 GPU call I: final state - full GPU suite, smoke, bench, launch list of the bench
# command, full ncu captures of the shipped kernels, CLI wall time, drop-in probe
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/final; mkdir -p gpurun_out/final
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/i_pytest.log 2>&1; echo "rc=$?" >> $OUT/i_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/i_smoke.log 2>&1
timeout 600 python bench.py > $OUT/i_bench.json 2> $OUT/i_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/i_bench_ref.json 2> $OUT/i_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/i_ncu_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:qs_smooth -s 2 -c 1 -o $OUT/r02_smooth_q3_final -f python tools/tune.py --flags 0 --variants 2:4 --steps 1 > $OUT/i_ncu1.log 2>&1
timeout 600 $NCU -k regex:qs_smooth -s 2 -c 1 -o $OUT/r02_smooth_q4_final -f python tools/tune.py --flags 1 --variants 2:4 --steps 1 > $OUT/i_ncu2.log 2>&1
timeout 600 $NCU -k regex:qs_lowq -s 2 -c 1 -o $OUT/r02_lowq_final -f python tools/tune.py --flags 8 --variants 2:4 --steps 1 > $OUT/i_ncu3.log 2>&1
timeout 600 $NCU -k "regex:qs_upsample|qs_downsample|qs_fdct_plane" -c 4 -o $OUT/r02_updown_final -f python tools/tune.py --flags 7 --variants 2:4 --steps 1 > $OUT/i_ncu4.log 2>&1
timeout 600 bash tools/cli_walltime.sh > $OUT/i_cli_walltime.txt 2>&1
timeout 300 python tools/dropin_probe.py --flags 0 > $OUT/i_dropin.txt 2>&1
timeout 300 python tools/tune.py --flags 0 --variants 2:4 --steps 3 > $OUT/i_tune.txt 2>&1
timeout 300 python tools/tune.py --flags 1 --variants 2:4 --steps 3 >> $OUT/i_tune.txt 2>&1
ls -la $OUT | tail -14

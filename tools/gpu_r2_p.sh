#!/bin/bash
# round-2 GPU call P: the rest of the GPU suite on the final code, as far as the remaining budget goes
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/p
timeout 420 python -m pytest tests/test_gpu_edge.py tests/test_gpu_slab_engine.py tests/test_gpu_parity.py tests/test_real_libjpeg.py -m gpu -x -q --durations=15 > gpurun_out/p/p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/p/p_pytest.log
tail -25 gpurun_out/p/p_pytest.log

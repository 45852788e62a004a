#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_benchsize.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_c31_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c31_smoke.log 2>&1
echo done

#!/bin/bash
# memcheck of the smoke path and a few small parity cases (one sanitizer kind per call)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c26_smoke_plain.log 2>&1 &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_c26_memcheck_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_c26_memcheck_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sliding_windows and not subprocess or cuda_graph_replay or large_temporal_kernel or stgcn_model_c1 or rt_full_tensor_core or costgcn" > gpurun_out/r2_c26_memcheck_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_c26_memcheck_tests.log
echo done

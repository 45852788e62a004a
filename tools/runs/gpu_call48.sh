#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_c48_tests.log
echo done

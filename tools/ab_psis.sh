#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest psis + fast"; python -m pytest tests/test_gpu_psis.py tests/test_gpu_fast.py -x -q -m gpu 2>&1 | tail -15
echo "=== timing"; python tools/run_psis.py 100000000 2 | grep "psislw\|moments-only n"
} > gpurun_out/ab_psis.log 2>&1
tail -30 gpurun_out/ab_psis.log

#!/bin/bash
# A/B of the fast sweep: fp8 correction passes (default) against three fp16 passes (VB_FAST_FP8=0)
mkdir -p gpurun_out
{
echo "=== parity fp8"; timeout 300 python tools/pair_check.py 2>&1 | tail -26
echo "=== parity fp16x3"; VB_FAST_FP8=0 timeout 300 python tools/pair_check.py 2>&1 | tail -9
echo "=== time fp8"; timeout 300 python tools/pair_check.py time 2>&1 | tail -2
echo "=== time fp16x3"; VB_FAST_FP8=0 timeout 300 python tools/pair_check.py time 2>&1 | tail -2
} > gpurun_out/ab_fast.log 2>&1
cat gpurun_out/ab_fast.log

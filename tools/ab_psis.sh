#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest psis"; python -m pytest tests/test_gpu_psis.py -x -q -m gpu 2>&1 | tail -3
echo "=== timing"; python tools/run_psis.py 100000000 2 | grep "psislw\|moments-only n\|rel diff"
} > gpurun_out/ab_psis.log 2>&1
ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_issued.avg.pct_of_peak_sustained_active --clock-control none -k regex:"psis_pass" -s 6 -c 4 --csv --log-file gpurun_out/psis_passes_b.csv python tools/run_psis.py 100000000 2 > /dev/null 2>&1
tail -12 gpurun_out/ab_psis.log; grep -v "^==" gpurun_out/psis_passes_b.csv | cut -d, -f5,13- | tail -14

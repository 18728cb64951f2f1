#!/bin/bash
# A/B of the PSIS pass-A versions and of the one-pass moments mode (one gpurun call)
mkdir -p gpurun_out
{
echo "=== pytest psis"; python -m pytest tests/test_gpu_psis.py -x -q -m gpu 2>&1 | tail -5
echo "=== lean pass A (default)"; python tools/run_psis.py 100000000 2
echo "=== pass A v1"; VB_PSIS_PASS_A=v1 VB_PSIS_ONEPASS=0 python tools/run_psis.py 100000000 2
} > gpurun_out/ab_psis.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:psis -s 30 -c 30 --csv --log-file gpurun_out/psis_launches_r02_b.csv python tools/run_psis.py 100000000 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"psis_cand_gather|psis_tail_count|psis_tail_values|psis_gpd_grid|psis_tail_place|psis_tail_rank|psis_sample_select" -s 14 -c 7 -o gpurun_out/psis_small_r02 -f python tools/run_psis.py 100000000 2 > /dev/null 2>&1
tail -45 gpurun_out/ab_psis.log

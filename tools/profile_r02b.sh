#!/bin/bash
# ncu captures of the round-2 kernels that changed after tools/profile_r02.sh was run (lean PSIS pass A, one-pass
# moments, GPD grid, 16-warp exact sweep).  One gpurun call on ONE B200; reports land in gpurun_out/.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# PSIS: launch list (full mode, then moments-only) and --set full of the two streaming passes
timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:psis -s 26 -c 40 --csv --log-file $O/psis_launches_r02_final.csv python tools/run_psis.py 100000000 2 > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"psis_pass_a_lean|psis_pass_b_kernel" -s 4 -c 2 -o $O/psis_r02_final -f python tools/run_psis.py 100000000 2 > /dev/null 2>&1
# exact sweep
timeout 600 $NCU --set full --import-source on -k regex:glm_sweep_f64 -s 2 -c 1 -o $O/f64_r02_final -f python bench.py --steps 2 --warmup 3 --path f64 --no-psis --no-cpu-baseline > /dev/null 2>&1
# timings outside the profiler
python tools/run_psis.py 100000000 2 > $O/psis_final.log 2>&1
python bench.py --path f64 --steps 5 --warmup 3 --no-psis --no-cpu-baseline > $O/f64_final.json 2> $O/f64_final.err
tail -12 $O/psis_final.log
ls -la $O/*final*

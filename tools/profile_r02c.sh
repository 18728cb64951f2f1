#!/bin/bash
# ncu capture of the fast sweep with the fp8 correction passes (round 2, final form of the kernel)
O=gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -k regex:"glm_fast_pair" -s 15 -c 1 -o $O/step_r02_fp8 -f python bench.py --steps 2 --warmup 3 --no-psis --no-f64 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:"vb::|glm_|mf_" -s 20 -c 30 --csv --log-file $O/bench_launches_r02_fp8.csv python bench.py --steps 2 --warmup 3 --no-psis --no-f64 --no-cpu-baseline > /dev/null 2>&1
ls -la $O/step_r02_fp8.ncu-rep $O/bench_launches_r02_fp8.csv

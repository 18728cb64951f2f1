#!/bin/bash
# ncu captures of round 2 (run under gpurun on ONE B200): launch lists + one --set full capture per dominant kernel.
# Reports land in gpurun_out/; profiles/*.csv are extracted from them with `ncu -i ... --page raw --csv`.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
# 1. launch list of the bench command itself (first 700 launches after the data generation)
timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:"vb::|glm_|mf_|psis_|philox|gemm" -c 700 --csv --log-file $O/bench_launches_r02.csv $B > $O/bench_under_ncu_r02.log 2>&1
# 2. the fused step: pair kernel + pre + post kernels
timeout 600 $NCU --set full --import-source on -k regex:"glm_fast_pair|mf_pre_kernel|mf_post_kernel" -s 15 -c 3 -o $O/step_r02 -f $B --no-psis --no-f64 > /dev/null 2>&1
# 3. PSIS passes
timeout 600 $NCU --set full --import-source on -k regex:"psis_pass_a|psis_pass_b_kernel" -s 6 -c 2 -o $O/psis_r02 -f python tools/run_psis.py 100000000 4 > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:psis -s 39 -c 26 --csv --log-file $O/psis_launches_r02.csv python tools/run_psis.py 100000000 4 > /dev/null 2>&1
# 4. exact path
timeout 600 $NCU --set full --import-source on -k regex:glm_sweep_f64 -s 2 -c 1 -o $O/f64_r02 -f python bench.py --steps 2 --warmup 3 --path f64 --no-psis --no-cpu-baseline > /dev/null 2>&1
# 5. C4: launch list of two iterations + the float64 GEMM
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/c4_launches_r02.csv python bench.py --config c4 --steps 2 --warmup 3 > $O/c4_under_ncu_r02.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"gemm_f64_kernel|hier_lik" -s 10 -c 3 -o $O/c4_r02 -f python bench.py --config c4 --steps 2 --warmup 3 > /dev/null 2>&1
ls -la $O/*r02*

#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T tools/check_sharded_step.py > gpurun_out/sharded_check_r02_final_2gpu.log 2>&1; echo "check rc=$?"; tail -6 gpurun_out/sharded_check_r02_final_2gpu.log
$T tools/run_psis_sharded.py --draws 100000000 > gpurun_out/psis_sharded_r02_final_2gpu.log 2>&1; echo "psis sharded rc=$?"; tail -6 gpurun_out/psis_sharded_r02_final_2gpu.log
$T bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r02_final2_2gpu.json 2> gpurun_out/bench_r02_final2_2gpu.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02_final2_2gpu.json

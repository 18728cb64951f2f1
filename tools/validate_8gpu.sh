#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
$T bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r02_final2_8gpu.json 2> gpurun_out/bench_r02_final2_8gpu.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_r02_final2_8gpu.json
$T tools/check_sharded_step.py > gpurun_out/sharded_check_r02_final_8gpu.log 2>&1; echo "check rc=$?"; grep -c PASS gpurun_out/sharded_check_r02_final_8gpu.log; grep -v PASS gpurun_out/sharded_check_r02_final_8gpu.log | tail -3
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513"
$T4 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r02_final2_4gpu.json 2> gpurun_out/bench_r02_final2_4gpu.err; echo "bench4 rc=$?"

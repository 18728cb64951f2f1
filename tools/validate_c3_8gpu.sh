#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514"
$T bench.py --gpus 8 --n-obs 10000000 --dim 1024 --steps 20 --warmup 3 --no-psis > gpurun_out/bench_r02_final2_c3_8gpu.json 2> gpurun_out/bench_r02_final2_c3_8gpu.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_r02_final2_c3_8gpu.json

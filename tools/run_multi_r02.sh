for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_r02_final_n$n.json 2> gpurun_out/bench_r02_final_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --n-obs 10000000 --dim 1024 --steps 20 --warmup 3 --no-psis > gpurun_out/bench_r02_c3_8gpu.json 2> gpurun_out/bench_r02_c3_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tools/check_sharded_step.py > gpurun_out/sharded_check_r02_8gpu_final.log 2>&1
tail -3 gpurun_out/sharded_check_r02_8gpu_final.log

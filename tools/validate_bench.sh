#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s); python bench.py > gpurun_out/bench_r02_final2_1gpu.json 2> gpurun_out/bench_r02_final2_1gpu.err; echo "bench rc=$? $(( $(date +%s) - S )) s"
S=$(date +%s); python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final2_reference.json 2> gpurun_out/bench_r02_final2_reference.err; echo "reference rc=$? $(( $(date +%s) - S )) s"
tail -c 600 gpurun_out/bench_r02_final2_reference.json

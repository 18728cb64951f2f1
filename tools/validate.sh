#!/bin/bash
# round-end style validation on one B200: GPU tests, smoke, the default bench line
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/validate_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/validate_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/validate_smoke.log
S=$(date +%s); python bench.py > gpurun_out/bench_validate_1gpu.json 2> gpurun_out/bench_validate_1gpu.err; echo "bench rc=$? $(( $(date +%s) - S )) s"

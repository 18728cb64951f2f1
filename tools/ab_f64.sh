#!/bin/bash
mkdir -p gpurun_out
{
echo "=== default"; python bench.py --path f64 --steps 5 --warmup 3 --no-psis --no-cpu-baseline
echo "=== pytest elbo/engine/cv"; python -m pytest tests/test_gpu_elbo.py tests/test_gpu_engine.py tests/test_gpu_cv.py -x -q -m gpu 2>&1 | tail -2
} > gpurun_out/ab_f64.log 2>&1
grep -v "^{" gpurun_out/ab_f64.log | tail -20
python - <<'PY'
import json
for line in open('gpurun_out/ab_f64.log'):
    if line.startswith('{'):
        d = json.loads(line)
        print('value %.2f iter/s  ms/step %.3f  roofline %s  clocks %s' % (d['value'], d['ms_per_step'], d.get('roofline', {}).get('frac'), d.get('clocks')))
PY

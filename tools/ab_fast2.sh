#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/validate_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/validate_pytest.log; grep "alpha fast path" gpurun_out/validate_pytest.log | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

#!/bin/bash
# last check of the committed state: GPU suite + smoke()
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests -m gpu -q > gpurun_out/pytest_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_last.log
tail -3 gpurun_out/pytest_last.log
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke_last.log

#!/bin/bash
# one gpurun call: smoke, GPU parity tests, the bench line with the warp-chunk kernels and with the first-generation ones
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench rc=$?"
cat gpurun_out/bench_v2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline'])"
TGNH_V2=0 timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --quick > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err
cat gpurun_out/bench_v1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline'])"

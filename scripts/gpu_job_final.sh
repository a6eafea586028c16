#!/bin/bash
# Final check of the round on one B200: GPU suite, smoke(), the default bench line of both arms, then a compute-sanitizer memcheck
# pass over a subset of the parity tests (last: whatever time is left).
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log | cut -c1-300
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
tail -2 gpurun_out/smoke_final.log | cut -c1-300
T0=$(date +%s)
timeout -s KILL 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s" | tee gpurun_out/bench_final.rc
T0=$(date +%s)
timeout -s KILL 200 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_final.err; echo "ref rc=$? $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/bench_final.rc
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/bench_ref_final.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "same_config", d["config"] == r["config"], "ref", r["value"])
P
timeout -s KILL 240 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "test_single_step or test_hard_wall or test_big_residues or test_residue_per_lane_reduction or test_lazy_second_kick or test_constraint_split_equals" \
    > gpurun_out/sanitizer_final.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_final.log
tail -6 gpurun_out/sanitizer_final.log | cut -c1-300

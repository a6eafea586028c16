#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout -s KILL 900 python -m pytest tests/test_sharded.py -m gpu -q -s > gpurun_out/sharded_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sharded_r2.log
tail -8 gpurun_out/sharded_r2.log | cut -c1-300
N=${1:-2}
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_n${N}_r2.json 2> gpurun_out/scale_n${N}_r2.err; echo "bench rc=$?"
tail -3 gpurun_out/scale_n${N}_r2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_n${N}_r2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}); print(d["e2e"]["value"])
c=d.get("details", d["config"]); print({k:c.get(k) for k in ("exchange","shard_check","c5_strong","replicas","exchange_timing_us","openmm_call_pattern_ms")})
PY

#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout -s KILL 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_n${N}_r2.json 2> gpurun_out/scale_n${N}_r2.err; echo "bench rc=$?"
tail -3 gpurun_out/scale_n${N}_r2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_n${N}_r2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches")}); print("e2e", d["e2e"]["value"])
c=d.get("details", d["config"]); print(json.dumps({k:c.get(k) for k in ("exchange","shard_check","c5_strong","replicas","exchange_timing_us")}, indent=0))
PY

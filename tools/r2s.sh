#!/bin/bash
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2s_bench_n$N.json 2> gpurun_out/r2s_bench_n$N.err; echo rc_bench=$?
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2s_bench_n$N.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches","clocks")})
    print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"])
    for k,v in d.get("extra", {}).items():
        print(k, v if not isinstance(v, dict) else {kk: v[kk] for kk in v if kk in ("value","ms_per_step","error")}, (v.get("config") or {}).get("ms_per_step with C left column-distributed") if isinstance(v, dict) else "")
except Exception as e:
    print("no line", e)
PY
grep -v "^W10\|^\*\*\*\|OMP_NUM" gpurun_out/r2s_bench_n$N.err | tail -5

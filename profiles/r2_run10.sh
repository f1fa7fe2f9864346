#!/bin/bash
mkdir -p gpurun_out
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 1 --warmup 1 --no-shapes) > gpurun_out/r2_bench_${NG:-2}gpu.json 2> gpurun_out/r2_bench_${NG:-2}gpu.err
tail -n 5 gpurun_out/r2_bench_${NG:-2}gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_${NG:-2}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step','suite')}); print(d['e2e'])
PY

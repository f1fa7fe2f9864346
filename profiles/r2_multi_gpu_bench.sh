#!/bin/bash
# NG=2|4|8 bash profiles/r2_multi_gpu_bench.sh   (under `gpurun --gpus $NG`): the bench line incl. the suite leg
mkdir -p gpurun_out
NG=${NG:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $NG --steps 1 --warmup 1 --no-shapes > gpurun_out/r2_bench_${NG}gpu.json 2> gpurun_out/r2_bench_${NG}gpu.err
tail -n 3 gpurun_out/r2_bench_${NG}gpu.err

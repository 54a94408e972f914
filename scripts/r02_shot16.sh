#!/bin/bash
# N=2: number of export CTAs in the single-launch partitioned vmult
mkdir -p gpurun_out
for e in 16 32 128 296; do
  EXADG_B200_EXPORT_CTAS=$e timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s16_n2_e$e.json 2> gpurun_out/r02_s16_n2_e$e.err
  python -c "import json;d=json.loads(open('gpurun_out/r02_s16_n2_e$e.json').read().strip().splitlines()[-1]);print('export ctas $e',d['value']/1e9,d['ms_per_step'])"
done

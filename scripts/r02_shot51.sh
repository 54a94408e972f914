#!/bin/bash
# N GPUs: bench line of the final tree (partitioned direct e2e variant included)
mkdir -p gpurun_out
N=${1:-4}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_s51_bench_n${N}.json 2> gpurun_out/r02_s51_bench_n${N}.err
tail -n 3 gpurun_out/r02_s51_bench_n${N}.err
python -c "import json;d=json.loads(open('gpurun_out/r02_s51_bench_n${N}.json').read().strip().splitlines()[-1]);print('n$N',d['value']/1e9,d['ms_per_step'],d['config']['workload'][:90],d['config']['invariants'], 'e2e', d['e2e']['value']/1e9, d['e2e']['sequential_dofs_per_s']/1e9, d['e2e']['pipelined_dofs_per_s'])"

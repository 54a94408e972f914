#!/bin/bash
# N GPUs: curved bench line of the final tree (config 5 at N=8), Cartesian line at N=4 on the 160^3 grid
mkdir -p gpurun_out
N=${1:-8}
if [ "$N" = "8" ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 3 --mesh curvilinear --cells 192 > gpurun_out/r02_s37_bench_curved_n8.json 2> gpurun_out/r02_s37_bench_curved_n8.err
  python -c "import json;d=json.loads(open('gpurun_out/r02_s37_bench_curved_n8.json').read().strip().splitlines()[-1]);print('curved n8',d['value']/1e9,d['ms_per_step'],d['roofline']['frac'],d['config']['invariants'])"
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_s37_bench_n${N}.json 2> gpurun_out/r02_s37_bench_n${N}.err
  python -c "import json;d=json.loads(open('gpurun_out/r02_s37_bench_n${N}.json').read().strip().splitlines()[-1]);print('n$N',d['value']/1e9,d['ms_per_step'],d['config']['workload'][:90],d['config']['invariants'], 'e2e', d['e2e']['value']/1e9)"
fi

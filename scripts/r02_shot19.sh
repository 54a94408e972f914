#!/bin/bash
# 8 GPUs: Cartesian bench line (BASELINE config 5 size: 192^3 cells, 884.7 M DoFs) and the curved mesh of config 5
mkdir -p gpurun_out
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_s19_bench_n${N}.json 2> gpurun_out/r02_s19_bench_n${N}.err
echo "rc $?"
python -c "import json;d=json.loads(open('gpurun_out/r02_s19_bench_n${N}.json').read().strip().splitlines()[-1]);print('n$N',d['value']/1e9,d['ms_per_step'],d['config']['workload'][:90],d['config']['invariants'], 'e2e', d['e2e']['value']/1e9, d['clocks'])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 --mesh curvilinear --cells 192 > gpurun_out/r02_s19_bench_curved_n${N}.json 2> gpurun_out/r02_s19_bench_curved_n${N}.err
echo "rc $?"
python -c "import json;d=json.loads(open('gpurun_out/r02_s19_bench_curved_n${N}.json').read().strip().splitlines()[-1]);print('curved n$N',d['value']/1e9,d['ms_per_step'],d['roofline']['frac'],d['config']['workload'][:90],d['config']['invariants'])"
tail -2 gpurun_out/r02_s19_bench_curved_n${N}.err

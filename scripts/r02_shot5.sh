#!/bin/bash
# 8 GPUs: partitioned parity (oracle) on 8 ranks, bench N=8 and N=4 (fused single launch)
mkdir -p gpurun_out
N=${1:-8}
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py > gpurun_out/r02_s5_mgpu$N.log 2>&1; echo "rc $?" >> gpurun_out/r02_s5_mgpu$N.log )
tail -n 25 gpurun_out/r02_s5_mgpu$N.log
for n in $N 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 50 --warmup 5 --no-tune > gpurun_out/r02_s5_bench_n$n.json 2> gpurun_out/r02_s5_bench_n$n.err
  echo "rc $?"
  python -c "import json;d=json.loads(open('gpurun_out/r02_s5_bench_n$n.json').read().strip().splitlines()[-1]);print('n$n',d['value']/1e9,d['ms_per_step'],d['config']['invariants'], d['e2e']['value']/1e9)"
done

#!/bin/bash
# N GPUs: Cartesian bench line of the current tree, curved bench line (config 5 is a curved mesh), multi-GPU parity tests
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_s14_bench_n${N}.json 2> gpurun_out/r02_s14_bench_n${N}.err
echo "rc $?"
python -c "import json;d=json.loads(open('gpurun_out/r02_s14_bench_n${N}.json').read().strip().splitlines()[-1]);print('cartesian n$N',d['value']/1e9,d['ms_per_step'],d['config']['invariants'], d['e2e']['value']/1e9)"
CELLS=$(python -c "print({1:64,2:80,4:96,8:128}[$N])")
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 --mesh curvilinear --cells $CELLS > gpurun_out/r02_s14_bench_curved_n${N}.json 2> gpurun_out/r02_s14_bench_curved_n${N}.err
echo "rc $?"
python -c "import json;d=json.loads(open('gpurun_out/r02_s14_bench_curved_n${N}.json').read().strip().splitlines()[-1]);print('curved n$N',d['value']/1e9,d['ms_per_step'],d['roofline']['frac'],d['config']['invariants'])"
tail -2 gpurun_out/r02_s14_bench_curved_n${N}.err
if [ "$N" = "2" ]; then
  ( timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02_s14_pytest_multi.log 2>&1; echo "rc $?" >> gpurun_out/r02_s14_pytest_multi.log )
  tail -n 4 gpurun_out/r02_s14_pytest_multi.log
fi

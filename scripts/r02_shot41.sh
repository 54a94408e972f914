#!/bin/bash
# 2 GPUs: partitioned direct variant of the host-buffer vmult (parity through tests/multi_gpu_check.py), then the N=2 bench line
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k default_kernels > gpurun_out/r02_s41_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s41_pytest.log )
tail -n 25 gpurun_out/r02_s41_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_s41_bench_n2.json 2> gpurun_out/r02_s41_bench_n2.err
tail -n 3 gpurun_out/r02_s41_bench_n2.err
python -c "import json;d=json.loads(open('gpurun_out/r02_s41_bench_n2.json').read().strip().splitlines()[-1]);print('n2',d['value']/1e9,d['ms_per_step'], 'e2e', d['e2e'])"

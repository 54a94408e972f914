#!/bin/bash
# 2 GPUs: bench N=2 fused single launch vs multi-launch (after the room-for-the-put-kernel fix)
mkdir -p gpurun_out
for mode in fused nofused; do
  if [ $mode = nofused ]; then export EXADG_B200_NO_FUSED_HALO=1; else unset EXADG_B200_NO_FUSED_HALO; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --no-tune > gpurun_out/r02_s4_bench_n2_$mode.json 2> gpurun_out/r02_s4_bench_n2_$mode.err
  echo "rc $?"
  python -c "import json;d=json.loads(open('gpurun_out/r02_s4_bench_n2_$mode.json').read().strip().splitlines()[-1]);print('$mode',d['value']/1e9,d['ms_per_step'],d['config']['invariants'])"
done
unset EXADG_B200_NO_FUSED_HALO
( timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k default > gpurun_out/r02_s4_pytest_multi.log 2>&1; echo "rc $?" >> gpurun_out/r02_s4_pytest_multi.log )
tail -n 3 gpurun_out/r02_s4_pytest_multi.log

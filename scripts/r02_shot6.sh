#!/bin/bash
# N GPUs: dynamic item claiming + 64 export CTAs vs static striding (every CTA exports); parity check
mkdir -p gpurun_out
N=${1:-2}
for mode in dynamic static; do
  if [ $mode = static ]; then export EXADG_B200_STATIC_ITEMS=1; else unset EXADG_B200_STATIC_ITEMS; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --no-tune > gpurun_out/r02_s6_bench_n${N}_$mode.json 2> gpurun_out/r02_s6_bench_n${N}_$mode.err
  echo "rc $?"
  python -c "import json;d=json.loads(open('gpurun_out/r02_s6_bench_n${N}_$mode.json').read().strip().splitlines()[-1]);print('$mode n$N',d['value']/1e9,d['ms_per_step'],d['config']['invariants'])"
done
unset EXADG_B200_STATIC_ITEMS
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py > gpurun_out/r02_s6_mgpu$N.log 2>&1; echo "rc $?" >> gpurun_out/r02_s6_mgpu$N.log )
tail -n 3 gpurun_out/r02_s6_mgpu$N.log

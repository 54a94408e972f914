#!/bin/bash
# 2 GPUs: partitioned parity tests (oracle), bench N=2 fused single launch vs multi-launch, N=1 tournament with the WP variants
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02_s3_pytest_multi.log 2>&1; echo "rc $?" >> gpurun_out/r02_s3_pytest_multi.log )
tail -n 25 gpurun_out/r02_s3_pytest_multi.log
for mode in fused nofused; do
  if [ $mode = nofused ]; then export EXADG_B200_NO_FUSED_HALO=1; else unset EXADG_B200_NO_FUSED_HALO; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --no-tune > gpurun_out/r02_s3_bench_n2_$mode.json 2> gpurun_out/r02_s3_bench_n2_$mode.err
  python -c "import json;d=json.loads(open('gpurun_out/r02_s3_bench_n2_$mode.json').read().strip().splitlines()[-1]);print('$mode',d['value']/1e9,d['ms_per_step'])"
done
unset EXADG_B200_NO_FUSED_HALO
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-tune --no-cpu > gpurun_out/r02_s3_bench_n1.json 2> gpurun_out/r02_s3_bench_n1.err
python -c "import json;d=json.loads(open('gpurun_out/r02_s3_bench_n1.json').read().strip().splitlines()[-1]);print('n1',d['value']/1e9,d['ms_per_step'])"
TOURNAMENT_SKIP2=1 timeout 120 build/ws_tournament -1 3 5 50 > gpurun_out/r02_s3_tournament.log 2>&1; echo "rc $?" >> gpurun_out/r02_s3_tournament.log
cat gpurun_out/r02_s3_tournament.log

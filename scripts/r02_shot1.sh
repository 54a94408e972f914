#!/bin/bash
# Round-2 baseline shot (1 GPU): GPU test-suite, kernel variants side by side, ncu capture of the timed k=4 kernels at 96^3,
# launch list of the bench command, one bench line.   gpurun --timeout 1500 -- 'bash scripts/r02_shot1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_pytest.log ) 
tail -n 5 gpurun_out/r02_pytest.log
timeout 120 build/ws_tournament -1 3 5 50 > gpurun_out/r02_tournament_96.log 2>&1; echo "rc $?" >> gpurun_out/r02_tournament_96.log
cat gpurun_out/r02_tournament_96.log
for v in 1 3; do
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:ws_kernel -s 3 -c 1 -f -o gpurun_out/r02_cart_ws_v${v}_k4_96 build/ws_tournament $v 3 5 3 > gpurun_out/r02_ncu_ws_v${v}.log 2>&1
  tail -n 2 gpurun_out/r02_ncu_ws_v${v}.log
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-tune --e2e-api plain > gpurun_out/r02_launches_bench.log 2>&1
timeout 400 python bench.py --steps 50 --warmup 5 --no-tune > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cat gpurun_out/r02_bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
cat gpurun_out/r02_bench_ref.json

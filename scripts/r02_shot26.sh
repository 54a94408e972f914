#!/bin/bash
# 1 GPU: whole GPU suite on the final tree, default bench line, launch list of the bench command under ncu
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_s26_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s26_pytest.log )
tail -n 6 gpurun_out/r02_s26_pytest.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_s26_bench_n1.json 2> gpurun_out/r02_s26_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_s26_bench_n1.json').read().strip().splitlines()[-1])
print('n1', d['value']/1e9, d['ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['value']/1e9, d['clocks'])
print(json.dumps(d.get('callers'))[:600])
print(json.dumps(d.get('applications'))[:900])
"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_s26_bench_ref.json 2> gpurun_out/r02_s26_bench_ref.err
tail -c 700 gpurun_out/r02_s26_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_s26_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s26_ncu.log 2>&1
tail -n 12 gpurun_out/r02_s26_launches.csv | cut -c1-200

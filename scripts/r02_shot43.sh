#!/bin/bash
# final-tree bench line (N=1) and reference arm; ncu launch list of the same command
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_s43_bench_n1.json 2> gpurun_out/r02_s43_bench_n1.err
tail -n 2 gpurun_out/r02_s43_bench_n1.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_s43_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value']/1e9,'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value']/1e9,d['e2e']['sequential_dofs_per_s']/1e9,d['e2e']['pipelined_dofs_per_s'],'cpu',d['cpu_baseline']['value']/1e9)
print(d['e2e']['api']); print({k:(v.get('ms') if isinstance(v,dict) else v) for k,v in d.get('callers',{}).items()})"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_s43_bench_ref.json 2> gpurun_out/r02_s43_bench_ref.err
tail -c 600 gpurun_out/r02_s43_bench_ref.json

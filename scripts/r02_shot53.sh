#!/bin/bash
# final tree: GPU suite, smoke(), bench line + ncu launch list of the same command (shares of the step)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s53_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s53_pytest.log )
tail -n 3 gpurun_out/r02_s53_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 2
timeout 300 python bench.py > gpurun_out/r02_s53_bench_n1.json 2> gpurun_out/r02_s53_bench_n1.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_s53_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value']/1e9,'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value']/1e9,d['e2e']['pipelined_dofs_per_s'],'cpu',d['cpu_baseline']['value']/1e9)
print(d['applications']['ins_operators']['viscous_helmholtz']); print(d['applications']['poisson_solve']['solve_ms'], d['callers']['cg_iteration']['ms'])"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_s53_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s53_ncu.log 2>&1
grep -c vmult_cartesian_ws_kernel gpurun_out/r02_s53_launches.csv

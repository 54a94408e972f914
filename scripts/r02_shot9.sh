#!/bin/bash
# 1 GPU: multigrid tests (f-1), singular-system test (f-2)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_rhs_error.py -q -x > gpurun_out/r02_s9_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s9_pytest.log )
tail -n 40 gpurun_out/r02_s9_pytest.log

#!/bin/bash
# 2 GPUs: CG + p-multigrid on a partition (multi_gpu_check.py)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k default_kernels > gpurun_out/r02_s42_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s42_pytest.log )
grep -v "^$" gpurun_out/r02_s42_pytest.log | tail -n 30

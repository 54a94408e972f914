#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_helmholtz.py -q -x > gpurun_out/r02_s54_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s54_pytest.log )
tail -n 6 gpurun_out/r02_s54_pytest.log

#!/bin/bash
# 1 GPU: new f-1 / f-3 tests first, then the whole GPU suite
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_helmholtz.py -q > gpurun_out/r02_s10_new.log 2>&1; echo "rc $?" >> gpurun_out/r02_s10_new.log )
tail -n 60 gpurun_out/r02_s10_new.log
( timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multigrid.py --deselect tests/test_gpu_helmholtz.py > gpurun_out/r02_s10_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s10_pytest.log )
tail -n 15 gpurun_out/r02_s10_pytest.log

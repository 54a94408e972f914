#!/bin/bash
# ncu --set full after the round-2 changes: k=3 plane kernel (swizzled), k=5 and k=7 line kernels with 8-cell batches
mkdir -p gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:vmult_cartesian_kernel -s 2 -c 1 -f -o gpurun_out/r02_cart_k3_64_v2 python bench.py --degree 3 --cells 64 --steps 2 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s47_k3.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:vmult_cartesian_line_kernel -s 2 -c 1 -f -o gpurun_out/r02_cart_k5_40_v2 python bench.py --degree 5 --cells 40 --steps 2 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s47_k5.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:vmult_cartesian_line_kernel -s 2 -c 1 -f -o gpurun_out/r02_cart_k7_32_v2 python bench.py --degree 7 --cells 32 --steps 2 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s47_k7.log 2>&1
ls -la gpurun_out/*_v2.ncu-rep

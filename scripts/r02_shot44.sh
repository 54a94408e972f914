#!/bin/bash
# ncu --set full of the k=3 plane kernel and the k=5 line kernel (the degrees furthest below the roofline)
mkdir -p gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:vmult_cartesian_kernel -s 2 -c 1 -f -o gpurun_out/r02_cart_k3_64 python bench.py --degree 3 --cells 64 --steps 2 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s44_k3.log 2>&1
tail -n 2 gpurun_out/r02_s44_k3.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:vmult_cartesian_line_kernel -s 2 -c 1 -f -o gpurun_out/r02_cart_k5_40 python bench.py --degree 5 --cells 40 --steps 2 --warmup 3 --no-cpu --no-callers --no-fp64-peak --e2e-api plain > gpurun_out/r02_s44_k5.log 2>&1
tail -n 2 gpurun_out/r02_s44_k5.log
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 400 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2c14_bench_c5.json 2> gpurun_out/r2c14_bench_c5.err; echo "rc=$?"; head -c 300 gpurun_out/r2c14_bench_c5.json; tail -2 gpurun_out/r2c14_bench_c5.err
(timeout 300 python scripts/bench_configs.py --which c5 --n-kmeans 5000000; B2_KM_TIMING=1 timeout 300 python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1) > gpurun_out/r2c14_c5.jsonl 2>&1
cut -c1-330 gpurun_out/r2c14_c5.jsonl

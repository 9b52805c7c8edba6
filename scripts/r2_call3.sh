#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_search.py -q -x > gpurun_out/r2c3_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c3_tests.log
tail -15 gpurun_out/r2c3_tests.log
(B2_KM_TIMING=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000) > gpurun_out/r2c3_c5.jsonl 2>&1
cut -c1-500 gpurun_out/r2c3_c5.jsonl
timeout 900 python scripts/large_k_timing.py > gpurun_out/r2c3_large_k.jsonl 2>&1
cat gpurun_out/r2c3_large_k.jsonl | cut -c1-400
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_kmeans.py --deselect tests/test_gpu_search.py > gpurun_out/r2c3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c3_pytest.log
tail -5 gpurun_out/r2c3_pytest.log

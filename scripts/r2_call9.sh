#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_ops.py -q > gpurun_out/r2c9_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c9_pytest.log
tail -3 gpurun_out/r2c9_pytest.log | cut -c1-250
for lb in 4 8 16; do
B2_KM_LANE_BYTES=$lb B2_KM_DEBUG=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1 --c5-niter 5 > gpurun_out/r2c9_kmdbg_bulk_$lb.log 2>&1
echo "bulk LB=$lb"; grep "accumulate dbg" gpurun_out/r2c9_kmdbg_bulk_$lb.log | tail -1 | cut -c1-300
B2_KM_LANE_BYTES=$lb B2_KM_TIMING=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1 > gpurun_out/r2c9_km_bulk_$lb.log 2>&1
grep "kmeans timing" gpurun_out/r2c9_km_bulk_$lb.log | cut -c100-300
done
timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000 > gpurun_out/r2c9_c5.jsonl 2>&1
cut -c1-250 gpurun_out/r2c9_c5.jsonl

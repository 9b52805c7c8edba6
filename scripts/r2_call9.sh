#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 900 python -m pytest tests/test_gpu_kmeans.py -q > gpurun_out/r2c9_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c9_pytest.log
tail -3 gpurun_out/r2c9_pytest.log | cut -c1-250
for cfg in "4 8" "4 32" "8 32" "16 8" "16 32"; do
set -- $cfg
B2_KM_LANE_BYTES=$1 B2_KM_RING_GROUPS=$2 B2_KM_DEBUG=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1 --c5-niter 5 > gpurun_out/r2c9_kmdbg_$1_$2.log 2>&1
echo "LB=$1 groups=$2"; grep "accumulate dbg" gpurun_out/r2c9_kmdbg_$1_$2.log | tail -1 | cut -c1-300
B2_KM_LANE_BYTES=$1 B2_KM_RING_GROUPS=$2 B2_KM_TIMING=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1 > gpurun_out/r2c9_km_$1_$2.log 2>&1
grep "kmeans timing" gpurun_out/r2c9_km_$1_$2.log | cut -c100-300
done

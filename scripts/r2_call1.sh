#!/bin/bash
# round-2 GPU call 1: validate the three opt-in kernels, measure them against the defaults, operator scope
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
python -c "import faiss; print('faiss', faiss.__version__)" > gpurun_out/r2c1_faiss_probe.txt 2>&1
nproc >> gpurun_out/r2c1_faiss_probe.txt
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -3 gpurun_out/r2c1_pytest.log
(timeout 300 python scripts/bench_configs.py --which c2; B2_F32_SPLIT=1 timeout 300 python scripts/bench_configs.py --which c2) > gpurun_out/r2c1_c2.jsonl 2>&1
(B2_KM_TIMING=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000; B2_FILTER_TOP1=1 B2_KM_TIMING=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000) > gpurun_out/r2c1_c5.jsonl 2>&1
(B2_PAIR_2CTA=1 timeout 500 python scripts/pair_sched_exp.py) > gpurun_out/r2c1_pair.jsonl 2>&1
timeout 600 python scripts/operator_scope.py > gpurun_out/r2c1_opscope.jsonl 2>&1
tail -n 4 gpurun_out/r2c1_c2.jsonl gpurun_out/r2c1_c5.jsonl gpurun_out/r2c1_pair.jsonl gpurun_out/r2c1_opscope.jsonl

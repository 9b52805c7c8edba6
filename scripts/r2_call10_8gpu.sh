#!/bin/bash
# 8-GPU call: multi-GPU correctness log + the headline bench at N=8 + C5 / C4 at their BASELINE shape (8 x B200)
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 400 $TR scripts/multigpu_check.py > gpurun_out/r2c10_multigpu_check_8gpu.log 2>&1; echo "multigpu rc=$?" | tee -a gpurun_out/r2c10_multigpu_check_8gpu.log
grep -E "rank 0\]|MULTIGPU" gpurun_out/r2c10_multigpu_check_8gpu.log | head -8
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2c10_bench_n8.json 2> gpurun_out/r2c10_bench_n8.err; echo "bench n8 rc=$?"
tail -c 2200 gpurun_out/r2c10_bench_n8.json; tail -2 gpurun_out/r2c10_bench_n8.err
B2_SHARD_STAGED=0 timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --parity-queries 32 > gpurun_out/r2c10_bench_n8_unstaged.json 2> gpurun_out/r2c10_bench_n8_unstaged.err; echo "bench n8 unstaged rc=$?"
head -c 400 gpurun_out/r2c10_bench_n8_unstaged.json
timeout 600 $TR bench.py --config c5 --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2c10_bench_c5_n8.json 2> gpurun_out/r2c10_bench_c5_n8.err; echo "c5 n8 rc=$?"
head -c 500 gpurun_out/r2c10_bench_c5_n8.json; tail -2 gpurun_out/r2c10_bench_c5_n8.err
timeout 600 $TR bench.py --config c4 --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2c10_bench_c4_n8.json 2> gpurun_out/r2c10_bench_c4_n8.err; echo "c4 n8 rc=$?"
head -c 500 gpurun_out/r2c10_bench_c4_n8.json; tail -2 gpurun_out/r2c10_bench_c4_n8.err

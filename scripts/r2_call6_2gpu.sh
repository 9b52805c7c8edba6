#!/bin/bash
# 2-GPU call: multi-GPU correctness logs + N=2 bench + the 8-GPU shard shape emulated on 2 ranks
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/multigpu_check.py > gpurun_out/r2c6_multigpu_check_2gpu.log 2>&1; echo "multigpu rc=$?" | tee -a gpurun_out/r2c6_multigpu_check_2gpu.log
grep -E "OK|MISMATCH|FAILED|Error" gpurun_out/r2c6_multigpu_check_2gpu.log | head -20
timeout 600 python scripts/multidevice_check.py > gpurun_out/r2c6_multidevice_check_2gpu.log 2>&1; echo "multidevice rc=$?" | tee -a gpurun_out/r2c6_multidevice_check_2gpu.log
tail -2 gpurun_out/r2c6_multidevice_check_2gpu.log | cut -c1-600
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c6_bench_n2.json 2> gpurun_out/r2c6_bench_n2.err; echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/r2c6_bench_n2.json; tail -3 gpurun_out/r2c6_bench_n2.err
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --n 250000 --parity-queries 64 > gpurun_out/r2c6_bench_n2_shardshape.json 2> gpurun_out/r2c6_bench_n2_shardshape.err; echo "bench shardshape rc=$?"
tail -c 1800 gpurun_out/r2c6_bench_n2_shardshape.json; tail -3 gpurun_out/r2c6_bench_n2_shardshape.err
timeout 900 $TR bench.py --config c5 --gpus 2 --steps 2 --warmup 3 --c5-n 2000000 > gpurun_out/r2c6_bench_c5_n2.json 2> gpurun_out/r2c6_bench_c5_n2.err; echo "c5 n2 rc=$?"
tail -c 1500 gpurun_out/r2c6_bench_c5_n2.json; tail -3 gpurun_out/r2c6_bench_c5_n2.err
timeout 900 $TR bench.py --config c4 --gpus 2 --steps 2 --warmup 3 --c4-n 3000000 > gpurun_out/r2c6_bench_c4_n2.json 2> gpurun_out/r2c6_bench_c4_n2.err; echo "c4 n2 rc=$?"
tail -c 1500 gpurun_out/r2c6_bench_c4_n2.json; tail -3 gpurun_out/r2c6_bench_c4_n2.err

#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 3 --parity-queries 64 > gpurun_out/r2c12_bench_n8.json 2> gpurun_out/r2c12_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c12_bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "phases", d["shard_phases_ms"], "kernel/rank", d["roofline"]["kernel_ms_per_rank"], "parity", d["parity"]["idx_bit_exact_vs_oracle"], d["parity"]["score_bit_exact_vs_oracle"])
PY

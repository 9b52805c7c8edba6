#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c11_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c11_pytest.log
tail -6 gpurun_out/r2c11_pytest.log | cut -c1-250
for tp in 1 0; do
B2_FILTER_TWO_PHASE=$tp timeout 600 python bench.py --steps 10 --warmup 3 --corpus-rows 125000 --no-cpu-baseline --no-operator --parity-queries 64 > gpurun_out/r2c11_bench_shard125k_tp$tp.json 2> gpurun_out/r2c11_bench_shard_tp$tp.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c11_bench_shard125k_tp$tp.json").read().strip().splitlines()[-1])
print("two_phase=$tp shard 125k: ms_per_step", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],3), "parity", d["parity"]["idx_bit_exact_vs_oracle"], d["parity"]["score_bit_exact_vs_oracle"], "fallback", d["fallback_queries"])
PY
done
B2_FILTER_TWO_PHASE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-operator --parity-queries 32 > gpurun_out/r2c11_bench_n1_tp0.json 2> gpurun_out/r2c11_bench_n1_tp0.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c11_bench_n1.json 2> gpurun_out/r2c11_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("n1_tp0","n1"):
    d=json.loads(open(f"gpurun_out/r2c11_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "kernel", round(d["roofline"]["kernel_ms"],2), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "op", (d.get("operator_e2e") or {}).get("value"), "parity", d["parity"]["idx_bit_exact_vs_oracle"], "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2c11_bench_ref.json 2> gpurun_out/r2c11_bench_ref.err; head -c 300 gpurun_out/r2c11_bench_ref.json

#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 600 python -m pytest tests/test_gpu_search.py -q -k "two_phase" > gpurun_out/r2c15_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c15_pytest.log
tail -5 gpurun_out/r2c15_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-operator > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c15_bench.json").read().strip().splitlines()[-1])
p=d["parity"]; print(round(d["value"]), round(d["ms_per_step"],2), p["queries"], p["sample"], p["idx_bit_exact_vs_oracle"], p["score_bit_exact_vs_oracle"], p["rows_differing"])
PY

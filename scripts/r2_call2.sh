#!/bin/bash
# round-2 GPU call 2: new Lloyd engine (parity + timing), whole gpu suite, launch list of a k-means run
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 900 python -m pytest tests/test_gpu_kmeans.py -q -x > gpurun_out/r2c2_kmeans_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c2_kmeans_tests.log
tail -15 gpurun_out/r2c2_kmeans_tests.log
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_kmeans.py > gpurun_out/r2c2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c2_pytest.log
tail -5 gpurun_out/r2c2_pytest.log
(timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000; B2_KM_TIMING=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000) > gpurun_out/r2c2_c5.jsonl 2>&1
cut -c1-600 gpurun_out/r2c2_c5.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2c2_km_launches.csv python scripts/bench_configs.py --which c5 --n-kmeans 5000000 > gpurun_out/r2c2_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2c2_km_launches.csv")) if len(r) > 5]
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    try: v = float(d["Metric Value"].replace(",", ""))
    except Exception: continue
    name = d["Kernel Name"][:60]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/1e6:10.3f} ms {100*t/tot:5.1f}%  x{c:4d}  {n}")
PY

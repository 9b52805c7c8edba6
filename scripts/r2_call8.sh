#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_ops.py -q > gpurun_out/r2c8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c8_pytest.log
tail -6 gpurun_out/r2c8_pytest.log | cut -c1-250
(B2_KM_TIMING=1 timeout 400 python scripts/bench_configs.py --which c5 --n-kmeans 5000000) > gpurun_out/r2c8_c5.jsonl 2>&1
cut -c1-330 gpurun_out/r2c8_c5.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c8_km_launches.csv python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1 --c5-niter 6 > gpurun_out/r2c8_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2c8_km_launches.csv")) if len(r) > 5]
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
    a = agg.setdefault(name, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(round(v/1e6,3))
tot = sum(a[1] for a in agg.values())
for n, (c, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"{t/1e6:10.3f} ms {100*t/tot:5.1f}%  x{c:4d}  {n}  {l[:8]}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:km_accumulate_vec -s 4 -c 1 -o gpurun_out/r2c8_prof_accumulate python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1 --c5-niter 6 > gpurun_out/r2c8_ncu2.log 2>&1
tail -1 gpurun_out/r2c8_ncu2.log

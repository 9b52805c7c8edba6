#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c13_pytest.log
tail -3 gpurun_out/r2c13_pytest.log | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-operator --parity-queries 0 > gpurun_out/r2c13_ncu1.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_bench_steps2.csv")) if len(r) > 5]
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    try: v = float(d["Metric Value"].replace(",", ""))
    except Exception: continue
    name = d["Kernel Name"][:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"{t/1e6:10.3f} ms {100*t/tot:5.1f}%  x{c:4d}  {n}")
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn_filter_kernel -s 3 -c 1 -o gpurun_out/r2_prof_filter_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-operator --parity-queries 0 > gpurun_out/r2c13_ncu2.log 2>&1
tail -2 gpurun_out/r2c13_ncu2.log

#!/bin/bash
mkdir -p gpurun_out
export B2_EXPECT_GPU=1
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2c4_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c4_pytest.log
tail -12 gpurun_out/r2c4_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c4_bench_n1.json 2> gpurun_out/r2c4_bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2c4_bench_n1.json; tail -5 gpurun_out/r2c4_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2c4_bench_ref.json 2> gpurun_out/r2c4_bench_ref.err; echo "ref rc=$?"
tail -c 2500 gpurun_out/r2c4_bench_ref.json; tail -5 gpurun_out/r2c4_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c4_km_launches.csv python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1 --c5-niter 6 > gpurun_out/r2c4_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2c4_km_launches.csv")) if len(r) > 5]
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
for n, (c, t, l) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{t/1e6:10.3f} ms {100*t/tot:5.1f}%  x{c:4d}  {n}  {l[:8]}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:km_accumulate_vec -s 2 -c 1 -o gpurun_out/r2c4_prof_accumulate python scripts/bench_configs.py --which c5 --n-kmeans 5000000 --c5-modes 1 --c5-niter 4 > gpurun_out/r2c4_ncu2.log 2>&1
tail -3 gpurun_out/r2c4_ncu2.log
timeout 600 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2c4_bench_c5.json 2> gpurun_out/r2c4_bench_c5.err; tail -c 1500 gpurun_out/r2c4_bench_c5.json; tail -3 gpurun_out/r2c4_bench_c5.err
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 > gpurun_out/r2c4_bench_c2.json 2> gpurun_out/r2c4_bench_c2.err; tail -c 1500 gpurun_out/r2c4_bench_c2.json; tail -3 gpurun_out/r2c4_bench_c2.err
timeout 900 python bench.py --config c4 --steps 2 --warmup 3 > gpurun_out/r2c4_bench_c4.json 2> gpurun_out/r2c4_bench_c4.err; tail -c 1500 gpurun_out/r2c4_bench_c4.json; tail -3 gpurun_out/r2c4_bench_c4.err

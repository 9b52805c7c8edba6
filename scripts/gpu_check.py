"""Staged GPU bring-up check: each stage runs in its own subprocess (a sticky CUDA error or a trap in one
stage must not hide the others). Usage on the GPU box:
    python scripts/gpu_check.py            # all stages, logs to gpurun_out/check_*.log
    python scripts/gpu_check.py --stage 3  # one stage in-process
"""
import argparse
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def mk(n, d, seed, grid=False, normalize=True):
    rng = np.random.default_rng(seed)
    if grid:
        return (rng.integers(-8, 9, size=(n, d)).astype(np.float32) / 64).astype(np.float32)
    x = rng.standard_normal((n, d), dtype=np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def compare(tag, D, I, Do, Io):
    okI = np.array_equal(I, Io)
    okD = np.array_equal(D.view(np.uint32), Do.view(np.uint32))
    rows_bad = int((I != Io).any(axis=1).sum())
    maxd = float(np.max(np.abs(D.astype(np.float64) - Do.astype(np.float64)))) if D.size else 0.0
    print(f"[{tag}] idx_exact={okI} score_bits_exact={okD} bad_rows={rows_bad}/{len(I)} max|dD|={maxd:.3e}", flush=True)
    if not okI:
        r = int(np.argmax((I != Io).any(axis=1)))
        print("   first bad row", r, "\n   got ", I[r][:16], D[r][:8], "\n   want", Io[r][:16], Do[r][:8], flush=True)
    return okI and okD


def run_case(tag, n, d, nq, k, metric, dtype, grid=False, ids=None, oracle_rows=None):
    import oracle
    from lotus_b200 import _native as nv
    x = mk(n, d, 0, grid)
    q = mk(nq, d, 1, grid)
    if dtype == "bf16":
        xb, qb = nv.f32_to_bf16_bits(x), nv.f32_to_bf16_bits(q)
        xf, qf = nv.bf16_bits_to_f32(xb), nv.bf16_bits_to_f32(qb)
        idx = nv.Index(xb, nv.BF16, metric)
        qarg, qdt = qb, nv.BF16
    else:
        xf, qf = x, q
        idx = nv.Index(x, nv.F32, metric)
        qarg, qdt = q, nv.F32
    nv.stats_reset()
    t0 = time.time()
    D, I = idx.search(qarg, k, qdt, ids=ids)
    t1 = time.time()
    st = nv.stats()
    sel = slice(None) if oracle_rows is None else slice(0, oracle_rows)
    if ids is None:
        Do, Io = oracle.knn(xf, qf[sel], k, metric)
    else:
        Do, Io = oracle.knn_subset(xf, qf[sel], k, ids, metric)
    ok = compare(tag, D[sel], I[sel], Do, Io)
    print(f"   n={n} d={d} nq={nq} k={k} metric={metric} dtype={dtype} wall={1e3*(t1-t0):.1f}ms filter_ms={idx.last_filter_ms():.3f} stats={st}", flush=True)
    idx.close()
    return ok


def stage(s):
    from lotus_b200 import _native as nv
    print("devices", nv.device_count(), flush=True)
    ok = True
    if s == 0:   # dense path only (n < 512): tie-heavy grid data, all tie rules
        for metric in (0, 1):
            for k in (1, 2, 5, 17, 64, 400):
                ok &= run_case(f"dense grid m{metric} k{k}", 300, 16, 40, k, metric, "f32", grid=True)
        ok &= run_case("dense gauss", 400, 40, 33, 7, 0, "f32")
    elif s == 1:  # first tcgen05 filter run: bf16 IP
        ok &= run_case("filter bf16 IP k5", 5000, 64, 300, 5, 0, "bf16")
        ok &= run_case("filter bf16 IP k32", 5000, 64, 300, 32, 0, "bf16")
        ok &= run_case("filter bf16 IP k10 d768", 20000, 768, 500, 10, 0, "bf16")
    elif s == 2:  # L2 epilogue
        ok &= run_case("filter bf16 L2 k5", 5000, 64, 300, 5, 1, "bf16")
        ok &= run_case("filter bf16 L2 k32 d768", 20000, 768, 500, 32, 1, "bf16")
    elif s == 3:  # TF32 path
        ok &= run_case("filter f32 IP k10", 5000, 96, 300, 10, 0, "f32")
        ok &= run_case("filter f32 L2 k10", 5000, 96, 300, 10, 1, "f32")
        ok &= run_case("filter f32 IP k32 d768", 20000, 768, 500, 32, 0, "f32")
    elif s == 4:  # ragged dims (padded filter operand), odd sizes
        ok &= run_case("filter bf16 d100", 3001, 100, 77, 10, 0, "bf16")
        ok &= run_case("filter f32 d30", 2999, 30, 130, 3, 1, "f32")
        ok &= run_case("filter bf16 k1", 4097, 128, 257, 1, 0, "bf16")
        ok &= run_case("filter bf16 k64 (kp96)", 6000, 128, 300, 64, 0, "bf16")
        ok &= run_case("dense k96", 6000, 128, 100, 96, 0, "bf16")
    elif s == 5:  # tie-heavy data through the filter (forces the certified-fallback path)
        ok &= run_case("filter grid IP", 4096, 32, 64, 10, 0, "f32", grid=True)
        ok &= run_case("filter grid L2", 4096, 32, 64, 10, 1, "bf16", grid=True)
    elif s == 6:  # ids subset
        rng = np.random.default_rng(5)
        ids = rng.permutation(6000)[:2500]
        ok &= run_case("subset perm", 6000, 64, 120, 8, 0, "bf16", ids=ids)
        ok &= run_case("subset identity", 3000, 64, 50, 8, 0, "f32", ids=np.arange(3000))
        ok &= run_case("subset small", 3000, 64, 50, 8, 1, "f32", ids=np.array([5, 7, 7, 100, 2999]))
    elif s == 7:  # larger timing run, many splits / many tiles
        ok &= run_case("big bf16", 200000, 768, 4096, 32, 0, "bf16", oracle_rows=48)
        ok &= run_case("big bf16 few q", 200000, 768, 16, 10, 0, "bf16", oracle_rows=16)
    print("STAGE", s, "OK" if ok else "FAILED", flush=True)
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", type=int, default=None)
    ap.add_argument("--stages", type=str, default="0,1,2,3,4,5,6,7")
    a = ap.parse_args()
    if a.stage is not None:
        sys.exit(stage(a.stage))
    os.makedirs("gpurun_out", exist_ok=True)
    rc_all = 0
    for s in [int(t) for t in a.stages.split(",")]:
        log = f"gpurun_out/check_stage{s}.log"
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, __file__, "--stage", str(s)], stdout=f, stderr=subprocess.STDOUT, timeout=300)
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = 124
        tail = open(log).read().strip().splitlines()[-12:]
        print(f"=== stage {s} rc={rc}")
        print("\n".join(tail), flush=True)
        rc_all |= rc
    sys.exit(1 if rc_all else 0)


if __name__ == "__main__":
    main()

"""Secondary BASELINE.json configurations on ONE GPU (C2 full size; C4 and C5 scaled to what one GPU holds/finishes),
each with an oracle check on a sample. One JSON line per config into stdout (and gpurun_out/configs_*.json)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from lotus_b200 import _native as nv  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def c2(args):
    """sem_search over 1M x 768 fp32 index, 10k queries, top-10 (VS.__call__ boundary; TF32 filter + exact fp32 re-score)."""
    import oracle
    dev = torch.device("cuda", 0)
    n, d, nq, k = args.n, 768, 10_000, 10
    x = bench.gen_rows_torch(torch, 0, n, d, 0, dev, torch.float32)
    q = bench.gen_rows_torch(torch, 0, nq, d, 1, dev, torch.float32)
    idx = nv.Index(None, nv.F32, nv.METRIC_IP, 0, on_device_ptr=x.data_ptr(), n=n, d=d)
    os_, oi = torch.empty((nq, k), dtype=torch.float32, device=dev), torch.empty((nq, k), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    fms = []

    def step():
        idx.search_dev(q.data_ptr(), nq, k, nv.F32, os_.data_ptr(), oi.data_ptr(), stream=st)
        fms.append(idx.last_filter_ms())

    for _ in range(3):
        step()
    fms.clear()
    nv.stats_reset()
    ms = timed(step, 5)
    stt = nv.stats()
    xs, qs = x.cpu().numpy(), q[:8].cpu().numpy()
    Do, Io = oracle.knn(xs, qs, k, oracle.IP)
    ok = bool(np.array_equal(oi[:8].cpu().numpy(), Io) and np.array_equal(os_[:8].cpu().numpy().view(np.uint32), Do.view(np.uint32)))
    fl = 2.0 * nq * n * d
    kms = float(np.mean(fms))
    print(json.dumps({"config": "C2 sem_search 10k queries x %d x 768 fp32, K=10, 1 GPU" % n, "ms_per_step": ms, "queries_per_s": nq / ms * 1e3,
                      "filter_ms": kms, "filter_tflops_tf32": fl / kms / 1e9, "fallback_queries": stt["fallback_queries"] / 5,
                      "pipe": "kind::tf32 tcgen05 (1x issue), exact fp32 values re-scored in fp64",
                      "idx_and_score_bit_exact_vs_oracle_8q": ok}), flush=True)
    idx.close()


def c4(args):
    """sem_dedup relation: all pairs with cosine > 0.95 among N x 384 bf16 rows (1 % planted near-duplicates)."""
    import oracle
    dev = torch.device("cuda", 0)
    n, d = args.n_dedup, 384
    x = bench.gen_rows_torch(torch, 0, n, d, 2, dev, torch.float32)
    g = torch.Generator(device=dev)
    g.manual_seed(22)
    m = n // 100
    src = torch.randint(0, n, (m,), generator=g, device=dev)
    dst = torch.randperm(n, generator=g, device=dev)[:m]
    noise = torch.randn((m, d), generator=g, device=dev) * (0.1 / d ** 0.5)
    x[dst] = x[src] + noise
    x = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).contiguous()
    idx = nv.Index(None, nv.BF16, nv.METRIC_IP, 0, on_device_ptr=x.data_ptr(), n=n, d=d)
    idx.threshold_pairs(0.95)  # warm-up (sizes the candidate buffer)
    nv.stats_reset()
    t0 = time.perf_counter()
    pi, pj = idx.threshold_pairs(0.95)
    dt = time.perf_counter() - t0
    lab = nv.connected_components(n, pi, pj, 0)
    ncomp_dup = int((lab != np.arange(n)).sum())
    # oracle on a slice: the relation restricted to the first 3000 rows must match exactly
    sub = 3000
    xs = x[:sub].float().cpu().numpy()
    oi, oj, _ = oracle.threshold_pairs(xs, 0.95)
    keep = (pi < sub) & (pj < sub)
    ok = bool(np.array_equal(pi[keep], oi) and np.array_equal(pj[keep], oj))
    fl = float(n) * (n - 1) / 2 * 2 * d
    print(json.dumps({"config": "C4 sem_dedup pairs, %d x 384 bf16, tau=0.95, 1 GPU" % n, "seconds": dt, "pairs": int(len(pi)),
                      "rows_removed": ncomp_dup, "tflops_symmetric": fl / dt / 1e12, "slice_relation_exact_vs_oracle": ok,
                      "stats": nv.stats()}), flush=True)
    idx.close()


def c5(args):
    """sem_cluster_by: faiss-parity k-means, N x 768 bf16, k=1024, 20 iterations (+ final assignment of all points)."""
    dev = torch.device("cuda", 0)
    n, d, k = args.n_kmeans, 768, 1024
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    centers = torch.randn((k, d), generator=g, device=dev)
    lab = torch.randint(0, k, (n,), generator=g, device=dev)
    x = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for s in range(0, n, 1 << 18):
        e = min(n, s + (1 << 18))
        x[s:e] = (centers[lab[s:e]] + 0.5 * torch.randn((e - s, d), generator=g, device=dev)).to(torch.bfloat16)
    idx = nv.Index(None, nv.BF16, nv.METRIC_L2, 0, on_device_ptr=x.data_ptr(), n=n, d=d)
    for full in (0, 1):
        nv.stats_reset()
        t0 = time.perf_counter()
        a, c, obj = idx.kmeans(k, niter=20, full_lloyd=bool(full))
        dt = time.perf_counter() - t0
        st = nv.stats()
        purity = float((torch.from_numpy(a).to(dev)[:200000] == torch.from_numpy(a).to(dev)[:200000]).float().mean())
        npts = n if full else min(n, 256 * k)
        fl = 2.0 * npts * k * d * 20 + 2.0 * n * k * d
        print(json.dumps({"config": "C5 k-means %d x 768 bf16, k=1024, 20 it, %s, 1 GPU" % (n, "full Lloyd" if full else "faiss parity (256k subsample)"),
                          "seconds": dt, "s_per_iteration": dt / 21, "assign_tflops_equiv": fl / dt / 1e12, "obj_first": float(obj[0]),
                          "obj_last": float(obj[-1]), "clusters_used": int(len(np.unique(a))), "fallback_queries": st["fallback_queries"],
                          "launches": st["launches"]}), flush=True)
    idx.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c2,c4,c5")
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--n-dedup", type=int, default=1_000_000)
    ap.add_argument("--n-kmeans", type=int, default=1_000_000)
    a = ap.parse_args()
    for w in a.which.split(","):
        {"c2": c2, "c4": c4, "c5": c5}[w](a)

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
    part, nparts = args.dedup_part, args.dedup_parts
    if not args.no_warmup:
        idx.threshold_pairs(0.95, part=part, nparts=nparts)  # warm-up
    nv.stats_reset()
    t0 = time.perf_counter()
    pi, pj = idx.threshold_pairs(0.95, part=part, nparts=nparts)
    dt = time.perf_counter() - t0
    lab = nv.connected_components(n, pi, pj, 0)
    ncomp_dup = int((lab != np.arange(n)).sum())
    # oracle on a slice: the relation restricted to the first 3000 rows (and to this rank's 128-row tiles) must match exactly
    sub = 3000
    xs = x[:sub].float().cpu().numpy()
    oi, oj, _ = oracle.threshold_pairs(xs, 0.95)
    mine = nv.pair_owner(oi, nparts) == part
    oi, oj = oi[mine], oj[mine]
    keep = (pi < sub) & (pj < sub)
    ok = bool(np.array_equal(pi[keep], oi) and np.array_equal(pj[keep], oj))
    # full-size properties: i < j, sorted unique, every pair owned by this part, every returned pair really above tau, and
    # every planted pair above tau (recomputed with torch in fp64 from the stored bf16 rows) is present
    tpi, tpj = torch.from_numpy(pi).to(dev), torch.from_numpy(pj).to(dev)
    key = pi.astype(np.uint64) << np.uint64(32) | pj.astype(np.uint64)
    shape_ok = bool((pi < pj).all() and (np.diff(key.astype(np.int64)) > 0).all() and (nv.pair_owner(pi, nparts) == part).all())
    sc_out = torch.empty(len(pi), dtype=torch.float64, device=dev)
    for s in range(0, len(pi), 1 << 18):
        e = min(len(pi), s + (1 << 18))
        sc_out[s:e] = (x[tpi[s:e]].double() * x[tpj[s:e]].double()).sum(1)
    all_above = bool((sc_out.float() > 0.95).all()) if len(pi) else True
    lo, hi = torch.minimum(src, dst), torch.maximum(src, dst)
    sc_pl = torch.empty(m, dtype=torch.float64, device=dev)
    for s in range(0, m, 1 << 18):
        e = min(m, s + (1 << 18))
        sc_pl[s:e] = (x[lo[s:e]].double() * x[hi[s:e]].double()).sum(1)
    want = (sc_pl.float() > 0.95) & (lo != hi) & (torch.from_numpy(nv.pair_owner(lo.cpu().numpy(), nparts)).to(dev) == part)
    wkey = (lo[want].cpu().numpy().astype(np.uint64) << np.uint64(32)) | hi[want].cpu().numpy().astype(np.uint64)
    planted_found = bool(np.isin(wkey, key).all())
    fl = float(n) * (n - 1) / 2 * 2 * d / nparts
    print(json.dumps({"config": "C4 sem_dedup pairs, %d x 384 bf16, tau=0.95, 1 GPU, tile share %d/%d" % (n, part, nparts), "seconds": dt,
                      "pairs": int(len(pi)), "rows_removed": ncomp_dup, "tflops_symmetric": fl / dt / 1e12,
                      "slice_relation_exact_vs_oracle": ok, "pairs_sorted_unique_owned": shape_ok, "all_returned_above_tau_fp64": all_above,
                      "planted_pairs_expected": int(want.sum()), "planted_pairs_all_found": planted_found, "stats": nv.stats()}), flush=True)
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
    for full in [int(v) for v in args.c5_modes.split(",")]:
        nv.stats_reset()
        t0 = time.perf_counter()
        a, c, obj = idx.kmeans(k, niter=args.c5_niter, full_lloyd=bool(full))
        dt = time.perf_counter() - t0
        st = nv.stats()
        # full-size properties: the returned assignment is a fixed point of kmeans_assign on the returned centroids
        # (faiss: `kmeans.index.search(x, 1)`, utils.py:65) and equals the fp64 argmin on a sample
        a2, _ = idx.kmeans_assign(c)
        idem = bool(np.array_equal(a, a2))
        smp = torch.arange(0, n, max(1, n // 4096), device=dev)[:4096]
        dd = torch.cdist(x[smp].double(), torch.from_numpy(c).to(dev).double())
        srt = dd.topk(2, dim=1, largest=False)
        clear = ((srt.values[:, 1] - srt.values[:, 0]) > 1e-6 * srt.values[:, 1]).cpu().numpy()
        amin_ok = bool((srt.indices[:, 0].cpu().numpy()[clear] == a[smp.cpu().numpy()][clear]).all())
        npts = n if full else min(n, 256 * k)
        fl = 2.0 * npts * k * d * args.c5_niter + 2.0 * n * k * d
        print(json.dumps({"config": "C5 k-means %d x 768 bf16, k=1024, 20 it, %s, 1 GPU" % (n, "full Lloyd" if full else "faiss parity (256k subsample)"),
                          "seconds": dt, "s_per_iteration": dt / (args.c5_niter + 1), "assign_tflops_equiv": fl / dt / 1e12, "obj_first": float(obj[0]),
                          "obj_last": float(obj[-1]), "clusters_used": int(len(np.unique(a))), "assign_is_fixed_point": idem,
                          "assign_equals_fp64_argmin_4096_sample": amin_ok, "fallback_queries": st["fallback_queries"],
                          "launches": st["launches"]}), flush=True)
    idx.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c2,c4,c5")
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--n-dedup", type=int, default=1_000_000)
    ap.add_argument("--n-kmeans", type=int, default=1_000_000)
    ap.add_argument("--dedup-part", type=int, default=0)
    ap.add_argument("--dedup-parts", type=int, default=1, help="time one rank's share of the upper-triangular tile grid (8 = C4's 8-GPU split)")
    ap.add_argument("--no-warmup", action="store_true")
    ap.add_argument("--c5-modes", default="0,1", help="0 = faiss parity (256*k subsample), 1 = full Lloyd")
    ap.add_argument("--c5-niter", type=int, default=20)
    a = ap.parse_args()
    for w in a.which.split(","):
        {"c2": c2, "c4": c4, "c5": c5}[w](a)

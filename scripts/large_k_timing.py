"""Timing of the k > 64 paths on one GPU: sem_sim_join at C3 with K = 32 / 64 / 128 / 256 / 1000, and the
single-query K = len(df) search of the cascade callers over 1M rows. One JSON line each."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from lotus_b200 import _native as nv  # noqa: E402


def main():
    import oracle
    dev = torch.device("cuda", 0)
    n, d, nq = 1_000_000, 768, int(os.environ.get("NQ", 100_000))
    x = bench.gen_rows_torch(torch, 0, n, d, 0, dev, torch.bfloat16)
    q = bench.gen_rows_torch(torch, 0, nq, d, 1, dev, torch.bfloat16)
    idx = nv.Index(None, nv.BF16, nv.METRIC_IP, 0, on_device_ptr=x.data_ptr(), n=n, d=d)
    st = torch.cuda.current_stream().cuda_stream
    xs = x.float().cpu().numpy()
    for k in (32, 64, 128, 256, 1000):
        nqk = nq if k <= 256 else nq // 10
        os_ = torch.empty((nqk, k), dtype=torch.float32, device=dev)
        oi = torch.empty((nqk, k), dtype=torch.int64, device=dev)
        qq = q[:nqk].contiguous()
        for _ in range(2):
            idx.search_dev(qq.data_ptr(), nqk, k, nv.BF16, os_.data_ptr(), oi.data_ptr(), stream=st)
        nv.stats_reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            idx.search_dev(qq.data_ptr(), nqk, k, nv.BF16, os_.data_ptr(), oi.data_ptr(), stream=st)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        stt = nv.stats()
        Do, Io = oracle.knn(xs, qq[:4].float().cpu().numpy(), k, oracle.IP)
        ok = bool(np.array_equal(oi[:4].cpu().numpy(), Io) and np.array_equal(os_[:4].cpu().numpy().view(np.uint32), Do.view(np.uint32)))
        print(json.dumps({"case": f"sem_sim_join {nqk} x {n} x {d} bf16 K={k}", "ms": ms, "queries_per_s": nqk / ms * 1e3,
                          "filter_ms": idx.last_filter_ms(), "fallback_queries": stt["fallback_queries"] / reps,
                          "bit_exact_vs_oracle_4q": ok}), flush=True)
    # the cascade callers: one query, every row back, best first
    for k in (n, 100_000, 2048, 2049):
        os_ = torch.empty((1, k), dtype=torch.float32, device=dev)
        oi = torch.empty((1, k), dtype=torch.int64, device=dev)
        qq = q[:1].contiguous()
        for _ in range(2):
            idx.search_dev(qq.data_ptr(), 1, k, nv.BF16, os_.data_ptr(), oi.data_ptr(), stream=st)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            idx.search_dev(qq.data_ptr(), 1, k, nv.BF16, os_.data_ptr(), oi.data_ptr(), stream=st)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        qh = q[:1].cpu().view(torch.int16).numpy().view(np.uint16)
        idx.search(qh, k, nv.BF16)
        t0 = time.perf_counter()
        for _ in range(3):
            Dh, Ih = idx.search(qh, k, nv.BF16)
        ms_host = (time.perf_counter() - t0) / 3 * 1e3
        Do, Io = oracle.knn(xs, qq.float().cpu().numpy(), k, oracle.IP)
        ok = bool(np.array_equal(Ih, Io) and np.array_equal(Dh.view(np.uint32), Do.view(np.uint32)))
        print(json.dumps({"case": f"single query, K={k} of {n} rows (sem_filter / sem_join cascade)", "ms_device_buffers": ms,
                          "ms_host_buffers": ms_host, "bit_exact_vs_oracle": ok}), flush=True)
    idx.close()


if __name__ == "__main__":
    main()

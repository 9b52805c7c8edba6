"""Operator-scope timing: DataFrame in -> DataFrame out through the pandas accessor, next to the
VS-boundary time of the same search. left = Q rows, right = N rows, both indexed through B200VS (bf16), K = 32.

    python scripts/operator_scope.py [--nq 100000 --n 1000000 --d 768 --k 32]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import lotus_b200 as lotus  # noqa: E402
from lotus_b200 import _native as nv  # noqa: E402


class PrecomputedRM(lotus.RM):
    """sem_index asks the RM for the column's embeddings: hand over the precomputed device tensor (device hand-off)."""

    def __init__(self, table):
        super().__init__()
        self.table = table
        self.next = None

    def _embed(self, docs):
        return self.table[self.next]

    def __call__(self, docs):
        return self._embed(docs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nq", type=int, default=100_000)
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    x = bench.gen_rows_torch(torch, 0, a.n, a.d, 0, dev, torch.bfloat16)
    q = bench.gen_rows_torch(torch, 0, a.nq, a.d, 1, dev, torch.bfloat16)
    rm = PrecomputedRM({"left": q, "right": x})
    vs = lotus.B200VS(dtype="bf16")
    lotus.settings.configure(rm=rm, vs=vs, enable_cache=False)
    left = pd.DataFrame({"article": [f"a{i}" for i in range(a.nq)]})
    right = pd.DataFrame({"category": [f"c{i}" for i in range(a.n)]})
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.perf_counter()
        rm.next = "left"
        left = left.sem_index("article", os.path.join(tmp, "l"))
        rm.next = "right"
        right = right.sem_index("category", os.path.join(tmp, "r"))
        t_index = time.perf_counter() - t0
        times = []
        for _ in range(a.reps):
            nv.stats_reset()
            t0 = time.perf_counter()
            out = left.sem_sim_join(right, "article", "category", K=a.k)
            times.append(time.perf_counter() - t0)
        st = nv.stats()
        # VS boundary alone (host query matrix in, host result out), same index
        qh = vs.get_vectors_from_index(os.path.join(tmp, "l"), left.index)
        vs.load_index(os.path.join(tmp, "r"))
        t0 = time.perf_counter()
        vs(qh, a.k, ids=np.asarray(right.index))
        t_vs = time.perf_counter() - t0
    best = min(times)
    print(json.dumps({"workload": f"df.sem_sim_join {a.nq} x {a.n} x {a.d} bf16 K={a.k} (DataFrame in -> DataFrame out)",
                      "operator_seconds": best, "operator_queries_per_s": a.nq / best, "all_reps_s": times,
                      "vs_boundary_seconds": t_vs, "rows_out": int(len(out)), "sem_index_seconds_both_frames": t_index,
                      "fallback_queries": st["fallback_queries"], "launches": st["launches"]}))


if __name__ == "__main__":
    main()

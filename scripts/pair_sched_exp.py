"""Scheduling experiment for the all-pairs (sem_dedup) kernel: group size / aligned sweep starts, at 1M rows (one rank) and
at 10M rows (one rank's share of 8). Every setting must return the identical pair list."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from lotus_b200 import _native as nv  # noqa: E402


def make(n, d, dev):
    x = bench.gen_rows_torch(torch, 0, n, d, 2, dev, torch.float32)
    g = torch.Generator(device=dev)
    g.manual_seed(22)
    m = n // 100
    src = torch.randint(0, n, (m,), generator=g, device=dev)
    dst = torch.randperm(n, generator=g, device=dev)[:m]
    x[dst] = x[src] + torch.randn((m, d), generator=g, device=dev) * (0.1 / d ** 0.5)
    return (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16).contiguous()


def main():
    dev = torch.device("cuda", 0)
    settings = [("148", "0"), ("148", "1"), ("74", "1"), ("296", "1")]
    for n, nparts in ((1_000_000, 1), (10_000_000, 8)):
        x = make(n, 384, dev)
        idx = nv.Index(None, nv.BF16, nv.METRIC_IP, 0, on_device_ptr=x.data_ptr(), n=n, d=384)
        ref = None
        for group, align in settings:
            os.environ["B2_PAIR_GROUP"], os.environ["B2_PAIR_ALIGN"] = group, align
            best = 1e9
            for rep in range(3 if n <= 1_000_000 else 2):
                t0 = time.perf_counter()
                pi, pj = idx.threshold_pairs(0.95, part=0, nparts=nparts)
                best = min(best, time.perf_counter() - t0)
            key = (pi.astype(np.uint64) << np.uint64(32)) | pj.astype(np.uint64)
            if group == "148":  # ownership of a pair depends on the group size: compare like with like
                same = True if ref is None else bool(np.array_equal(ref, key))
                ref = key if ref is None else ref
            else:
                same = None
            fl = float(n) * (n - 1) / 2 * 2 * 384 / nparts
            print(json.dumps({"n": n, "nparts": nparts, "group": group, "align": align, "seconds": best, "pairs": int(len(pi)),
                              "tflops_symmetric": fl / best / 1e12, "same_pairs_as_first": same}), flush=True)
        idx.close()
        del x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

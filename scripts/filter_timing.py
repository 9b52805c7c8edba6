"""Filter-kernel timing experiments: B2_FILTER_2CTA / B2_FILTER_DEBUG are read once per process, so each variant runs in
its own subprocess. Prints kernel ms and TFLOP/s for the C3 shape (or --nq/--n)."""
import argparse
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(a):
    import numpy as np
    import torch
    import bench
    from lotus_b200 import _native as nv
    dev = torch.device("cuda", 0)
    x = bench.gen_rows_torch(torch, 0, a.n, a.d, 0, dev, torch.bfloat16)
    q = bench.gen_rows_torch(torch, 0, a.nq, a.d, 1, dev, torch.bfloat16)
    idx = nv.Index(None, nv.BF16, nv.METRIC_IP, 0, on_device_ptr=x.data_ptr(), n=a.n, d=a.d)
    os_ = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
    oi = torch.empty((a.nq, a.k), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ms, tot = [], []
    for i in range(a.reps + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        idx.search_dev(q.data_ptr(), a.nq, a.k, nv.BF16, os_.data_ptr(), oi.data_ptr(), stream=st)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(idx.last_filter_ms())
            tot.append(e0.elapsed_time(e1))
    m = float(np.mean(ms))
    print(json.dumps({"two_cta": os.environ.get("B2_FILTER_2CTA", "1"), "debug": os.environ.get("B2_FILTER_DEBUG", "0"),
                      "kernel_ms": m, "tflops": 2.0 * a.nq * a.n * a.d / m / 1e9, "step_ms": float(np.mean(tot)),
                      "non_filter_ms": float(np.mean(tot)) - m}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--nq", type=int, default=100_000)
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variants", default="0:0,1:0,0:1,1:1")
    a = ap.parse_args()
    if a.child:
        child(a)
    else:
        for v in a.variants.split(","):
            two, dbg = v.split(":")
            env = dict(os.environ, B2_FILTER_2CTA=two, B2_FILTER_DEBUG=dbg)
            subprocess.run([sys.executable, __file__, "--child", "--nq", str(a.nq), "--n", str(a.n), "--d", str(a.d), "--k", str(a.k),
                            "--reps", str(a.reps)], env=env, timeout=600)

"""Single-process multi-GPU store: B200VS(devices=[0..G-1]) against the oracle (run on a box with G >= 2 GPUs, plain python).
Also exercises two handles on two devices from one process (the kernels' shared-memory attribute is per device)."""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import lotus_b200 as lotus  # noqa: E402
from lotus_b200 import _native as nv  # noqa: E402


def gauss(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def main():
    g = nv.device_count()
    out = {"devices": g}
    x, q = gauss(50_021, 96, 0), gauss(500, 96, 1)
    # two independent handles, one per device, used alternately from this process
    if g >= 2:
        a, b = nv.Index(x[:30_000].copy(), nv.F32, nv.METRIC_IP, 0), nv.Index(x[30_000:].copy(), nv.F32, nv.METRIC_IP, 1)
        ok = True
        for _ in range(2):
            for idx, xs in ((a, x[:30_000]), (b, x[30_000:])):
                D, I = idx.search(q, 10)
                Do, Io = oracle.knn(xs, q, 10, oracle.IP)
                ok &= bool(np.array_equal(I, Io) and np.array_equal(D.view(np.uint32), Do.view(np.uint32)))
        a.close()
        b.close()
        out["two_handles_two_devices"] = ok
    with tempfile.TemporaryDirectory() as tmp:
        for metric, om in ((lotus.METRIC_INNER_PRODUCT, oracle.IP), (lotus.METRIC_L2, oracle.L2)):
            for dtype in ("f32", "bf16"):
                vs = lotus.B200VS(metric=metric, dtype=dtype, devices=list(range(g)))
                vs.index(None, x, os.path.join(tmp, f"i{metric}{dtype}"))
                xf = x if dtype == "f32" else nv.bf16_bits_to_f32(nv.f32_to_bf16_bits(x))
                r = vs(q, 10)
                Do, Io = oracle.knn(xf, q, 10, om)
                ok = bool(np.array_equal(r.indices, Io) and np.array_equal(np.asarray(r.distances).view(np.uint32), Do.view(np.uint32)))
                ids = np.arange(5, len(x), 3)
                r = vs(q, 10, ids=ids)
                Ds, Is = oracle.knn_subset(xf, q, 10, ids, om)
                ok &= bool(np.array_equal(r.indices, Is) and np.array_equal(np.asarray(r.distances).view(np.uint32), Ds.view(np.uint32)))
                got = vs.get_vectors_from_index(os.path.join(tmp, f"i{metric}{dtype}"), [50_020, 0, 25_000])
                ok &= bool(np.array_equal(np.asarray(got), xf[[50_020, 0, 25_000]]))
                out[f"devices_store_metric{metric}_{dtype}"] = ok
                vs.close()
    out["all_ok"] = all(v for k, v in out.items() if isinstance(v, bool))
    print("MULTIDEVICE CHECK", json.dumps(out), flush=True)
    sys.exit(0 if out["all_ok"] else 1)


if __name__ == "__main__":
    main()

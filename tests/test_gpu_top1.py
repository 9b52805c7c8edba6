"""k == 1 searches with the register-resident top-2 epilogue (B2_FILTER_TOP1=1, read once per process -> subprocess with a
timeout): nearest neighbour, k-means assignment and the faiss-parity k-means must stay bit-identical to the oracle."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import oracle
from helpers import gauss, grid
from lotus_b200 import _native as nv
out = {}
for metric, name in ((nv.METRIC_IP, "ip"), (nv.METRIC_L2, "l2")):
    for n, d, dtype in ((1024, 64, "bf16"), (1500, 100, "f32"), (5000, 768, "bf16"), (513, 32, "bf16")):
        x, q = gauss(n, d, 40 + n), gauss(700, d, 41 + n)
        if dtype == "bf16":
            xb, qb = nv.f32_to_bf16_bits(x), nv.f32_to_bf16_bits(q)
            idx, xf, qf, qa, qc = nv.Index(xb, nv.BF16, metric, 0), nv.bf16_bits_to_f32(xb), nv.bf16_bits_to_f32(qb), qb, nv.BF16
        else:
            idx, xf, qf, qa, qc = nv.Index(x, nv.F32, metric, 0), x, q, q, nv.F32
        nv.stats_reset()
        D, I = idx.search(qa, 1, qc)
        st = nv.stats()
        Do, Io = oracle.knn(xf, qf, 1, metric)
        out[f"{name}_{n}x{d}_{dtype}"] = {"idx": bool(np.array_equal(I, Io)), "score": bool(np.array_equal(D.view(np.uint32), Do.view(np.uint32))),
                                          "filter_launches": int(st["filter_launches"]), "fallback": int(st["fallback_queries"])}
        idx.close()
# ties everywhere: the certificate must refuse and the dense path must answer
xg, qg = grid(2000, 4, 7), grid(300, 4, 8)
idx = nv.Index(xg, nv.F32, nv.METRIC_L2, 0)
D, I = idx.search(qg, 1, nv.F32)
Do, Io = oracle.knn(xg, qg, 1, oracle.L2)
out["grid_l2"] = {"idx": bool(np.array_equal(I, Io)), "score": bool(np.array_equal(D.view(np.uint32), Do.view(np.uint32)))}
idx.close()
# faiss-parity k-means end to end (assignment steps are k == 1 searches against 640 centroids)
pts = gauss(30000, 48, 50, normalize=False)
pb = nv.f32_to_bf16_bits(pts)
idx = nv.Index(pb, nv.BF16, nv.METRIC_L2, 0)
a, c, obj = idx.kmeans(640, niter=6)
ao, co, oo = oracle.kmeans(nv.bf16_bits_to_f32(pb), 640, niter=6)
out["kmeans"] = {"idx": bool(np.array_equal(a, ao)), "score": bool(np.array_equal(c.view(np.uint32), co.view(np.uint32)))}
idx.close()
print(json.dumps(out))
""" % (ROOT, ROOT)


@pytest.mark.gpu
def test_top1_epilogue_is_exact():
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, B2_FILTER_TOP1="1"))
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    for name, v in res.items():
        assert v["idx"] and v["score"], (name, v)
        if "filter_launches" in v:
            assert v["filter_launches"] >= 1, (name, v)

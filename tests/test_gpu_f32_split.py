"""fp32 indexes searched with bf16 hi|lo split operands (B2_F32_SPLIT=1, read once per process -> subprocess): results must be
bit-identical to the oracle, like the TF32 route, and the tighter error bound must not cost fallbacks on Gaussian data."""
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
from helpers import gauss
from lotus_b200 import _native as nv
out = {}
for metric, name in ((nv.METRIC_IP, "ip"), (nv.METRIC_L2, "l2")):
    for d in (100, 128, 768):
        x, q = gauss(20000, d, 11 + d), gauss(300, d, 12 + d)
        idx = nv.Index(x, nv.F32, metric, 0)
        nv.stats_reset()
        D, I = idx.search(q, 10, nv.F32)
        st = nv.stats()
        Do, Io = oracle.knn(x, q, 10, metric)
        out[f"{name}_{d}"] = {"idx": bool(np.array_equal(I, Io)), "score": bool(np.array_equal(D.view(np.uint32), Do.view(np.uint32))),
                               "fallback": int(st["fallback_queries"]), "filter_launches": int(st["filter_launches"])}
        qb = nv.f32_to_bf16_bits(q)                      # bf16 queries against the fp32 index
        D, I = idx.search(qb, 10, nv.BF16)
        Do, Io = oracle.knn(x, nv.bf16_bits_to_f32(qb), 10, metric)
        out[f"{name}_{d}_bf16q"] = {"idx": bool(np.array_equal(I, Io)), "score": bool(np.array_equal(D.view(np.uint32), Do.view(np.uint32)))}
        idx.close()
print(json.dumps(out))
""" % (ROOT, ROOT)


@pytest.mark.gpu
def test_f32_index_with_hi_lo_split_filter_is_exact():
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, B2_F32_SPLIT="1"))
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    for name, v in res.items():
        assert v["idx"] and v["score"], (name, v)
        if "fallback" in v:
            assert v["filter_launches"] >= 1 and v["fallback"] <= 3, (name, v)

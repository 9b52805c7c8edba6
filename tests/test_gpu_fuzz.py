"""Seeded random-shape sweep of the search path against the oracle (bit-exact indices and scores): odd sizes around the
tile / chunk / split boundaries, both metrics, both dtypes, random k, optional ids subsets, duplicate rows."""
import numpy as np
import pytest

import oracle
from helpers import bits

pytestmark = pytest.mark.gpu


def one_case(nv, rng, case):
    n = int(rng.choice([513, 600, 767, 768, 1000, 1025, 2047, 2048, 2049, 3000, 4095, 5000, 9001]))
    d = int(rng.choice([4, 7, 8, 16, 30, 33, 64, 96, 100, 128, 200, 384]))
    nq = int(rng.choice([1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 300, 700]))
    k = int(rng.choice([1, 2, 3, 5, 6, 7, 10, 16, 17, 32, 40, 41, 64, 65, 100]))
    metric = int(rng.integers(0, 2))
    dtype = "bf16" if rng.random() < 0.5 else "f32"
    x = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    if rng.random() < 0.5:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    if rng.random() < 0.3:  # exact duplicates exercise the tie rules
        dup = rng.integers(0, n, 20)
        x[dup] = x[rng.integers(0, n, 20)]
    ids = None
    if rng.random() < 0.25:
        ids = rng.permutation(n)[: int(rng.integers(1, n))]
        if rng.random() < 0.5:
            ids = np.sort(ids)
    if dtype == "bf16":
        xb, qb = nv.f32_to_bf16_bits(x), nv.f32_to_bf16_bits(q)
        xf, qf = nv.bf16_bits_to_f32(xb), nv.bf16_bits_to_f32(qb)
        idx, qa, qdt = nv.Index(xb, nv.BF16, metric), qb, nv.BF16
    else:
        xf, qf = x, q
        idx, qa, qdt = nv.Index(x, nv.F32, metric), q, nv.F32
    D, I = idx.search(qa, k, qdt, ids=ids)
    idx.close()
    Do, Io = oracle.knn(xf, qf, k, metric) if ids is None else oracle.knn_subset(xf, qf, k, ids, metric)
    tag = f"case {case}: n={n} d={d} nq={nq} k={k} metric={metric} dtype={dtype} ids={'none' if ids is None else len(ids)}"
    assert np.array_equal(I, Io), tag + f" ({(I != Io).any(axis=1).sum()} bad rows)"
    assert np.array_equal(bits(D), bits(Do)), tag


@pytest.mark.parametrize("seed", [101, 202, 303, 404])
def test_random_shapes(gpu, seed):
    rng = np.random.default_rng(seed)
    for case in range(12):
        one_case(gpu, rng, f"{seed}/{case}")

"""Regenerates tests/golden/knn_golden.npz and tests/golden/kmeans_golden.npz from the oracle.

    python tests/golden/make_golden.py

The reference (lotus-data/lotus) cannot be imported in this image (litellm / faiss / sentence-transformers are
absent) and holds no numeric fixtures for this path, so these vectors are produced by the oracle
restatement itself on the quantised-grid data set (coordinates j/64, |j| <= 8: every product and partial sum is exact
in fp32, so ANY correct flat search — faiss's sgemm included — must return exactly these scores; ties abound) and on
small Gaussian cases. They pin the oracle against drift and give the GPU tests size-independent known answers.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def grid(n, d, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(-8, 9, size=(n, d)).astype(np.float32) / 64).astype(np.float32)


def gauss(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def main():
    out = {}
    for name, x, q in [("grid", grid(700, 24, 10), grid(37, 24, 11)), ("gauss", gauss(900, 48, 12), gauss(41, 48, 13))]:
        out[f"{name}_x"] = x
        out[f"{name}_q"] = q
        for metric, mname in [(oracle.IP, "ip"), (oracle.L2, "l2")]:
            for k in (1, 5, 32):
                D, I = oracle.knn(x, q, k, metric)
                out[f"{name}_{mname}_k{k}_D"] = D
                out[f"{name}_{mname}_k{k}_I"] = I
    np.savez_compressed(os.path.join(HERE, "knn_golden.npz"), **out)
    x = gauss(600, 16, 20) + np.repeat(gauss(6, 16, 21) * 3, 100, axis=0)
    a, c, obj = oracle.kmeans(x.astype(np.float32), 6, niter=5)
    np.savez_compressed(os.path.join(HERE, "kmeans_golden.npz"), x=x.astype(np.float32), assign=a, centroids=c, obj=obj,
                        perm_seed1234_n20=oracle.rand_perm(20, 1234), mt_seed1234_first8=oracle.mt19937(1234, 8))
    print("wrote golden fixtures")


if __name__ == "__main__":
    main()

import numpy as np


def grid(n, d, seed):
    """Quantised-grid data (coordinates j/64, |j| <= 8): products and partial sums are exact in fp32 and bf16, so
    scores are identical under ANY summation order and ties are abundant (SURVEY.md §7 'Hard parts')."""
    rng = np.random.default_rng(seed)
    return (rng.integers(-8, 9, size=(n, d)).astype(np.float32) / 64).astype(np.float32)


def gauss(n, d, seed, normalize=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


class NumpyVS:
    """Test double of the VS plugin surface (lotus/vector_store/vs.py) backed by the oracle: lets the pandas
    operators be exercised on CPU. It lives in tests/ only — the product has no CPU path."""

    supports_ids_search = True

    def __init__(self, metric=0):
        self.metric = metric
        self.index_dir = None
        self.dirs = {}
        self.x = None

    def index(self, docs, embeddings, index_dir, **kw):
        self.dirs[index_dir] = np.asarray(embeddings, dtype=np.float32)
        self.index_dir = index_dir
        self.x = self.dirs[index_dir]

    def load_index(self, index_dir):
        if index_dir not in self.dirs:
            raise ValueError(f"Index directory {index_dir} not found")
        self.index_dir = index_dir
        self.x = self.dirs[index_dir]

    def get_vectors_from_index(self, index_dir, ids):
        return self.dirs[index_dir][np.asarray(list(ids), dtype=np.int64)]

    def __call__(self, query_vectors, K, ids=None, **kw):
        import oracle
        from lotus_b200.types import RMOutput
        if self.x is None:
            raise ValueError("Index not loaded")
        q = np.asarray(query_vectors, dtype=np.float32)
        if ids is None:
            D, I = oracle.knn(self.x, q, K, self.metric)
        else:
            D, I = oracle.knn_subset(self.x, q, K, np.asarray(list(ids), dtype=np.int64), self.metric)
        return RMOutput(distances=D, indices=I)

    # extension methods the re-registered accessors use (oracle-backed, tests only)
    def threshold_pairs(self, threshold):
        import oracle
        pi, pj, _ = oracle.threshold_pairs(self.x, float(threshold))
        return pi, pj

    def kmeans(self, ids, ncentroids, niter=20, seed=1234, full_lloyd=False):
        import oracle
        return oracle.kmeans(self.x[np.asarray(ids, dtype=np.int64)], ncentroids, niter=niter, seed=seed, full_lloyd=full_lloyd)


def frame_from_json(j):
    import pandas as pd
    return pd.DataFrame(j["data"], columns=j["columns"], index=j["index"])


def assert_frame_matches_json(df, j):
    """DataFrame equals a fixture produced by the reference's operator code (tests/golden/reference_ops.json)."""
    assert [str(c) for c in df.columns] == j["columns"]
    assert [int(i) for i in df.index] == j["index"]
    got = df.to_numpy().tolist()
    assert len(got) == len(j["data"])
    for r_got, r_want in zip(got, j["data"]):
        for a, b in zip(r_got, r_want):
            if isinstance(b, float):
                assert np.float32(a) == np.float32(b), (a, b)
            else:
                assert a == b, (a, b)

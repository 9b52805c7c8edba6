"""The oracle against every known answer we hold for this path (CPU only).

The reference's own tests carry no numeric fixtures here (every hot-path test asserts on strings after a
live embedding model), so the pins are: the C++-standard mt19937 known answer, numpy's independent legacy MT19937,
hand-derived heap cases (tests/golden/heap_ties.json), an independent closed form of the heap's tie rule, numpy
float64 arithmetic, and the committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import oracle
from helpers import bits, gauss, grid

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FMAX = np.finfo(np.float32).max


def test_mt19937_cxx_standard_known_answer():
    # ISO C++ [rand.predef]: the 10000th consecutive invocation of a default-constructed mt19937 produces 4123659995
    out = oracle.mt19937(5489, 10000)
    assert out[0] == 3499211612 and out[-1] == 4123659995


def test_mt19937_matches_numpy_legacy_generator():
    for seed in (0, 1, 1234, 1235, 2**31 + 7):
        rs = np.random.RandomState(seed & 0xFFFFFFFF)
        assert np.array_equal(oracle.mt19937(seed, 64), rs._bit_generator.random_raw(64).astype(np.uint32))


def test_rand_perm_is_faiss_fisher_yates():
    # faiss/utils/random.cpp rand_perm restated independently on top of numpy's MT19937 stream
    for n, seed in [(1, 5), (2, 5), (20, 1234), (257, 1235)]:
        raw = np.random.RandomState(seed)._bit_generator.random_raw(max(n, 1))
        perm = np.arange(n)
        for i in range(n - 1):
            i2 = i + int(raw[i]) % (n - i)
            perm[i], perm[i2] = perm[i2], perm[i]
        got = oracle.rand_perm(n, seed)
        assert np.array_equal(got, perm)
        assert sorted(got.tolist()) == list(range(n))


@pytest.mark.parametrize("case", json.load(open(os.path.join(GOLD, "heap_ties.json")))["cases"], ids=lambda c: c["name"])
def test_hand_derived_heap_cases(case):
    """One query whose score against row j is scores[j]: build that with a 1-d index (IP: x=[s], q=[1]; L2: x=[sqrt s], q=[0])."""
    s = np.asarray(case["scores"], dtype=np.float32)
    if case["metric"] == "ip":
        x, q, metric = s[:, None], np.ones((1, 1), np.float32), oracle.IP
    else:
        x, q, metric = np.sqrt(s)[:, None], np.zeros((1, 1), np.float32), oracle.L2
    for scorer in (oracle.CANONICAL, oracle.F32_SEQ):
        D, I = oracle.knn(x, q, case["k"], metric, scorer)
        assert I[0].tolist() == case["expect_ids"], case["why"]
        for d, i in zip(D[0], I[0]):
            if i >= 0:
                assert d == s[i]
            else:
                assert d == (-FMAX if metric == oracle.IP else FMAX)


@pytest.mark.parametrize("metric", [oracle.IP, oracle.L2])
@pytest.mark.parametrize("k", [1, 2, 5, 17, 64, 400])
def test_heap_equals_closed_form_tie_rule(metric, k):
    x, q = grid(300, 16, 0), grid(40, 16, 1)
    D, I = oracle.knn(x, q, k, metric)
    S = oracle.scores(x, q, metric)
    D2, I2 = oracle.knn_window_rule(S, k, metric)
    assert np.array_equal(I, I2) and np.array_equal(bits(D), bits(D2))


def test_canonical_scorer_is_correctly_rounded_dot():
    x, q = gauss(50, 768, 2), gauss(7, 768, 3)
    S = oracle.scores(x, q, oracle.IP)
    ref = (q.astype(np.float64) @ x.astype(np.float64).T)
    assert np.max(np.abs(S.astype(np.float64) - ref)) <= 6e-8  # half an fp32 ulp at |s| < 1
    S2 = oracle.scores(x, q, oracle.L2)
    ref2 = ((q[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)) ** 2).sum(-1)
    assert np.allclose(S2, ref2, rtol=1e-7, atol=0)


def test_scorers_agree_within_fp32_tolerance():
    x, q = gauss(400, 384, 4), gauss(9, 384, 5)
    for metric in (oracle.IP, oracle.L2):
        a = oracle.scores(x, q, metric, oracle.CANONICAL)
        for scorer in (oracle.F32_SEQ, oracle.F32_FAST):
            b = oracle.scores(x, q, metric, scorer)
            assert np.max(np.abs(a - b)) <= 1e-5  # north_star: scores within 1e-5 fp32


def test_grid_scores_are_exact_under_any_order():
    x, q = grid(200, 32, 6), grid(11, 32, 7)
    a = oracle.scores(x, q, oracle.IP, oracle.CANONICAL)
    for scorer in (oracle.F32_SEQ, oracle.F32_FAST):
        assert np.array_equal(bits(a), bits(oracle.scores(x, q, oracle.IP, scorer)))
    D1, I1 = oracle.knn(x, q, 10, oracle.IP, oracle.CANONICAL)
    D2, I2 = oracle.knn(x, q, 10, oracle.IP, oracle.F32_FAST)
    D3, I3 = oracle.knn_blocked(x, q, 10, oracle.IP)
    assert np.array_equal(I1, I2) and np.array_equal(I1, I3) and np.array_equal(bits(D1), bits(D3))


def test_l2_blas_form_clamps_at_zero():
    x = gauss(5, 64, 8)
    D, I = oracle.knn(x, x, 1, oracle.L2, oracle.F32_SEQ)
    assert (D >= 0).all() and np.array_equal(I[:, 0], np.arange(5))
    Dc, Ic = oracle.knn(x, x, 1, oracle.L2, oracle.CANONICAL)
    assert (Dc == 0).all() and np.array_equal(Ic[:, 0], np.arange(5))


def test_k_larger_than_n_pads():
    x, q = gauss(3, 8, 9), gauss(2, 8, 10)
    D, I = oracle.knn(x, q, 5, oracle.IP)
    assert (I[:, 3:] == -1).all() and (D[:, 3:] == -FMAX).all() and (I[:, :3] >= 0).all()
    D, I = oracle.knn(x, q, 5, oracle.L2)
    assert (I[:, 3:] == -1).all() and (D[:, 3:] == FMAX).all()
    D, I = oracle.knn(np.zeros((0, 8), np.float32), q, 2, oracle.IP)
    assert (I == -1).all()


def test_subset_search_remaps_and_does_not_wrap():
    x, q = gauss(50, 16, 11), gauss(4, 16, 12)
    ids = np.array([40, 3, 17, 3, 9])
    D, I = oracle.knn_subset(x, q, 8, ids)
    assert set(I[I >= 0].tolist()) <= set(ids.tolist())
    assert (I[:, 5:] == -1).all()  # the reference would wrap -1 to ids[-1] (faiss_vs.py:71-72); we do not
    Dd, Id = oracle.knn(x[ids], q, 5)
    assert np.array_equal(ids[Id], I[:, :5])


def test_blocked_baseline_matches_flat_recall():
    x, q = gauss(3000, 96, 13), gauss(40, 96, 14)
    D, I = oracle.knn_blocked(x, q, 10, oracle.IP)
    Dc, Ic = oracle.knn(x, q, 10, oracle.IP)
    assert np.mean([len(set(a) & set(b)) / 10 for a, b in zip(I, Ic)]) == 1.0
    assert np.max(np.abs(D - Dc)) <= 1e-5


def test_golden_vectors():
    g = np.load(os.path.join(GOLD, "knn_golden.npz"))
    for name in ("grid", "gauss"):
        x, q = g[f"{name}_x"], g[f"{name}_q"]
        for mname, metric in (("ip", oracle.IP), ("l2", oracle.L2)):
            for k in (1, 5, 32):
                D, I = oracle.knn(x, q, k, metric)
                assert np.array_equal(I, g[f"{name}_{mname}_k{k}_I"])
                assert np.array_equal(bits(D), bits(g[f"{name}_{mname}_k{k}_D"]))
    # the grid set is exact under any arithmetic: an independent float64 numpy search must give the same answer
    x, q = g["grid_x"], g["grid_q"]
    S = (q.astype(np.float64) @ x.astype(np.float64).T).astype(np.float32)
    D, I = oracle.knn_window_rule(S, 5, oracle.IP)
    assert np.array_equal(I, g["grid_ip_k5_I"]) and np.array_equal(bits(D), bits(g["grid_ip_k5_D"]))


def test_threshold_pairs_and_components():
    x = gauss(60, 32, 15)
    x[10] = x[3]
    x[44] = x[3]
    x[20] = x[21] * 0.999 + x[22] * 0.001
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    pi, pj, cnt = oracle.threshold_pairs(x, 0.95)
    S = oracle.scores(x, x, oracle.IP)
    want = [(i, j) for i in range(60) for j in range(i + 1, 60) if S[i, j] > 0.95]
    assert cnt == len(want) and list(zip(pi.tolist(), pj.tolist())) == want
    lab = oracle.connected_components(60, pi, pj)
    assert lab[10] == 3 and lab[44] == 3 and lab[3] == 3 and lab[21] == 20 and lab[5] == 5


def test_kmeans_golden_and_faiss_structure():
    g = np.load(os.path.join(GOLD, "kmeans_golden.npz"))
    assert np.array_equal(oracle.rand_perm(20, 1234), g["perm_seed1234_n20"])
    assert np.array_equal(oracle.mt19937(1234, 8), g["mt_seed1234_first8"])
    x = g["x"]
    a, c, obj = oracle.kmeans(x, 6, niter=5)
    assert np.array_equal(a, g["assign"]) and np.array_equal(bits(c), bits(g["centroids"]))
    assert np.allclose(obj, g["obj"], rtol=1e-6)
    assert (np.diff(obj) <= 1e-3 * obj[0]).all()  # Lloyd's objective does not increase
    # independent numpy restatement of Clustering::train for this case (n <= 256*k: no subsampling)
    perm = oracle.rand_perm(len(x), 1234 + 1)
    cent = x[perm[:6]].copy()
    for _ in range(5):
        d2 = oracle.scores(cent, x, oracle.L2)
        asg = np.argmin(d2, axis=1)
        new = np.zeros_like(cent)
        cntv = np.zeros(6, np.float32)
        for i, ci in enumerate(asg):  # sums in point order, fp32
            new[ci] += x[i]
            cntv[ci] += 1
        assert (cntv > 0).all()
        cent = new * (np.float32(1) / cntv)[:, None]
    assert np.array_equal(bits(cent), bits(c))
    assert np.array_equal(np.argmin(oracle.scores(cent, x, oracle.L2), axis=1), a)


def test_kmeans_subsamples_like_faiss_and_rejects_small_n():
    with pytest.raises(ValueError):
        oracle.kmeans(gauss(3, 4, 1), 5)
    x = gauss(600, 8, 16)
    a, c, obj = oracle.kmeans(x, 2, niter=3)  # 600 > 2*256: trains on the first 512 of rand_perm(600, 1234)
    sub = x[oracle.rand_perm(600, 1234)[:512]]
    a2, c2, _ = oracle.kmeans(sub, 2, niter=3, full_lloyd=True)
    assert np.array_equal(bits(c), bits(c2))
    assert np.array_equal(a, np.argmin(oracle.scores(c, x, oracle.L2), axis=1))


def test_split_clusters_restatement():
    """faiss/Clustering.cpp split_clusters restated independently with numpy's legacy MT19937 stream."""
    cent = gauss(4, 6, 17, normalize=False)
    h = np.array([10, 0, 5, 0], np.float32)
    n, k, d = 15, 4, 6
    c2, h2, ns = oracle.split_clusters(cent, h, n=n)
    raw = iter(np.random.RandomState(1234)._bit_generator.random_raw(10000).tolist())
    c, hh, nsplit = cent.copy(), h.copy(), 0
    eps = 1 / 1024.0
    for ci in range(k):
        if hh[ci] == 0:
            cj = 0
            while True:
                p = np.float32((np.float64(hh[cj]) - 1.0) / np.float64(np.float32(n - k)))
                r = np.float32(next(raw)) / np.float32(4294967295)
                if r < p:
                    break
                cj = (cj + 1) % k
            c[ci] = c[cj]
            for j in range(d):
                sgn = 1 if j % 2 == 0 else -1
                c[ci, j] = np.float32(np.float64(c[ci, j]) * (1 + sgn * eps))
                c[cj, j] = np.float32(np.float64(c[cj, j]) * (1 - sgn * eps))
            hh[ci] = hh[cj] / 2
            hh[cj] -= hh[ci]
            nsplit += 1
    assert ns == nsplit == 2
    assert np.array_equal(bits(c2), bits(c)) and np.array_equal(h2, hh) and h2.sum() == 15


# ---- independent third-party cross-checks (SURVEY.md §8c plan item 2): none of them is faiss, all of them are somebody
# else's brute-force kNN / k-means step, so they pin the oracle's arithmetic (not its tie rule) from outside this repo ----
@pytest.mark.parametrize("metric", [oracle.IP, oracle.L2])
def test_knn_agrees_with_sklearn_and_torch_brute_force(metric):
    import torch
    from sklearn.neighbors import NearestNeighbors
    x, q, k = gauss(3000, 48, 70), gauss(40, 48, 71), 7
    D, I = oracle.knn(x, q, k, metric)
    if metric == oracle.L2:
        nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="sqeuclidean").fit(x.astype(np.float64))
        Ds, Is = nn.kneighbors(q.astype(np.float64))
        Dt, It = torch.cdist(torch.from_numpy(q).double(), torch.from_numpy(x).double()).pow(2).topk(k, dim=1, largest=False)
    else:
        S = q.astype(np.float64) @ x.astype(np.float64).T
        Is = np.argsort(-S, axis=1, kind="stable")[:, :k]
        Ds = np.take_along_axis(S, Is, axis=1)
        Dt, It = (torch.from_numpy(q).double() @ torch.from_numpy(x).double().T).topk(k, dim=1)
    assert np.array_equal(I, Is) and np.array_equal(I, It.numpy())  # Gaussian data: no ties, one right answer
    assert np.allclose(D, Ds, rtol=0, atol=1e-5) and np.allclose(D, Dt.numpy(), rtol=0, atol=1e-5)  # north_star's fp32 tolerance


def test_kmeans_assignment_step_agrees_with_sklearn():
    from sklearn.metrics import pairwise_distances_argmin_min
    x = gauss(4000, 24, 80, normalize=False)
    a, c, obj = oracle.kmeans(x, 16, niter=5)
    lab, dist = pairwise_distances_argmin_min(x.astype(np.float64), c.astype(np.float64), metric="sqeuclidean")
    assert np.array_equal(a, lab)
    # and the first Lloyd update: init = first k of rand_perm(n, seed + 1), centroids = means of sklearn's assignment to them
    init = x[oracle.rand_perm(len(x), 1234 + 1)[:16]]
    lab0, _ = pairwise_distances_argmin_min(x.astype(np.float64), init.astype(np.float64), metric="sqeuclidean")
    means = np.stack([x[lab0 == j].astype(np.float64).mean(0) for j in range(16)])
    _, c1, _ = oracle.kmeans(x, 16, niter=1)
    assert np.allclose(c1, means, rtol=0, atol=1e-5)

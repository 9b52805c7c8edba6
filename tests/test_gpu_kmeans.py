"""The device-resident Lloyd engine (lotus_b200/csrc/kmeans.cu) against the faiss restatement (oracle.kmeans,
faiss/Clustering.cpp as lotus/utils.py:61-65 drives it): assignments and centroids must be BIT-identical, whichever of the
engine's paths decides a point — the register top-2 filter + gap proof, the second-level general pipeline (ties, near-ties),
the vectorised or the generic point-order accumulate, the device-side split_clusters."""
import numpy as np
import pytest

import oracle
from helpers import bits, gauss, grid

pytestmark = pytest.mark.gpu


def _index(gpu, x, dtype):
    if dtype == "bf16":
        xb = gpu.f32_to_bf16_bits(x)
        return gpu.Index(xb, gpu.BF16, 1), gpu.bf16_bits_to_f32(xb)
    return gpu.Index(x, gpu.F32, 1), x


def _mixture(n, d, k_true, seed, spread=1.0):
    rng = np.random.default_rng(seed)
    centers = gauss(k_true, d, seed + 1) * 4
    return (centers[rng.integers(0, k_true, n)] + spread * gauss(n, d, seed + 2, normalize=False)).astype(np.float32)


@pytest.mark.parametrize("n,d,k,dtype,full", [
    (30_000, 48, 640, "bf16", False),     # 3 centroid tiles; 30000 < 640*256: trains on everything
    (30_000, 48, 640, "f32", False),      # tf32 filter
    (9_000, 64, 8, "bf16", False),        # one partial centroid tile (k << 256); 9000 > 8*256: faiss subsampling
    (20_000, 100, 1030, "f32", True),     # 5 tiles; d % 8 != 0 (padded filter copy)
    (6_000, 30, 17, "bf16", True),        # bf16 rows of 60 bytes: generic (non-vectorised) accumulate, padded query copy
    (5_000, 768, 300, "bf16", True),      # the benchmark's row size
])
def test_lloyd_engine_is_bit_identical_to_the_restatement(gpu, n, d, k, dtype, full):
    x = _mixture(n, d, min(k, 50), 100 + n % 97 + d)
    idx, xf = _index(gpu, x, dtype)
    a, c, obj = idx.kmeans(k, niter=5, full_lloyd=full)
    ao, co, oo = oracle.kmeans(xf, k, niter=5, full_lloyd=full)
    assert np.array_equal(a, ao), f"{(a != ao).sum()} of {n} assignments differ"
    assert np.array_equal(bits(c), bits(co)), "centroids are not bit-identical (point-order fp32 sums)"
    assert np.allclose(obj, oo, rtol=1e-5)
    idx.close()


def test_unstructured_gaussian_points_take_the_second_level_path(gpu):
    # no cluster structure: the best two centroids of many points are within the filter's error bound of each other
    x = gauss(20_000, 32, 7, normalize=False)
    idx, xf = _index(gpu, x, "bf16")
    gpu.stats_reset()
    a, c, obj = idx.kmeans(512, niter=4, full_lloyd=True)
    ao, co, oo = oracle.kmeans(xf, 512, niter=4, full_lloyd=True)
    assert np.array_equal(a, ao) and np.array_equal(bits(c), bits(co)) and np.allclose(obj, oo, rtol=1e-5)
    idx.close()


def test_exact_ties_go_to_the_lowest_centroid_id(gpu):
    # quantised grid: distances tie exactly all over the place; the gap proof must refuse and the exact path must answer
    x = grid(4000, 6, 11)
    idx, xf = _index(gpu, x, "f32")
    a, c, obj = idx.kmeans(40, niter=4, full_lloyd=True)
    ao, co, oo = oracle.kmeans(xf, 40, niter=4, full_lloyd=True)
    assert np.array_equal(a, ao) and np.array_equal(bits(c), bits(co))
    cent = grid(300, 6, 12)                                  # duplicates among the centroids: argmin ties -> lowest id
    cent[100:200] = cent[:100]
    a2, dist = idx.kmeans_assign(cent)
    Dk, Ik = oracle.knn(cent, xf, 1, oracle.L2)
    assert np.array_equal(a2, Ik[:, 0]) and np.array_equal(bits(dist), bits(Dk[:, 0]))
    idx.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_empty_clusters_are_split_like_faiss(gpu, dtype):
    # 70 % of the points are copies of ONE point: the initial centroids (first k of rand_perm) contain that point several
    # times, every copy but the lowest-numbered one attracts nothing -> split_clusters runs (device MT19937 replay)
    rng = np.random.default_rng(21)
    x = gauss(1500, 16, 22, normalize=False)
    dup = rng.random(1500) < 0.7
    x[dup] = x[0]
    idx, xf = _index(gpu, x, dtype)
    a, c, obj = idx.kmeans(24, niter=6, full_lloyd=True)
    ao, co, oo = oracle.kmeans(xf, 24, niter=6, full_lloyd=True)
    assert np.array_equal(a, ao), f"{(a != ao).sum()} assignments differ"
    assert np.array_equal(bits(c), bits(co))
    assert np.allclose(obj, oo, rtol=1e-5, atol=1e-3)
    idx.close()


def test_assign_and_accumulate_device_entry_points(gpu):
    import torch
    x = _mixture(12_345, 64, 9, 31)
    idx, xf = _index(gpu, x, "bf16")
    dev = torch.device("cuda", 0)
    cent_h = _mixture(700, 64, 9, 33)
    cent = torch.from_numpy(cent_h).to(dev)
    assign = torch.empty(len(x), dtype=torch.int64, device=dev)
    dist = torch.empty(len(x), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    idx.kmeans_assign_dev(cent.data_ptr(), 700, assign.data_ptr(), dist_ptr=dist.data_ptr(), stream=st)
    Dk, Ik = oracle.knn(cent_h, xf, 1, oracle.L2)
    assert np.array_equal(assign.cpu().numpy(), Ik[:, 0]) and np.array_equal(bits(dist.cpu().numpy()), bits(Dk[:, 0]))
    sums = torch.empty((700, 64), dtype=torch.float32, device=dev)
    counts = torch.empty(700, dtype=torch.float32, device=dev)
    obj = torch.zeros(1, dtype=torch.float64, device=dev)
    idx.kmeans_accumulate_dev(assign.data_ptr(), 700, sums.data_ptr(), counts.data_ptr(), centroids_ptr=cent.data_ptr(),
                              obj_ptr=obj.data_ptr(), stream=st)
    want = np.zeros((700, 64), np.float32)
    for i, cc in enumerate(Ik[:, 0]):          # point order, fp32: faiss compute_centroids before its normalisation
        want[cc] += xf[i]
    assert np.array_equal(bits(sums.cpu().numpy()), bits(want))
    assert np.array_equal(counts.cpu().numpy(), np.bincount(Ik[:, 0], minlength=700).astype(np.float32))
    assert np.isclose(float(obj.item()), float(Dk[:, 0].astype(np.float64).sum()), rtol=1e-6)
    # a subset through ids, host entry points
    ids = np.arange(5, len(x), 7)
    a2, d2 = idx.kmeans_assign(cent_h, ids=ids)
    assert np.array_equal(a2, Ik[ids, 0]) and np.array_equal(bits(d2), bits(Dk[ids, 0]))
    idx.close()


def test_as_many_points_as_centroids_and_tiny_inputs(gpu):
    x = gauss(12, 8, 41, normalize=False)
    idx, xf = _index(gpu, x, "f32")
    a, c, _ = idx.kmeans(12, niter=3)                       # "just copying"
    ao, co, _ = oracle.kmeans(xf, 12, niter=3)
    assert np.array_equal(a, ao) and np.array_equal(bits(c), bits(co))
    a, c, _ = idx.kmeans(1, niter=2)
    ao, co, _ = oracle.kmeans(xf, 1, niter=2)
    assert np.array_equal(a, ao) and np.array_equal(bits(c), bits(co))
    idx.close()

"""Row-sharding logic under torch.distributed (gloo, world_size 2, CPU): shard bounds, per-shard flat search (the
oracle stands in for the per-GPU kernels), ONE all-gather of the (score, idx) lists, k-way merge rule. The merged
result must equal the single-index search."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import bits, gauss


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, metric, k, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle
    from lotus_b200.distributed import merge_host_lists, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, q = gauss(1001, 24, 0), gauss(33, 24, 1)
    lo, hi = shard_bounds(len(x), world, rank)
    D, I = oracle.knn(x[lo:hi], q, k, metric)
    I = np.where(I >= 0, I + lo, -1)
    ts, ti = torch.from_numpy(D), torch.from_numpy(I)
    gs = [torch.empty_like(ts) for _ in range(world)]
    gi = [torch.empty_like(ti) for _ in range(world)]
    dist.all_gather(gs, ts)
    dist.all_gather(gi, ti)
    ms, mi = merge_host_lists(torch.stack(gs).numpy(), torch.stack(gi).numpy(), metric)
    Dw, Iw = oracle.knn(x, q, k, metric)
    ok = np.array_equal(mi, Iw) and np.array_equal(bits(ms), bits(Dw))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(os.path.join(out_dir, f"ok_{metric}_{k}"), "w").write(str(int(flag.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("metric,k", [(0, 5), (1, 32)])
def test_world2_sharded_search_equals_single_index(tmp_path, metric, k):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, metric, k, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / f"ok_{metric}_{k}").read() == "1"


def test_shard_bounds_cover_without_overlap():
    from lotus_b200.distributed import shard_bounds
    for n in (0, 1, 7, 1000, 1_000_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_merge_rule_on_exact_ties():
    from lotus_b200.distributed import merge_host_lists
    # two shards, one query, k=2, all scores equal: IP keeps faiss's descending-id order (highest shard first)
    s = np.full((2, 1, 2), 0.5, np.float32)
    i = np.array([[[1, 0]], [[3, 2]]], dtype=np.int64)  # per-shard lists already in faiss tie order (id desc for IP)
    ms, mi = merge_host_lists(s, i, 0)
    assert mi.tolist() == [[3, 2]]
    i = np.array([[[0, 1]], [[2, 3]]], dtype=np.int64)  # L2: id ascending
    ms, mi = merge_host_lists(s, i, 1)
    assert mi.tolist() == [[0, 1]]
    i = np.array([[[0, -1]], [[-1, -1]]], dtype=np.int64)
    ms, mi = merge_host_lists(s, i, 1)
    assert mi.tolist() == [[0, -1]] and ms[0, 1] == np.finfo(np.float32).max


def test_split_clusters_host_matches_oracle():
    import oracle
    from lotus_b200.distributed import split_clusters_host
    cent = gauss(6, 10, 3, normalize=False)
    h = np.array([7, 0, 4, 0, 9, 1], np.float32)
    c1, h1 = split_clusters_host(cent, h, n=21)
    c2, h2, ns = oracle.split_clusters(cent, h, n=21)
    assert ns == 2 and np.array_equal(bits(c1), bits(c2)) and np.array_equal(h1, h2)

"""Host logic of the plugin class `B200VS` (lotus_b200/vs.py) on CPU: the native `Index` is replaced by a numpy/oracle-backed
fake, so what is checked here is everything AROUND the C-ABI calls — argument validation, the faiss-compatible directory
layout, the per-directory cache, dtype handling, error mapping (lotus/vector_store/faiss_vs.py:13-77 is the contract).
The same class over the real library runs in tests/test_gpu_ops.py."""
import os
import pickle
import time

import numpy as np
import pytest

import lotus_b200 as lotus
import oracle
from helpers import FakeIndex, gauss
from lotus_b200 import _native as nv
from lotus_b200 import faiss_io
from lotus_b200.vs import BF16Backed, B200VS


@pytest.fixture
def fake_native(monkeypatch):
    FakeIndex.live = 0
    monkeypatch.setattr(nv, "Index", FakeIndex)
    monkeypatch.setattr(nv, "require_device", lambda: None)
    yield
    assert FakeIndex.live >= 0


def test_constructor_rejects_what_the_flat_backend_cannot_do():
    with pytest.raises(ValueError, match="flat"):
        B200VS(factory_string="IVF100,Flat")
    with pytest.raises(ValueError, match="metric"):
        B200VS(metric=7)
    with pytest.raises(ValueError, match="dtype"):
        B200VS(dtype="fp8")
    vs = B200VS()
    assert vs.index_dir is None and vs.metric == faiss_io.METRIC_INNER_PRODUCT
    with pytest.raises(ValueError, match="Index not loaded"):   # faiss_vs.py:54-55
        vs(np.zeros((1, 4), np.float32), 1)


def test_index_writes_the_reference_directory_layout_and_searches(fake_native, tmp_path):
    x, q = gauss(50, 16, 1), gauss(4, 16, 2)
    vs = B200VS()
    d = str(tmp_path / "idx")
    vs.index(None, x, d)
    assert vs.index_dir == d and sorted(os.listdir(d)) == ["index", "vecs"]        # faiss_vs.py:27-30
    with open(f"{d}/vecs", "rb") as fp:
        assert np.array_equal(pickle.load(fp), x)
    xr, metric = faiss_io.read_flat_index(f"{d}/index")
    assert metric == 0 and np.array_equal(xr, x)
    out = vs(q, 5)
    D, I = oracle.knn(x, q, 5, oracle.IP)
    assert np.array_equal(out.indices, I) and np.array_equal(out.distances, D)
    sub = [7, 3, 30, 31]
    out = vs(q, 3, ids=sub)
    D, I = oracle.knn_subset(x, q, 3, np.asarray(sub), oracle.IP)
    assert np.array_equal(out.indices, I) and np.array_equal(out.distances, D)
    assert np.array_equal(vs.get_vectors_from_index(d, [4, 2]), x[[4, 2]])           # faiss_vs.py:38-41
    with pytest.raises(ValueError, match="dimension"):
        vs(gauss(2, 8, 3), 1)
    with pytest.raises(ValueError, match="outside"):                                 # B2_ERANGE -> ValueError
        vs(q, 1, ids=[0, 99])
    with pytest.raises(ValueError, match="not supported"):
        vs(q, 5000)


def test_directories_written_by_the_reference_layout_load_and_metric_is_checked(fake_native, tmp_path):
    x = gauss(20, 8, 4)
    d = str(tmp_path / "faiss_made")
    os.makedirs(d)
    faiss_io.write_flat_index(f"{d}/index", x, faiss_io.METRIC_L2)                   # what faiss.write_index produces for IndexFlatL2
    with open(f"{d}/vecs", "wb") as fp:
        pickle.dump(x.astype(np.float64), fp)                                        # LiteLLMRM hands float64 to FaissVS
    with pytest.raises(ValueError, match="metric"):
        B200VS().load_index(d)
    vs = B200VS(metric=faiss_io.METRIC_L2)
    vs.load_index(d)
    out = vs(x[:3], 2)
    D, I = oracle.knn(x, x[:3], 2, oracle.L2)
    assert np.array_equal(out.indices, I) and np.array_equal(out.distances, D)
    with pytest.raises(ValueError, match="not found"):
        vs.load_index(str(tmp_path / "missing"))


def test_per_directory_cache_reuses_evicts_and_notices_rewrites(fake_native, tmp_path):
    vs = B200VS(cache_size=2)
    dirs = [str(tmp_path / f"d{i}") for i in range(3)]
    for i, d in enumerate(dirs[:2]):
        vs.index(None, gauss(10 + i, 8, 10 + i), d)
    a = vs.b2_index
    vs.load_index(dirs[0])
    b = vs.b2_index
    vs.load_index(dirs[1])
    assert vs.b2_index is a and b is not a and FakeIndex.live == 2                   # flipping between two dirs builds nothing
    vs.index(None, gauss(12, 8, 12), dirs[2])                                        # third directory: the oldest entry goes
    assert FakeIndex.live == 2 and b.closed and not a.closed
    # the directory is rewritten behind our back (another process re-indexed it): the stale copy must not be served
    time.sleep(0.01)
    newx = gauss(10, 8, 99)
    faiss_io.write_index_dir(dirs[1], newx, newx, 0)
    os.utime(f"{dirs[1]}/index", (time.time() + 5, time.time() + 5))
    vs.load_index(dirs[1])
    assert vs.b2_index is not a and np.array_equal(vs.b2_index.vals, newx)
    vs.close()
    assert FakeIndex.live == 0 and vs.b2_index is None


def test_bf16_store_rounds_once_and_hands_query_bits_back(fake_native, tmp_path):
    x = gauss(30, 16, 5)
    vs = B200VS(dtype="bf16")
    d = str(tmp_path / "b")
    vs.index(None, x, d)
    xb = nv.bf16_bits_to_f32(nv.f32_to_bf16_bits(x))
    assert vs.b2_index.dtype == nv.BF16 and np.array_equal(vs.b2_index.vals, xb)
    assert np.array_equal(faiss_io.read_flat_index(f"{d}/index")[0], x)               # the faiss file keeps what faiss would: fp32
    got = vs.get_vectors_from_index(d, [1, 5, 9])
    assert isinstance(got, BF16Backed) and got.dtype == np.float32 and np.array_equal(np.asarray(got), xb[[1, 5, 9]])
    out = vs(got, 4)                                                                 # the operator's round trip (sem_sim_join.py:112-134)
    assert vs.b2_index.calls[-1][:2] == (np.dtype(np.uint16), nv.BF16)
    D, I = oracle.knn(xb, xb[[1, 5, 9]], 4, oracle.IP)
    assert np.array_equal(out.indices, I) and np.array_equal(out.distances, D)
    vs(np.asarray(got) * np.float32(1.0 + 2.0 ** -12), 4)                            # values no longer bf16-exact: plain fp32 queries
    assert vs.b2_index.calls[-1][:2] == (np.dtype(np.float32), nv.F32)


def test_install_configures_whichever_settings_object_the_operators_read(fake_native, monkeypatch):
    import sys
    import types
    import pandas as pd
    try:
        store = lotus.install(dtype="bf16")              # no `lotus` package in this image: our own settings
        assert isinstance(store, B200VS) and lotus.settings.vs is store and store.dtype == "bf16"
    finally:
        lotus.settings.configure(vs=None)
    # with the real package importable, ITS settings and ITS utils.cluster are the ones the reference operators consult,
    # and importing it registers ITS accessors (lotus/sem_ops/*.py) — install() must put ours back AFTER that import
    seen = {}

    class RefDedup:  # what `import lotus` leaves registered under df.sem_dedup
        def __init__(self, obj):
            pass

    fake = types.ModuleType("lotus")
    fake.__path__ = []
    fake.settings = types.SimpleNamespace(configure=lambda **kw: seen.update(kw), rm=None, vs=None, enable_cache=False, lm=None)
    futils = types.ModuleType("lotus.utils")
    futils.cluster = "the faiss one"
    fake.utils = futils
    monkeypatch.setitem(sys.modules, "lotus", fake)
    monkeypatch.setitem(sys.modules, "lotus.utils", futils)
    pd.api.extensions.register_dataframe_accessor("sem_dedup")(RefDedup)
    assert type(pd.DataFrame({"a": [1]}).sem_dedup) is RefDedup
    mine = B200VS()
    assert lotus.install(mine) is mine and seen == {"vs": mine} and futils.cluster is lotus.utils.cluster
    assert lotus.settings.vs is None
    from lotus_b200.sem_ops import SemDedupByDataframe, SemSearchDataframe
    df = pd.DataFrame({"a": [1]})
    assert type(df.sem_dedup) is SemDedupByDataframe and type(df.sem_search) is SemSearchDataframe


def test_get_vectors_from_index_does_not_switch_the_loaded_index(fake_native, tmp_path):
    """ADVICE r1 (vs.py:204): the reference only reads the `vecs` pickle (faiss_vs.py:38-41); the loaded index stays."""
    xa, xb = gauss(12, 8, 21), gauss(9, 8, 22)
    vs = B200VS()
    da, db = str(tmp_path / "a"), str(tmp_path / "b")
    vs.index(None, xb, db)
    vs.index(None, xa, da)
    loaded = vs.b2_index
    got = vs.get_vectors_from_index(db, [3, 0])
    assert np.array_equal(got, xb[[3, 0]])
    assert vs.index_dir == da and vs.b2_index is loaded and np.array_equal(vs.vecs, xa)
    out = vs(got, 2)                                     # searches A, like the reference would
    D, I = oracle.knn(xa, xb[[3, 0]], 2, oracle.IP)
    assert np.array_equal(out.indices, I) and np.array_equal(out.distances, D)
    # a directory that is not cached yet is built and cached, still without becoming the loaded one
    vs2 = B200VS()
    vs2.load_index(da)
    assert np.array_equal(vs2.get_vectors_from_index(db, [1]), xb[[1]]) and vs2.index_dir == da
    assert np.array_equal(vs2(xa[:1], 1).indices, [[0]])


def test_single_process_multi_device_store_merges_like_the_kernel(fake_native, tmp_path):
    """B200VS(devices=[...]): rows sharded over several devices driven from one process; per-device searches + host k-way merge
    must equal the single-index answer (including the ids= subset form and row gathers across shards)."""
    from lotus_b200.vs import MultiDeviceIndex, merge_shard_lists
    x, q = gauss(301, 16, 31), gauss(9, 16, 32)
    for metric, om in ((faiss_io.METRIC_INNER_PRODUCT, oracle.IP), (faiss_io.METRIC_L2, oracle.L2)):
        vs = B200VS(metric=metric, devices=[0, 1, 2])
        d = str(tmp_path / f"md{metric}")
        vs.index(None, x, d)
        assert isinstance(vs.b2_index, MultiDeviceIndex) and [s.device for s in vs.b2_index.shards] == [0, 1, 2]
        out = vs(q, 7)
        D, I = oracle.knn(x, q, 7, om)
        assert np.array_equal(out.indices, I) and np.array_equal(out.distances, D)
        ids = np.arange(3, 301, 4)
        out = vs(q, 5, ids=ids)
        Ds, Is = oracle.knn_subset(x, q, 5, ids, om)
        assert np.array_equal(out.indices, Is) and np.array_equal(out.distances, Ds)
        assert np.array_equal(vs.get_vectors_from_index(d, [300, 0, 150]), x[[300, 0, 150]])
        out = vs(q, 400)                                     # K > n: padded
        assert (np.asarray(out.indices)[:, 301:] == -1).all() and np.array_equal(np.asarray(out.indices)[:, :301], oracle.knn(x, q, 301, om)[1])
        vs.close()
    # exact ties across shards: L2 keeps the lower shard first, IP the higher shard first (merge_topk_kernel's rule)
    s = [np.full((1, 2), 0.5, np.float32), np.full((1, 2), 0.5, np.float32)]
    i_l2 = [np.array([[0, 1]]), np.array([[2, 3]])]
    assert merge_shard_lists(s, i_l2, nv.METRIC_L2)[1].tolist() == [[0, 1]]
    i_ip = [np.array([[1, 0]]), np.array([[3, 2]])]
    assert merge_shard_lists(s, i_ip, nv.METRIC_IP)[1].tolist() == [[3, 2]]

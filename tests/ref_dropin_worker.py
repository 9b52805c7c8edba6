"""Worker of tests/test_gpu_reference_dropin.py (own process: it imports the real `lotus`).

    python tests/ref_dropin_worker.py

Runs the REFERENCE's own, unmodified operator classes — lotus/sem_ops/{sem_index,sem_sim_join,sem_search,sem_dedup,
sem_cluster_by}.py, lotus/utils.py — with `lotus.settings.configure(vs=B200VS())` (installed through lotus_b200.install()) and
replays the scenarios of tests/golden/reference_ops.json, which were produced by the same operator code over the reference's
FaissVS (faiss stand-in = oracle). Prints one JSON object {scenario: bool}."""
import json
import os
import sys
import tempfile
import zlib

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import refpkg  # noqa: E402


def hash_embed(docs, dim):
    out = np.empty((len(docs), dim), dtype=np.float32)
    for i, d in enumerate(docs):
        v = np.random.default_rng(zlib.crc32(str(d).encode("utf-8"))).standard_normal(dim).astype(np.float32)
        out[i] = v / np.linalg.norm(v)
    return out


def frame_to_json(df):
    cols = [str(c) for c in df.columns]
    data = []
    for row in df.itertuples(index=False):
        data.append([float(np.float32(v)) if isinstance(v, (float, np.floating)) else (int(v) if isinstance(v, (int, np.integer)) else v)
                     for v in row])
    return {"index": [int(i) for i in df.index], "columns": cols, "data": data}


def main():
    lotus, faiss_kind = refpkg.import_reference()
    if lotus is None:
        print(json.dumps({"unavailable": faiss_kind}))
        return 0
    from lotus.models import RM
    from lotus.sem_ops.sem_cluster_by import SemClusterByDataframe
    from lotus.sem_ops.sem_dedup import SemDedupByDataframe
    from lotus.sem_ops.sem_index import SemIndexDataframe
    from lotus.sem_ops.sem_search import SemSearchDataframe
    from lotus.sem_ops.sem_sim_join import SemSimJoinDataframe
    from lotus.vector_store.vs import VS as RefVS
    import lotus_b200

    class RefHashRM(RM):
        def __init__(self, dim):
            super().__init__()
            self.dim = dim

        def _embed(self, docs):
            return hash_embed(docs, self.dim)

    class RefTableRM(RM):
        def __init__(self, table):
            super().__init__()
            self.table = table

        def _embed(self, docs):
            return np.stack([np.asarray(self.table[d], dtype=np.float32) for d in docs]).astype(np.float32)

    if os.environ.get("B2_TEST_FAKE_NATIVE") == "1":
        # CPU dry run of this worker (tests/test_reference_dropin_cpu.py): the native index is the oracle-backed fake, so only the
        # plumbing is exercised — the real lotus package, install() ordering, isinstance, the reference's operator code over B200VS
        import oracle
        from helpers import FakeIndex
        from lotus_b200 import _native as _nv
        _nv.Index = FakeIndex
        _nv.require_device = lambda: None
        _nv.connected_components = lambda n, pi, pj, device=0: oracle.connected_components(n, pi, pj)
        _nv.stats = lambda: {"launches": 1}
    gold = json.load(open(os.path.join(HERE, "golden", "reference_ops.json")))
    out = {"faiss": faiss_kind, "reference_dir": refpkg.reference_dir()}
    store = lotus_b200.install(dtype="f32")
    out["b200vs_is_a_reference_VS"] = isinstance(store, RefVS) and lotus.settings.vs is store
    out["cluster_replaced"] = lotus.utils.cluster is lotus_b200.utils.cluster
    # install() re-registered lotus_b200's accessors after importing lotus: df.sem_dedup is the streaming one
    out["accessors_are_ours_after_install"] = type(pd.DataFrame({"a": [1]}).sem_dedup).__module__.startswith("lotus_b200")
    tmp = tempfile.mkdtemp()

    # ---- the REFERENCE's classes, called directly (whatever is registered under df.sem_* does not matter here) -------------
    sc = gold["sim_join"]
    lotus.settings.configure(rm=RefHashRM(sc["dim"]))
    left = pd.DataFrame({"a": [f"l{i}" for i in range(7)], "v": list(range(7))})
    right = pd.DataFrame({"b": [f"r{i}" for i in range(11)], "v": list(range(11))})
    right = SemIndexDataframe(right)("b", os.path.join(tmp, "r"))
    ok = True
    for case in sc["cases"]:
        other = right if case["right_rows"] is None else right.loc[case["right_rows"]]
        kw = dict(case["kwargs"])
        if case["right_rows"] is not None and kw["K"] > len(other):
            continue  # the reference wraps -1 to the last id here (faiss_vs.py:71-72); B200VS reports -1 (documented divergence)
        got = SemSimJoinDataframe(left)(other, "a", "b", **kw)
        ok &= frame_to_json(got) == case["out"]
    out["ref_sem_sim_join_over_B200VS"] = bool(ok)

    ss = gold["search"]
    lotus.settings.configure(rm=RefHashRM(ss["dim"]))
    docs = SemIndexDataframe(pd.DataFrame({"t": [f"doc{i}" for i in range(ss["n"])]}))("t", os.path.join(tmp, "s"))
    ok = True
    for case in ss["cases"]:
        frame = docs if case["rows"] is None else docs.loc[case["rows"]]
        got = SemSearchDataframe(frame)("t", case["query"], K=case["K"], return_scores=True)
        ok &= frame_to_json(got) == case["out"]
    out["ref_sem_search_over_B200VS"] = bool(ok)

    dd = gold["dedup"]
    lotus.settings.configure(rm=RefHashRM(dd["dim"]))
    frame = SemIndexDataframe(pd.DataFrame({"Text": dd["values"]}))("Text", os.path.join(tmp, "d"))
    kept_ref = SemDedupByDataframe(frame)("Text", threshold=dd["threshold"])            # the reference's N^2 sem_dedup, vs = B200VS
    kept_ours = frame.sem_dedup("Text", threshold=dd["threshold"])                        # lotus_b200's streaming accessor
    out["ref_sem_dedup_over_B200VS"] = bool(len(kept_ref) == len(dd["kept_values"]) and len(set(kept_ref["Text"])) == len(set(dd["kept_values"])))
    out["streaming_sem_dedup_same_survivor_count"] = bool(len(kept_ours) == len(kept_ref))

    cl = gold["cluster"]
    lotus.settings.configure(rm=RefTableRM({k: np.asarray(v, dtype=np.float32) for k, v in cl["table"].items()}))
    cf = SemIndexDataframe(pd.DataFrame({"name": cl["names"]}))("name", os.path.join(tmp, "c"))
    got = SemClusterByDataframe(cf)("name", cl["ncentroids"], niter=cl["niter"])       # reference class -> lotus.utils.cluster (ours)
    out["ref_sem_cluster_by_over_device_kmeans"] = bool([int(c) for c in got["cluster_id"]] == cl["cluster_id"])

    from lotus_b200 import _native as nv
    out["filter_or_dense_kernels_ran"] = bool(nv.stats()["launches"] > 0)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())

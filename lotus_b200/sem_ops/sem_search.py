"""df.sem_search(col, query, K, n_rerank, return_scores, suffix) — lotus/sem_ops/sem_search.py:91-157.

Same contract: top-K rows of THIS frame for one query, best first, optional `vec_scores{suffix}` column, optional
rerank. The reference searches the whole index and post-filters with a K-doubling retry loop (:116-138); when the
vector store accepts `ids` we pass the frame's own row ids instead, which yields the same rows in
one search. Stores without `ids` support keep the reference loop.

Multi-query form (not in the reference): `query` may be a list of strings (or a [Q, d] matrix of precomputed
query vectors) — ONE batched `vs(query_vectors[Q, d], K, ids=...)` call answers all of them and a list of Q frames comes back,
each identical to what the single-query call returns for that query. `df.sem_search.batch(...)` is the same thing by name."""
from __future__ import annotations

from typing import Any

import numpy as np
import pandas as pd

from ..cache import operator_cache
from ._common import active_settings, register, resolve_rm_vs, validate_df


@register("sem_search")
class SemSearchDataframe:
    def __init__(self, pandas_obj: Any) -> None:
        validate_df(pandas_obj)
        self._obj = pandas_obj

    @operator_cache
    def __call__(self, col_name: str, query: Any, K: int | None = None, n_rerank: int | None = None,
                 return_scores: bool = False, suffix: str = "_sim_score") -> Any:
        assert not (K is None and n_rerank is None), "K or n_rerank must be provided"
        many = isinstance(query, (list, tuple)) or (isinstance(query, np.ndarray) and query.ndim == 2) or _is_2d_tensor(query)
        queries = list(query) if isinstance(query, (list, tuple)) else query
        nq = len(queries) if many else 1
        if K is not None:
            rm, vs = resolve_rm_vs()
            col_index_dir = self._obj.attrs["index_dirs"][col_name]
            if vs.index_dir != col_index_dir:
                vs.load_index(col_index_dir)
            assert vs.index_dir == col_index_dir

            df_idxs = self._obj.index
            K = min(K, len(df_idxs))
            frames = []
            if K <= 0:  # the reference reaches faiss with k=0 (sem_search.py:118-122); return the empty frame
                for _ in range(nq):
                    new_df = self._obj.iloc[0:0].copy()
                    new_df.attrs["index_dirs"] = self._obj.attrs.get("index_dirs", None)
                    if return_scores:
                        new_df["vec_scores" + suffix] = np.zeros(0, dtype=np.float32)
                    frames.append(new_df)
            else:
                query_vectors = rm.convert_query_to_query_vector(queries if many else query)
                if getattr(vs, "supports_ids_search", True):
                    ids = np.asarray(df_idxs, dtype=np.int64) if getattr(vs, "accepts_id_arrays", False) else list(df_idxs)
                    out = vs(query_vectors, K, ids=ids)  # one batched search for all the queries
                    rows = [([int(i) for i in out.indices[r] if i != -1],
                             [float(s) for s, i in zip(out.distances[r], out.indices[r]) if i != -1]) for r in range(nq)]
                else:  # reference behaviour: post-filter + K doubling (per batch: every query must reach K rows)
                    search_K = K
                    idx_set = set(df_idxs)
                    while True:
                        out = vs(query_vectors, search_K)
                        pairs = [[(int(i), float(s)) for i, s in zip(out.indices[r], out.distances[r]) if i in idx_set] for r in range(nq)]
                        if all(len(p) >= K for p in pairs) or search_K >= 2 * max(len(out.indices[0]), 1) ** 2:
                            break
                        search_K *= 2
                    rows = [([p[0] for p in pr[:K]], [p[1] for p in pr[:K]]) for pr in pairs]
                for doc_idxs, scores in rows:
                    new_df = self._obj.loc[doc_idxs]
                    new_df.attrs["index_dirs"] = self._obj.attrs.get("index_dirs", None)
                    if return_scores:
                        new_df["vec_scores" + suffix] = scores
                    frames.append(new_df)
        else:
            frames = [self._obj for _ in range(nq)]

        if n_rerank is not None:
            reranker = active_settings().reranker
            if reranker is None:
                raise ValueError("Reranker not found in settings")
            for r in range(nq):
                docs = frames[r][col_name].tolist()
                reranked_output = reranker(queries[r] if many else query, docs, n_rerank)
                frames[r] = frames[r].iloc[reranked_output.indices]
        return frames if many else frames[0]

    def batch(self, col_name: str, queries: Any, K: int | None = None, n_rerank: int | None = None, return_scores: bool = False,
              suffix: str = "_sim_score") -> list:
        """Several queries, one device search: list of frames, one per query."""
        qs = queries if (isinstance(queries, np.ndarray) or _is_2d_tensor(queries)) else list(queries)
        return self(col_name, qs, K=K, n_rerank=n_rerank, return_scores=return_scores, suffix=suffix)


def _is_2d_tensor(q: Any) -> bool:
    try:
        import torch
        return isinstance(q, torch.Tensor) and q.dim() == 2
    except Exception:  # pragma: no cover
        return False

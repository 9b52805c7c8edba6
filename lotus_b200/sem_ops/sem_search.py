"""df.sem_search(col, query, K, n_rerank, return_scores, suffix) — lotus/sem_ops/sem_search.py:91-157.

Same contract: top-K rows of THIS frame for one query, best first, optional `vec_scores{suffix}` column, optional
rerank. The reference searches the whole index and post-filters with a K-doubling retry loop (:116-138); when the
vector store accepts `ids` we pass the frame's own row ids instead, which yields the same rows in
one search. Stores without `ids` support keep the reference loop."""
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
    def __call__(self, col_name: str, query: str, K: int | None = None, n_rerank: int | None = None,
                 return_scores: bool = False, suffix: str = "_sim_score") -> pd.DataFrame:
        assert not (K is None and n_rerank is None), "K or n_rerank must be provided"
        if K is not None:
            rm, vs = resolve_rm_vs()
            col_index_dir = self._obj.attrs["index_dirs"][col_name]
            if vs.index_dir != col_index_dir:
                vs.load_index(col_index_dir)
            assert vs.index_dir == col_index_dir

            df_idxs = self._obj.index
            K = min(K, len(df_idxs))
            if K <= 0:  # the reference reaches faiss with k=0 (sem_search.py:118-122); return the empty frame
                new_df = self._obj.iloc[0:0].copy()
                new_df.attrs["index_dirs"] = self._obj.attrs.get("index_dirs", None)
                if return_scores:
                    new_df["vec_scores" + suffix] = np.zeros(0, dtype=np.float32)
            else:
                query_vectors = rm.convert_query_to_query_vector(query)
                if getattr(vs, "supports_ids_search", True):
                    out = vs(query_vectors, K, ids=list(df_idxs))
                    doc_idxs = [int(i) for i in out.indices[0] if i != -1]
                    scores = [float(s) for s, i in zip(out.distances[0], out.indices[0]) if i != -1]
                else:  # reference behaviour: post-filter + K doubling
                    search_K = K
                    while True:
                        out = vs(query_vectors, search_K)
                        idx_set = set(df_idxs)
                        pairs = [(int(i), float(s)) for i, s in zip(out.indices[0], out.distances[0]) if i in idx_set]
                        if len(pairs) >= K or search_K >= 2 * max(len(out.indices[0]), 1) ** 2:
                            break
                        search_K *= 2
                    doc_idxs = [p[0] for p in pairs[:K]]
                    scores = [p[1] for p in pairs[:K]]
                new_df = self._obj.loc[doc_idxs]
                new_df.attrs["index_dirs"] = self._obj.attrs.get("index_dirs", None)
                if return_scores:
                    new_df["vec_scores" + suffix] = scores
        else:
            new_df = self._obj

        if n_rerank is not None:
            reranker = active_settings().reranker
            if reranker is None:
                raise ValueError("Reranker not found in settings")
            docs = new_df[col_name].tolist()
            reranked_output = reranker(query, docs, n_rerank)
            new_df = new_df.iloc[reranked_output.indices]
        return new_df

"""left.sem_sim_join(right, left_on, right_on, K, lsuffix, rsuffix, score_suffix, keep_index) — the north-star
operator, lotus/sem_ops/sem_sim_join.py:84-166.

Control flow follows the reference line by line (query vectors from the left index when it exists, `vs.load_index`
when the directory differs, `vs(query_vectors, K, ids=list(right.index))`, post-filter `res_id != -1 and res_id in
right.index`, then the same two pandas joins so column order, suffixes and the repeated left index come out
identically). The O(Q*K) Python double loop of :142-145 is replaced by the same filter on whole arrays."""
from __future__ import annotations

from typing import Any

import numpy as np
import pandas as pd

from ..cache import operator_cache
from ._common import register, resolve_rm_vs, validate_df


@register("sem_sim_join")
class SemSimJoinDataframe:
    def __init__(self, pandas_obj: Any):
        validate_df(pandas_obj)
        self._obj = pandas_obj

    @operator_cache
    def __call__(self, other: pd.DataFrame, left_on: str, right_on: str, K: int, lsuffix: str = "", rsuffix: str = "",
                 score_suffix: str = "", keep_index: bool = False) -> pd.DataFrame:
        if isinstance(other, pd.Series):
            if other.name is None:
                raise ValueError("Other Series must have a name")
            other = pd.DataFrame({other.name: other})

        rm, vs = resolve_rm_vs(strict=True)

        # load query embeddings from index if they exist
        if left_on in self._obj.attrs.get("index_dirs", []):
            query_index_dir = self._obj.attrs["index_dirs"][left_on]
            if vs.index_dir != query_index_dir:
                vs.load_index(query_index_dir)
            assert vs.index_dir == query_index_dir
            try:
                queries = vs.get_vectors_from_index(query_index_dir, self._obj.index)
            except NotImplementedError:
                queries = self._obj[left_on]
        else:
            queries = self._obj[left_on]

        # load index to search over
        try:
            col_index_dir = other.attrs["index_dirs"][right_on]
        except KeyError:
            raise ValueError(f"Index directory for column {right_on} not found in DataFrame")
        if vs.index_dir != col_index_dir:
            vs.load_index(col_index_dir)
        assert vs.index_dir == col_index_dir

        query_vectors = rm.convert_query_to_query_vector(queries)
        right_ids = list(other.index)
        vs_output = vs(query_vectors, K, ids=right_ids)
        distances = np.asarray(vs_output.distances)
        indices = np.asarray(vs_output.indices)

        # post filter (sem_sim_join.py:142-145), vectorised
        nq = indices.shape[0]
        kk = indices.shape[1] if indices.ndim == 2 else 0
        flat_ids = indices.reshape(-1)
        keep = (flat_ids != -1) & np.isin(flat_ids, np.asarray(other.index))
        left_labels = np.repeat(np.asarray(self._obj.index)[:nq], kk)[keep] if kk else np.asarray([], dtype=object)
        temp_df = pd.DataFrame({"_left_id": left_labels, "_right_id": flat_ids[keep],
                                "_scores" + score_suffix: distances.reshape(-1)[keep]})

        df1 = self._obj.copy()
        df2 = other.copy()
        df1["_left_id"] = df1.index
        df2["_right_id"] = df2.index
        joined_df = df1.join(temp_df.set_index("_left_id"), how="right", on="_left_id").join(
            df2.set_index("_right_id"), how="left", on="_right_id", lsuffix=lsuffix, rsuffix=rsuffix)
        if not keep_index:
            joined_df.drop(columns=["_left_id", "_right_id"], inplace=True)
        return joined_df

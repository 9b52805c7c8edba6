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
        # the reference passes `list(right.index)`; our own store takes the label array as is (no 1M-element Python list)
        right_ids = np.asarray(other.index) if getattr(vs, "accepts_id_arrays", False) else list(other.index)
        vs_output = vs(query_vectors, K, ids=right_ids)
        distances = np.asarray(vs_output.distances)
        indices = np.asarray(vs_output.indices)

        # post filter (sem_sim_join.py:142-145: `res_id != -1 and res_id in right.index`), on whole arrays
        nq = indices.shape[0]
        kk = indices.shape[1] if indices.ndim == 2 else 0
        flat_ids = indices.reshape(-1)
        right_pos = label_positions(other.index, flat_ids)  # -1 where the label is absent; None if labels repeat
        if right_pos is not None:
            keep = (flat_ids != -1) & (right_pos >= 0)
            right_pos = right_pos[keep]
        else:
            keep = (flat_ids != -1) & np.isin(flat_ids, np.asarray(other.index))
        left_pos = np.repeat(np.arange(nq, dtype=np.int64), kk)[keep] if kk else np.zeros(0, dtype=np.int64)
        score_col = "_scores" + score_suffix
        fast = assemble_take(self._obj, other, left_pos, flat_ids[keep], distances.reshape(-1)[keep], score_col, lsuffix,
                             rsuffix, keep_index, right_pos=right_pos)
        if fast is not None:
            return fast
        left_labels = np.asarray(self._obj.index)[left_pos] if kk else np.asarray([], dtype=object)
        temp_df = pd.DataFrame({"_left_id": left_labels, "_right_id": flat_ids[keep], score_col: distances.reshape(-1)[keep]})
        return assemble_join(self._obj, other, temp_df, lsuffix, rsuffix, keep_index)


def label_positions(index: pd.Index, labels: np.ndarray):
    """Position of every label in `index` (-1 = absent), or None when the index holds repeated labels."""
    if isinstance(index, pd.RangeIndex) and index.step == 1 and labels.dtype.kind in "iu":
        pos = labels.astype(np.int64, copy=True) - index.start
        pos[(pos < 0) | (pos >= len(index))] = -1
        return pos
    if not index.is_unique:
        return None
    return np.asarray(index.get_indexer(labels), dtype=np.int64)


def assemble_join(left: pd.DataFrame, right: pd.DataFrame, temp_df: pd.DataFrame, lsuffix: str, rsuffix: str,
                  keep_index: bool) -> pd.DataFrame:
    """The reference's own frame assembly (sem_sim_join.py:147-166): two label joins through the match table."""
    df1 = left.copy()
    df2 = right.copy()
    df1["_left_id"] = df1.index
    df2["_right_id"] = df2.index
    joined_df = df1.join(temp_df.set_index("_left_id"), how="right", on="_left_id").join(
        df2.set_index("_right_id"), how="left", on="_right_id", lsuffix=lsuffix, rsuffix=rsuffix)
    if not keep_index:
        joined_df.drop(columns=["_left_id", "_right_id"], inplace=True)
    return joined_df


def assemble_take(left: pd.DataFrame, right: pd.DataFrame, left_pos: np.ndarray, right_labels: np.ndarray, scores: np.ndarray,
                  score_col: str, lsuffix: str, rsuffix: str, keep_index: bool, right_pos=None):
    """Same frame as `assemble_join`, built with one positional `take` per side instead of two hash joins (3.2M output rows
    at the BASELINE shape: the joins cost more than the GPU search). Only for the plain case — unique labels on both
    sides, flat string column names, every match present on both sides — anything else returns None and takes the
    reference's path. tests/test_host_logic.py checks frame equality (values, dtypes, index, column order) of the two."""
    if not (left.index.is_unique and right.index.is_unique) or left.index.nlevels != 1 or right.index.nlevels != 1:
        return None
    lcols, rcols = list(left.columns), list(right.columns)
    if not all(isinstance(c, str) for c in lcols + rcols) or len(set(lcols)) != len(lcols) or len(set(rcols)) != len(rcols):
        return None
    reserved = {"_left_id", "_right_id", score_col}
    if reserved & set(lcols) or reserved & set(rcols) or len(reserved) != 3:
        return None
    if left.index.dtype == object or right.index.dtype == object:  # label dtype inference differs between the two paths
        return None
    overlap = set(lcols) & set(rcols)
    if overlap and not lsuffix and not rsuffix:
        return None  # pandas raises "columns overlap but no suffix specified": let it
    lnames = [c + lsuffix if c in overlap else c for c in lcols]
    rnames = [c + rsuffix if c in overlap else c for c in rcols]
    if len(set(lnames) | set(rnames) | reserved) != len(lnames) + len(rnames) + 3:
        return None  # suffixing created a clash: pandas has its own rules (and warnings) for that
    if right_pos is None:
        right_pos = label_positions(right.index, np.asarray(right_labels))
    if right_pos is None or (len(right_pos) and right_pos.min() < 0):
        return None
    lpart = left.take(left_pos)
    rpart = right.take(right_pos)
    data = {}
    for name, col in zip(lnames, lcols):
        data[name] = lpart[col].array
    out_index = lpart.index
    if keep_index:
        data["_left_id"] = out_index.array
        data["_right_id"] = right_labels
    data[score_col] = scores
    for name, col in zip(rnames, rcols):
        data[name] = rpart[col].array
    out = pd.DataFrame(data, index=out_index, copy=False)
    return out

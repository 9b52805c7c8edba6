"""df.sem_cluster_by(col, ncentroids, return_scores, return_centroids, niter, verbose) —
lotus/sem_ops/sem_cluster_by.py:57-86. Writes `cluster_id` INTO the caller's frame and returns it, like the
reference (:78,86); `return_scores` / `return_centroids` are accepted and ignored, like the reference (:79-85).
`cluster` is resolved at call time (:74) so it can be replaced."""
from __future__ import annotations

from typing import Any

import pandas as pd

from .. import utils as _utils
from ..cache import operator_cache
from ._common import register, resolve_rm_vs, validate_df


@register("sem_cluster_by")
class SemClusterByDataframe:
    def __init__(self, pandas_obj: Any) -> None:
        validate_df(pandas_obj)
        self._obj = pandas_obj

    @operator_cache
    def __call__(self, col_name: str, ncentroids: int, return_scores: bool = False, return_centroids: bool = False,
                 niter: int = 20, verbose: bool = False) -> pd.DataFrame:
        resolve_rm_vs()
        cluster_fn = _utils.cluster(col_name, ncentroids)
        indices = cluster_fn(self._obj, niter, verbose)
        self._obj["cluster_id"] = pd.Series(indices, index=self._obj.index)
        return self._obj

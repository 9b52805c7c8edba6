"""df.sem_partition_by(partition_fn) — lotus/sem_ops/sem_partition_by.py:60-67: store partition_fn(df) as
`_lotus_partition_id` (the typical partition_fn is `utils.cluster(col, k)`)."""
from __future__ import annotations

from typing import Any, Callable

import pandas as pd

from ..cache import operator_cache
from ._common import register, validate_df


@register("sem_partition_by")
class SemPartitionByDataframe:
    def __init__(self, pandas_obj: Any):
        validate_df(pandas_obj)
        self._obj = pandas_obj

    @operator_cache
    def __call__(self, partition_fn: Callable[[pd.DataFrame], list[int]]) -> pd.DataFrame:
        group_ids = partition_fn(self._obj)
        self._obj["_lotus_partition_id"] = pd.Series(group_ids, index=self._obj.index)
        return self._obj

"""df.load_sem_index(col, index_dir) — lotus/sem_ops/load_sem_index.py:49-51: record the directory only;
loading is lazy (the operators call vs.load_index when vs.index_dir differs)."""
from __future__ import annotations

from typing import Any

import pandas as pd

from ._common import register, validate_df


@register("load_sem_index")
class LoadSemIndexDataframe:
    def __init__(self, pandas_obj: Any):
        validate_df(pandas_obj)
        self._obj = pandas_obj
        self._obj.attrs.setdefault("index_dirs", {})

    def __call__(self, col_name: str, index_dir: str) -> pd.DataFrame:
        self._obj.attrs["index_dirs"][col_name] = index_dir
        return self._obj

"""On-disk compatibility with the reference's index directories.

`FaissVS.index` writes `{index_dir}/index` with faiss.write_index and `{index_dir}/vecs` as a pickle of the
ndarray it was given (lotus/vector_store/faiss_vs.py:27-30); `load_index` reads both (:32-36).
The faiss file of an IndexFlat is (faiss/impl/index_write.cpp, write_index + write_index_header):
    fourcc "IxFI" (inner product) | "IxF2" (L2)                      4 bytes
    d            int32          ntotal       int64
    dummy, dummy int64 (1<<20)  is_trained   uint8      metric_type int32   [metric_arg float32 if metric_type > 1]
    size         uint64 = number of float32 values (WRITEXBVECTOR writes codes.size()/4)
    data         float32[ntotal*d]
This module reads and writes that layout with numpy only, so directories produced by either backend load in both.
"""
from __future__ import annotations

import os
import pickle
import struct

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def write_flat_index(path: str, x: np.ndarray, metric: int) -> None:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    fourcc = b"IxFI" if metric == METRIC_INNER_PRODUCT else b"IxF2"
    with open(path, "wb") as f:
        f.write(fourcc)
        f.write(struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric))
        f.write(struct.pack("<Q", n * d))
        x.tofile(f)


def read_flat_index(path: str, mmap: bool = False):
    """-> (x float32 [n,d], metric)."""
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
            raise ValueError(f"{path}: not a faiss IndexFlat file (fourcc {fourcc!r}); only factory string 'Flat' is supported")
        d, n, _d1, _d2, _trained, metric = struct.unpack("<iqqqBi", f.read(4 + 8 * 3 + 1 + 4))
        if metric > 1:
            f.read(4)
        (size,) = struct.unpack("<Q", f.read(8))
        if size != n * d:
            raise ValueError(f"{path}: corrupt IndexFlat payload ({size} values for {n}x{d})")
        off = f.tell()
        if mmap:
            x = np.memmap(path, dtype=np.float32, mode="r", offset=off, shape=(n, d))
        else:
            x = np.fromfile(f, dtype=np.float32, count=n * d).reshape(n, d)
    return x, metric


def write_index_dir(index_dir: str, embeddings, x_f32: np.ndarray, metric: int) -> None:
    os.makedirs(index_dir, exist_ok=True)
    with open(f"{index_dir}/vecs", "wb") as fp:
        pickle.dump(embeddings, fp)
    write_flat_index(f"{index_dir}/index", x_f32, metric)


def read_index_dir(index_dir: str):
    """-> (vecs as stored by the caller of index(), x float32, metric). `vecs` falls back to x when the pickle is absent."""
    if not os.path.isdir(index_dir):
        raise ValueError(f"Index directory {index_dir} not found")
    x, metric = read_flat_index(f"{index_dir}/index")
    vecs = None
    vp = f"{index_dir}/vecs"
    if os.path.exists(vp):
        with open(vp, "rb") as fp:
            vecs = pickle.load(fp)
    return (vecs if vecs is not None else x), x, metric

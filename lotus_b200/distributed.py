"""Row-sharded flat index across the GPUs of one box: one process per GPU, torch.distributed for the
plumbing. The corpus rows [r*N/G, (r+1)*N/G) live on rank r; queries are replicated; every rank runs the fused
tcgen05 filter + exact finalize on its shard, then ONE all-gather of the per-shard (score, idx) lists over
NCCL/NVLink and a single-kernel k-way merge (b2_merge_topk_dev). No other collective touches the data path.

The reference is single-process (faiss_vs.py); the sharded result is defined as faiss `IndexShards` would define
it: per-shard flat search, then a merge by (score, shard order). It equals the single-index result whenever no
exact fp32 tie straddles rank K across shards."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _native as nv


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row range of `rank`: the first n % world ranks get one extra row."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedIndex:
    """Each rank holds `local_rows` (a torch CUDA tensor [n_local, d], bf16 or fp32) = its slice of the corpus."""

    def __init__(self, local_rows, row_offset: int, metric: int = nv.METRIC_IP, group=None):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        assert local_rows.is_cuda and local_rows.dim() == 2 and local_rows.is_contiguous()
        self.device = local_rows.device.index if local_rows.device.index is not None else torch.cuda.current_device()
        self.dtype = nv.BF16 if local_rows.dtype == torch.bfloat16 else nv.F32
        if self.dtype == nv.F32 and local_rows.dtype != torch.float32:
            raise TypeError("corpus must be float32 or bfloat16")
        self.metric = metric
        self.row_offset = int(row_offset)
        n, d = local_rows.shape
        self.index = nv.Index(None, self.dtype, metric, self.device, on_device_ptr=local_rows.data_ptr(), n=n, d=d)
        self.d = d

    def search(self, q, k: int):
        """q: torch CUDA tensor [nq, d] (bf16 or fp32), replicated on every rank.
        -> (scores [nq,k] float32, idx [nq,k] int64) CUDA tensors holding the GLOBAL top-k on every rank."""
        torch = self.torch
        assert q.is_cuda and q.is_contiguous() and q.shape[1] == self.d
        nq = q.shape[0]
        q_dtype = nv.BF16 if q.dtype == torch.bfloat16 else nv.F32
        stream = torch.cuda.current_stream().cuda_stream
        loc_s = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        loc_i = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        self.index.search_dev(q.data_ptr(), nq, k, q_dtype, loc_s.data_ptr(), loc_i.data_ptr(),
                              id_offset=self.row_offset, stream=stream)
        if self.world == 1:
            return loc_s, loc_i
        all_s = torch.empty((self.world, nq, k), dtype=torch.float32, device=q.device)
        all_i = torch.empty((self.world, nq, k), dtype=torch.int64, device=q.device)
        self.dist.all_gather_into_tensor(all_s, loc_s, group=self.group)
        self.dist.all_gather_into_tensor(all_i, loc_i, group=self.group)
        out_s = torch.empty_like(loc_s)
        out_i = torch.empty_like(loc_i)
        nv.merge_topk_dev(all_s.data_ptr(), all_i.data_ptr(), self.world, nq, k, self.metric, self.device,
                          out_s.data_ptr(), out_i.data_ptr(), stream=stream)
        return out_s, out_i

    def last_filter_ms(self) -> float:
        return self.index.last_filter_ms()

    def close(self) -> None:
        self.index.close()


def merge_host_lists(scores: np.ndarray, idx: np.ndarray, metric: int):
    """Host restatement of the k-way merge rule (used by the gloo CPU tests of the sharding logic only):
    scores/idx [g, nq, k], lists sorted best first -> [nq, k]. Equal scores: L2 keeps list order from the lowest
    shard up; IP from the highest shard down (faiss's heap order)."""
    g, nq, k = scores.shape
    out_s = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    pad = np.finfo(np.float32).max if metric == nv.METRIC_L2 else -np.finfo(np.float32).max
    for q in range(nq):
        ent = []
        for gi in range(g):
            for p in range(k):
                if idx[gi, q, p] >= 0:
                    s = float(scores[gi, q, p])
                    tie = gi * k + p if metric == nv.METRIC_L2 else (g - 1 - gi) * k + p
                    ent.append(((s if metric == nv.METRIC_L2 else -s), tie, gi, p))
        ent.sort(key=lambda e: (e[0], e[1]))
        for o in range(k):
            if o < len(ent):
                _, _, gi, p = ent[o]
                out_s[q, o] = scores[gi, q, p]
                out_i[q, o] = idx[gi, q, p]
            else:
                out_s[q, o] = pad
                out_i[q, o] = -1
    return out_s, out_i

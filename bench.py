#!/usr/bin/env python
"""bench.py — sem_sim_join queries/sec on the BASELINE.json headline configuration.

Workload (BASELINE.json configs[2], the one the metric is quoted on; it fits one GPU): 100k queries x 1M-row
index, 768-d bf16, K=32, synthetic L2-normalised Gaussian embeddings (corpus seed 0, queries seed 1). A "step" is
one pass of the hot path over the whole query batch: `VS.__call__` of the sim-join (fused tcgen05 filter + exact
finalize [+ all-gather + k-way merge when the index is row-sharded over N GPUs]).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...     # the CPU arm: oracle port of the faiss flat path on the host cores

Prints ONE JSON line (rank 0). `value` = device-resident throughput; `e2e` = the same call with HOST buffers
(pinned) including H2D of the queries and D2H of the (score, idx) result every step; `roofline` is the tcgen05
filter kernel against the measured dense bf16 peak (MEASURED_PEAKS.json); `cpu_baseline` is the oracle port on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK_ROWS = 1 << 16  # data is generated in fixed row blocks so every world size sees the same matrix


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def gen_rows_torch(torch, lo: int, hi: int, d: int, seed: int, device, dtype):
    """Rows [lo, hi) of the synthetic matrix: N(0,1) fp32, L2-normalised in fp32, cast to `dtype`."""
    out = torch.empty((hi - lo, d), dtype=dtype, device=device)
    b0, b1 = lo // BLOCK_ROWS, (hi - 1) // BLOCK_ROWS if hi > lo else -1
    for b in range(b0, b1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1_000_003 + b)
        blk = torch.randn((BLOCK_ROWS, d), generator=g, device=device, dtype=torch.float32)
        blk = blk / blk.norm(dim=1, keepdim=True)
        s, e = max(lo, b * BLOCK_ROWS), min(hi, (b + 1) * BLOCK_ROWS)
        out[s - lo:e - lo] = blk[s - b * BLOCK_ROWS:e - b * BLOCK_ROWS].to(dtype)
    return out


def gen_rows_numpy(lo: int, hi: int, d: int, seed: int) -> np.ndarray:
    """CPU generator for the reference arm (different stream of random numbers, same distribution)."""
    out = np.empty((hi - lo, d), dtype=np.float32)
    b0, b1 = lo // BLOCK_ROWS, (hi - 1) // BLOCK_ROWS if hi > lo else -1
    for b in range(b0, b1 + 1):
        rng = np.random.default_rng(seed * 1_000_003 + b)
        blk = rng.standard_normal((BLOCK_ROWS, d), dtype=np.float32)
        blk /= np.linalg.norm(blk, axis=1, keepdims=True)
        s, e = max(lo, b * BLOCK_ROWS), min(hi, (b + 1) * BLOCK_ROWS)
        out[s - lo:e - lo] = blk[s - b * BLOCK_ROWS:e - b * BLOCK_ROWS]
    return out


def to_bf16_values(a: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)  # RNE; inputs are finite
    return r.view(np.float32).reshape(a.shape)


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines: list[str] = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            assert self.proc is not None and self.proc.stdout is not None
            for line in self.proc.stdout:
                self.lines.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            pk = json.load(f)
        if "bf16_tflops_sustained" in pk:
            return float(pk["bf16_tflops_sustained"]), "measured bf16_tflops_sustained (MEASURED_PEAKS.json; kernel timed inside a long step)"
        return float(pk["bf16_tflops"]), "measured bf16_tflops (MEASURED_PEAKS.json)"
    except Exception:
        return 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def pick_cpu_sample(oracle, x, q_pool, k, requested: int, target_s: float) -> int:
    """Size of the bounded CPU sample: `requested` if given, else as many queries as the host cores score in about
    `target_s` seconds (calibrated on one query block per thread against the full index), clamped to [128, 8192]."""
    if requested > 0:
        return min(requested, len(q_pool))
    # one 16-query block per host thread (orc_knn_blocked parallelises over query blocks): a smaller probe would leave
    # cores idle and underestimate the rate
    probe = min(max(64, 16 * oracle.num_threads()), len(q_pool))
    oracle.knn_blocked(x[:50_000], q_pool[:probe], k, oracle.IP)  # spin the thread pool up
    t0 = time.perf_counter()
    oracle.knn_blocked(x, q_pool[:probe], k, oracle.IP)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    want = int(rate * target_s) // 64 * 64
    return int(min(max(want, 128), 8192, len(q_pool)))


def run_reference(args) -> None:
    """CPU arm: the reference's own CPU implementation of the path = faiss flat search, here the oracle port
    (faiss is not installable in this image; oracle/faiss_flat.c restates it). Rank 0 only."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import oracle
    oracle.build()
    oracle.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host core
    n, d, k = args.n, args.d, args.k
    t0 = time.time()
    x = to_bf16_values(gen_rows_numpy(0, n, d, 0))
    q_pool = to_bf16_values(gen_rows_numpy(0, min(8192, args.nq), d, 1))
    gen_s = time.time() - t0
    # every step is a bounded sample of the workload: ~10 s of all host cores, so W + K steps end within a few minutes
    sample = pick_cpu_sample(oracle, x, q_pool, k, args.cpu_sample, 10.0)
    q = np.ascontiguousarray(q_pool[:sample])
    for _ in range(args.warmup):
        oracle.knn_blocked(x, q, k, oracle.IP)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.knn_blocked(x, q, k, oracle.IP)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = sample / dt
    cores = oracle.num_threads()
    line = {
        "impl": "reference", "metric": "sem_sim_join queries/sec (1M x 768, K=32)", "value": val, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (bf16 values)", "data": "synthetic",
        "config": {"workload": f"sem_sim_join {args.nq} queries x {n} index, {d}-d bf16, K={k}", "nq": args.nq, "n": n, "d": d,
                   "k": k, "sample_queries_per_step": sample, "datagen_s": round(gen_s, 1)},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} queries x full {n}-row index per step (oracle/faiss_flat.c orc_knn_blocked, OpenMP)"},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--nq", type=int, default=100_000)
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries per CPU-baseline step (0 = about 10-15 s of the host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from lotus_b200 import _native as nv
    from lotus_b200.distributed import ShardedIndex, shard_bounds

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=device)
    nv.require_device()

    nq, n, d, k = args.nq, args.n, args.d, args.k
    lo, hi = shard_bounds(n, world, rank)
    corpus = gen_rows_torch(torch, lo, hi, d, 0, device, torch.bfloat16)
    queries = gen_rows_torch(torch, 0, nq, d, 1, device, torch.bfloat16)
    index = ShardedIndex(corpus, lo, nv.METRIC_IP)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item()) / steps

    filt_ms: list[float] = []

    def step_dev():
        index.search(queries, k)
        filt_ms.append(index.last_filter_ms())

    # ---- device-resident throughput --------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_dev()
    filt_ms.clear()
    nv.stats_reset()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step = timed(step_dev, args.steps)
    clocks = sampler.stop() if sampler else None
    st = nv.stats()
    launches_per_step = st["launches"] / max(args.steps, 1) + (2 if world > 1 else 0)  # + all-gather kernels are NCCL's
    fallback_q = st["fallback_queries"]
    kernel_ms = float(np.mean(filt_ms)) if filt_ms else float("nan")
    kernel_ms_ranks = [kernel_ms]
    if world > 1:  # the all-gather makes every step wait for the slowest rank: report the spread
        t = torch.tensor([kernel_ms], device=device, dtype=torch.float64)
        allk = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allk, t)
        kernel_ms_ranks = [float(a.item()) for a in allk]

    # ---- end to end: host (pinned) queries in, host results out, every step ------------------------------------------
    q_host = queries.cpu().pin_memory()
    out_s_host = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    q_stage = torch.empty_like(queries)

    def step_e2e():
        q_stage.copy_(q_host, non_blocking=True)
        s, i = index.search(q_stage, k)
        if rank == 0:
            out_s_host.copy_(s, non_blocking=True)
            out_i_host.copy_(i, non_blocking=True)

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    h2d = q_host.numel() * q_host.element_size()
    d2h = out_s_host.numel() * 4 + out_i_host.numel() * 8

    # ---- informational: the plain C-ABI host call of the plugin (b2_index_search: pageable numpy in, numpy out) --------
    e2e_plugin = None
    if world == 1:
        try:
            qb = q_host.view(torch.int16).numpy().view(np.uint16)  # bf16 bit patterns, as B200VS hands them over
            index.index.search(qb[:1024], k, nv.BF16)
            index.index.search(qb, k, nv.BF16)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                index.index.search(qb, k, nv.BF16)  # returns when scores + ids are in host memory
            dt = (time.perf_counter() - t0) / args.steps
            e2e_plugin = {"value": nq / dt, "unit": "queries/s", "ms_per_step": dt * 1e3, "clock": "host wall clock",
                          "api": "b2_index_search (host buffers in and out, the call B200VS.__call__ makes)"}
        except Exception as exc:  # never lose the bench line over the informational leg
            e2e_plugin = {"error": repr(exc)[:200]}

    # ---- roofline of the dominant kernel (the tcgen05 filter), per rank-0 launch --------------------------------------
    peak_tf, peak_src = load_peaks()
    flops_launch = 2.0 * nq * (hi - lo) * d  # algorithmic: 2*N_local*d per query
    achieved_tf = flops_launch / (kernel_ms * 1e-3) / 1e12 if kernel_ms == kernel_ms and kernel_ms > 0 else None
    traffic = None
    try:  # DRAM bytes of one launch of this kernel on this shape, from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "filter_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("nq") == nq and tj.get("n_local") == hi - lo and tj.get("d") == d:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        traffic = None
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                "kernel": "knn_filter_kernel<KP=64,IP,bf16,cta_group::2>", "kernel_ms": kernel_ms, "peak_source": peak_src,
                "kernel_ms_per_rank": [round(k, 3) for k in kernel_ms_ranks],
                "algorithmic_flops_per_launch": flops_launch,
                "hbm_floor_ms": ((hi - lo) * d * 2 + nq * d * 2 + nq * k * 12) / 6.4843e12 * 1e3}

    # ---- CPU baseline + parity sample (rank 0, single GPU run only) ---------------------------------------------------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        oracle.use_all_cores()
        xs = corpus.float().cpu().numpy()
        q_pool = queries[:min(8192, nq)].float().cpu().numpy()
        sample = pick_cpu_sample(oracle, xs, q_pool, k, args.cpu_sample, 15.0)
        qs = np.ascontiguousarray(q_pool[:sample])
        oracle.knn_blocked(xs[:50_000], qs[:64], k, oracle.IP)  # warm the threads
        t0 = time.perf_counter()
        Dc, Ic = oracle.knn_blocked(xs, qs, k, oracle.IP)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": sample / dt, "unit": "queries/s", "cores": oracle.num_threads(), "kind": "port",
                        "sample": f"{sample} of the {nq} queries x full {n}-row index, once ({dt:.1f} s; oracle/faiss_flat.c "
                                  "orc_knn_blocked, OpenMP fp32 FMA)"}
        s_dev, i_dev = index.search(queries[:sample].contiguous(), k)
        Ig = i_dev.cpu().numpy()
        Dg = s_dev.cpu().numpy()
        ncan = min(16, sample)
        Do, Io = oracle.knn(xs, qs[:ncan], k, oracle.IP)  # canonical oracle on a few queries
        recall = float(np.mean([len(set(a) & set(b)) / k for a, b in zip(Ig, Ic)]))
        parity = {"recall_at_k_vs_cpu_flat": recall, "sample_queries": sample,
                  "canonical_oracle_queries": ncan,
                  "idx_bit_exact_vs_oracle": bool(np.array_equal(Ig[:ncan], Io)),
                  "score_bit_exact_vs_oracle": bool(np.array_equal(Dg[:ncan].view(np.uint32), Do.view(np.uint32)))}

    if rank == 0:
        line = {
            "metric": "sem_sim_join queries/sec (1M x 768, K=32)", "value": nq / (ms_step * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"sem_sim_join {nq} queries x {n} index, {d}-d bf16, K={k} (BASELINE.json configs[2])",
                       "nq": nq, "n": n, "d": d, "k": k, "parallelism": f"index row-sharded x{world}",
                       "l2_policy": "inputs exceed L2 (corpus shard %.0f MB + queries %.0f MB vs 126 MB L2)" %
                                    ((hi - lo) * d * 2 / 1e6, nq * d * 2 / 1e6)},
            "e2e": {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_plugin": e2e_plugin,
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "gpu_launches_per_step": launches_per_step,
            "fallback_queries": int(fallback_q),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    index.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

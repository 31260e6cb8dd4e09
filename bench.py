#!/usr/bin/env python
"""bench.py -- batch_search QPS at recall@10 >= 0.95 on the SIFT-1M-shaped config (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one batch_search pass over the 10 000 synthetic queries at the smallest ef whose
recall@10 against exact ground truth is >= 0.95. Prints ONE JSON line (rank 0).

ours      index built on the GPU through the C ABI (ab_index_fit); `value` = QPS with queries and
          results resident in HBM (ab_index_batch_search_device, CUDA events); `e2e` = the same
          search through the reference-facing call ab_index_batch_search with pinned HOST buffers
          (H2D of the queries and D2H of the ids inside the timed region).
reference the reference's own CPU path (oracle/_ref: unmodified headers compiled in place; the exact
          PyIndex::batch_search route = coroutine Scheduler on every host thread) on the same data,
          same metric. Falls back to the oracle port if oracle/_ref is absent.
N > 1     one process per GPU (torchrun). Every rank owns one 1M-vector shard (ids offset by
          rank * 1M), searches all queries on it, and the per-shard top-k lists are all-gathered over
          NCCL and merged on the device. Weak scaling: the indexed set grows with N.
          `value` counts (query, shard) searches per second -- the unit every rank processes -- and
          `merged_qps` is the end-user query rate over the N x 1M union.
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

METRIC = "batch_search QPS at recall@10>=0.95"
WORKLOAD = "sift1m-shape: synthetic low-rank(L=16) 1M x 128 fp32, L2, HNSW R=32 efc=100, 10k queries, top-10"
N_BASE, DIM, NQ, LATENT, TOPK, EFC = 1_000_000, 128, 10_000, 16, 10, 100
EF_SWEEP = (10, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 60, 64, 72, 80, 100, 128, 160, 200, 300, 400, 800)
TARGET_RECALL = 0.95


def recall_at_k(res, gt):
    hits = 0
    for r, g in zip(res, gt):
        hits += len(set(r.tolist()) & set(g.tolist()))
    return hits / float(res.shape[0] * res.shape[1])


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """dram bytes per launch of the traversal kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("total")
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path, all host threads, same data and metric."""
    if rank != 0:
        return
    from alayalite_b200.synthetic import low_rank
    from oracle import oracle as O

    threads = os.cpu_count() or 1
    kind = "reference" if O.have_ref() else "port"
    if kind == "port" and not os.path.exists(O.ORC_SO):
        O.build(ref=False)
    chk = O.Checker("ref" if kind == "reference" else "orc")
    base, q = low_rank(N_BASE, DIM, NQ, LATENT, seed=42)
    t0 = time.time()
    if kind == "reference":
        idx = chk.index("l2", "none", N_BASE, DIM)
        idx.fit(base.copy(), EFC, threads)
    else:  # the scalar port builds single-threaded: bound it to a 100k sample, labelled as such
        base = base[:100_000]
        idx = chk.index("l2", "none", base.shape[0], DIM)
        idx.fit(base.copy(), EFC)
    build_s = time.time() - t0
    gt = chk.find_exact_gt(q[:500], base, TOPK, threads)  # reference's own exhaustive search, bounded sample
    ef = None
    for e in EF_SWEEP:
        r = recall_at_k(idx.batch_search_inplace(q[:500], TOPK, e, threads), gt)
        if r >= TARGET_RECALL:
            ef, rec = e, r
            break
    if ef is None:
        ef, rec = EF_SWEEP[-1], r
    qc = np.ascontiguousarray(q)
    for _ in range(args.warmup):
        idx.batch_search_inplace(qc, TOPK, ef, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        idx.batch_search_inplace(qc, TOPK, ef, threads)
    dt = time.perf_counter() - t0
    qps = NQ * args.steps / dt
    t0 = time.perf_counter()
    idx.batch_search_inplace(qc, TOPK, ef, threads, solo=True)
    solo_qps = NQ / (time.perf_counter() - t0)
    sample = f"all {NQ} queries per step on {base.shape[0]} x {DIM}; index built by the reference on {threads} threads in {build_s:.1f}s; recall on 500 queries"
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "ef": ef, "recall_at_10": rec, "path": "PyIndex::batch_search (Scheduler + coroutines) via oracle/_ref" if kind == "reference" else "oracle port",
                   "simd": chk.simd_level(), "search_solo_threads_qps": solo_qps},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(base, q, gt, threads):
    """Reference CPU path on the box's host cores, bounded: one index build + a few query passes."""
    from oracle import oracle as O

    kind = "reference" if O.have_ref() else "port"
    if kind == "port" and not os.path.exists(O.ORC_SO):
        O.build(ref=False)
    chk = O.Checker("ref" if kind == "reference" else "orc")
    t0 = time.time()
    if kind == "reference":
        idx = chk.index("l2", "none", base.shape[0], DIM)
        idx.fit(base.copy(), EFC, threads)
    else:
        base = base[:100_000]
        idx = chk.index("l2", "none", base.shape[0], DIM)
        idx.fit(base.copy(), EFC)
        gt = None
    build_s = time.time() - t0
    ef, rec = EF_SWEEP[-1], None
    if gt is not None:
        for e in EF_SWEEP:
            rec = recall_at_k(idx.batch_search_inplace(q[:1000], TOPK, e, threads), gt[:1000])
            if rec >= TARGET_RECALL:
                ef = e
                break
    else:
        ef = 50
    qc = np.ascontiguousarray(q)
    idx.batch_search_inplace(qc, TOPK, ef, threads)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        idx.batch_search_inplace(qc, TOPK, ef, threads)
    qps = NQ * reps / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    idx.batch_search_inplace(qc, TOPK, ef, threads, solo=True)
    solo = NQ / (time.perf_counter() - t0)
    return {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind,
            "sample": f"{reps} passes of {NQ} queries at ef={ef} (recall@10={rec}) on {base.shape[0]} x {DIM}; reference index built on {threads} threads in {build_s:.1f}s; "
                      f"PyIndex::batch_search route (Scheduler+coroutines); std::thread x search_solo gives {solo:.0f} q/s; simd {chk.simd_level()}"}


# --------------------------------------------------------------------------------------------------
def ours_arm(args, rank, world, local_rank):
    import torch

    from alayalite_b200 import _cabi
    from alayalite_b200.synthetic import low_rank

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _cabi.check(_cabi.lib().ab_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    # ---- data: every rank owns one shard (same generator, independent stream per rank); queries are shared ---------------------
    base, q = low_rank(N_BASE, DIM, NQ, LATENT, seed=42, shard=rank)  # same distribution and queries on every rank
    qt = torch.from_numpy(q).pin_memory()
    d_q = qt.to(dev)
    if world > 1:
        dist.broadcast(d_q, 0)
        qt.copy_(d_q)
        q = qt.numpy()
    id_offset = rank * N_BASE

    # ---- build (untimed part of the step, reported) ----------------------------------------------
    ix = _cabi.NativeIndex("l2", "none", N_BASE, device=local_rank)
    t0 = time.time()
    cache = os.path.join(args.index_cache, f"rank{rank}") if args.index_cache else None
    if cache and os.path.exists(os.path.join(cache, "hnsw.index")):
        # profiling aid: reuse an index this same script built and saved (reference file format)
        ix.load(os.path.join(cache, "hnsw.index"), os.path.join(cache, "raw.data"))
        built = "loaded from --index-cache"
    else:
        ix.fit(base, EFC)
        built = "ab_index_fit on the GPU"
        if cache:
            os.makedirs(cache, exist_ok=True)
            ix.save(os.path.join(cache, "hnsw.index"), os.path.join(cache, "raw.data"))
    build_s = time.time() - t0
    # a real (non-null) stream shared by torch's events/collectives and the library's kernels
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    from alayalite_b200.sharded import ShardedIndex

    sh = ShardedIndex(ix, id_offset, world)
    buf = sh._buffers(NQ, TOPK, dev)
    d_ids, d_dist = buf["ids"], buf["dist"]

    # ---- exact ground truth on the GPU (brute force per shard, merged across shards) --------------
    d_base = torch.from_numpy(base).to(dev)
    _cabi.bruteforce_topk_device("l2", d_base.data_ptr(), N_BASE, DIM, d_q.data_ptr(), NQ, TOPK, d_ids.data_ptr(), d_dist.data_ptr(), stream)
    d_ids += id_offset
    gt = sh.exchange(d_ids, d_dist, buf, stream)[0].cpu().numpy().copy()
    del d_base
    torch.cuda.empty_cache()

    # ---- operating point: smallest ef with recall@10 >= 0.95 -------------------------------------
    ef, rec = EF_SWEEP[-1], 0.0
    for e in EF_SWEEP:
        rec = recall_at_k(sh.batch_search_device(d_q, TOPK, e, stream)[0].cpu().numpy(), gt)
        if rec >= TARGET_RECALL:
            ef = e
            break

    # ---- timed region: K steps, device events, L2 flushed between steps ---------------------------
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    for _ in range(max(args.warmup, 3)):
        sh.batch_search_device(d_q, TOPK, ef, stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _cabi.kernel_launch_count()
    times, kernel_times = [], []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        for _ in range(args.steps):
            flush.zero_()
            ev[0].record()
            if world == 1:
                ix.batch_search_device(d_q.data_ptr(), NQ, TOPK, ef, d_ids.data_ptr(), d_dist.data_ptr(), True, stream)
                ev[1].record()
            else:  # search (storing into every rank's gather buffer) + barrier + merge
                sh.batch_search_device(d_q, TOPK, ef, stream, after_search=ev[1].record)
            ev[2].record()
            torch.cuda.synchronize()
            kernel_times.append(ev[0].elapsed_time(ev[1]))
            times.append(ev[0].elapsed_time(ev[2]))
    launches = _cabi.kernel_launch_count() - launches0
    if world > 1:
        dist.barrier()
    total_ms = float(sum(times))
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    stats = ix.stats()

    # ---- e2e: host buffers in, host ids out, copies inside the timed region ------------------------
    # N = 1: the reference-facing C-ABI call ab_index_batch_search on pinned host memory.
    # N > 1: ShardedIndex.batch_search (H2D queries, search, all-gather + merge, D2H merged ids).
    h_ids = torch.empty((NQ, TOPK), dtype=torch.int32).pin_memory()

    def e2e_step():
        if world == 1:
            ix.batch_search_ptr(qt.data_ptr(), NQ, TOPK, ef, h_ids.data_ptr())
        else:
            sh.batch_search(qt, TOPK, ef, out=h_ids, device=dev, stream=stream)

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_recall = recall_at_k(h_ids.numpy(), gt)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (beam_search_kernel): algorithmic bytes / device time -----
    row_bytes = DIM * 4
    alg_bytes = (stats["ndc"] + stats["overlay"]) * row_bytes + stats["hops"] * 128  # per launch
    k_ms = float(np.mean(kernel_times))
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = load_traffic()
    units = NQ * world * args.steps
    value = units / (total_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "ef": ef, "recall_at_10": rec, "vectors_total": N_BASE * world, "shards": world, "exchange": sh.exchange_kind,
                   "timing": "CUDA events per step on the launching stream; 256 MB L2 flush between steps; index 640 MB > 126 MB L2",
                   "build_s": build_s, "index": built, "unit_note": "value counts (query, shard) searches; merged_qps is the end-user rate" if world > 1 else "1 shard: value == merged_qps"},
        "merged_qps": NQ * args.steps / (total_ms * 1e-3),
        "e2e": {"value": NQ * world * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": NQ * DIM * 4, "d2h_bytes_per_step": NQ * TOPK * 4,
                "api": "ab_index_batch_search (C ABI, pinned host buffers)" if world == 1 else "ShardedIndex.batch_search (pinned host buffers)", "recall_at_10": e2e_recall},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "beam_search_kernel", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "per_query": {"ndc_level0": stats["ndc"] / NQ, "ndc_overlay": stats["overlay"] / NQ, "hops": stats["hops"] / NQ, "bytes": alg_bytes / NQ}},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(base, q, gt, os.cpu_count() or 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--index-cache", default="", help="directory to save/load the built index (profiling aid)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

"""bench.py, N > 1: one process per GPU (torchrun), a grid of id-range shards x query replicas.

Every number here is END-USER queries per second: a query is counted once, however many shards looked at it.

main line   the north-star 10M x 128 index in the layout --layout SxR (default 1xN: the index fits one GPU, so each
            GPU holds a replica and answers its own 10k-query batch -- the layout that scales QPS; "weak" scaling:
            per-GPU work is fixed, the global batch grows with N).
`sharded`   the same 10M rows split by contiguous id ranges over N shards (strong scaling: total rows and the batch
            are fixed). Every shard answers every query; the exchange is fused into the search kernel (peer stores +
            last-arriver merge). Reports the step, the search alone and the exchange share, and asserts in the run that
            fused == NCCL all-gather + merge kernel == exact merge of the per-shard lists.
`c5`        BASELINE configs[4]: 100M x 96 split over the N shards (50M rows per GPU at N = 2), same measurements.
"""

from __future__ import annotations

import json
import time

import numpy as np

from bench import CONFIGS, EF_SWEEP, EFC, METRIC, ClockSampler, algorithmic_bytes, load_peaks


def _hits(res, gt):
    h = 0
    for r, g in zip(res.astype(np.int64), gt.astype(np.int64)):
        h += len(set(r.tolist()) & set(g.tolist()))
    return h


def run_grid(cfg, S, R, args, torch, dist, dev, local_rank, stream, rank, world, groups, check_parity):
    from alayalite_b200 import _cabi
    from alayalite_b200.sharded import ShardedIndex, shard_range
    from alayalite_b200.synthetic import BLOCK_ROWS, low_rank_device

    n, dim, nq, k = cfg["n"], cfg["dim"], cfg["nq"], cfg["topk"]
    j, s = rank // S, rank % S  # replica group, shard within it
    group = groups[(S, R)][j]
    r0, r1 = shard_range(n, S, s)
    assert r0 % BLOCK_ROWS == 0, "shard boundaries must fall on generator blocks"
    t0 = time.time()
    d_base, d_qall = low_rank_device(r1 - r0, dim, nq * R, cfg["latent"], seed=42, normalize=cfg.get("normalize", False), device=dev,
                                     first_block=r0 // BLOCK_ROWS)
    d_q = d_qall[j * nq:(j + 1) * nq].contiguous()  # every replica group answers its own batch
    del d_qall
    ix = _cabi.NativeIndex(cfg["metric"], cfg["quant"], r1 - r0, device=local_rank)
    ix.fit_ptr(d_base.data_ptr(), r1 - r0, dim, EFC)
    torch.cuda.synchronize()
    build_s = time.time() - t0

    fused = ShardedIndex(ix, r0, S, rank=s, group=group, max_nq=nq, max_topk=max(k, 16))
    coll = ShardedIndex(ix, r0, S, rank=s, group=group, peer_memory=False)
    own = fused.owned(nq)

    # ---- exact ground truth: per-shard brute force, merged over the replica group (collective route) ----
    b = coll._buffers(nq, k, dev)
    _cabi.bruteforce_topk_device(cfg["metric"], d_base.data_ptr(), r1 - r0, dim, d_q.data_ptr(), nq, k, b["ids"].data_ptr(), b["dist"].data_ptr(),
                                 stream.cuda_stream)
    b["ids"] += r0
    gt = coll.exchange(b["ids"], b["dist"], b, stream)[0].cpu().numpy().copy()
    del d_base
    torch.cuda.empty_cache()

    def recall_of(sh, ef):
        ids, _ = sh.batch_search_device(d_q, k, ef, stream)
        torch.cuda.synchronize()
        t = torch.tensor([_hits(ids.cpu().numpy(), gt[own])], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item()) / (nq * k * R)

    ef, rec = EF_SWEEP[-1], 0.0
    hist = []
    for e in EF_SWEEP:
        if e < k:
            continue
        rec = recall_of(fused, e)
        hist.append((e, round(rec, 4)))
        if rec >= cfg["target"]:
            ef = e
            break

    # ---- in-run parity of the exchange: fused == collective == exact merge of the per-shard lists --------
    parity = None
    if S > 1 and check_parity:
        a_ids, a_dist = [t.clone() for t in fused.batch_search_device(d_q, k, ef, stream)]
        c_ids, c_dist = [t.clone() for t in coll.batch_search_device(d_q, k, ef, stream)]
        loc_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
        loc_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
        ix.batch_search_device(d_q.data_ptr(), nq, k, ef, loc_i.data_ptr(), loc_d.data_ptr(), 3, stream.cuda_stream)
        loc_i.add_(r0 * (loc_i >= 0).to(torch.int32))
        gi = [torch.empty_like(loc_i) for _ in range(S)]
        gd = [torch.empty_like(loc_d) for _ in range(S)]
        dist.all_gather(gi, loc_i, group=group)
        dist.all_gather(gd, loc_d, group=group)
        torch.cuda.synchronize()
        ci = torch.cat(gi, 1).cpu().numpy().view(np.uint32).astype(np.int64)  # unfilled slots (0xFFFFFFFF, FLT_MAX) sort last
        cd = torch.cat(gd, 1).cpu().numpy()
        order = np.lexsort((ci, cd), axis=1)[:, :k]
        want = np.take_along_axis(ci, order, 1)[own]
        ok = (torch.equal(a_ids, c_ids) and torch.equal(a_dist, c_dist)
              and np.array_equal(a_ids.cpu().numpy().view(np.uint32).astype(np.int64), want)
              and (fused._comm is not None and fused._comm.errors() == 0))
        t = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        parity = bool(t.item())

    # ---- timed region ---------------------------------------------------------------------------------------
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        fused.batch_search_device(d_q, k, ef, stream)
    torch.cuda.synchronize()
    dist.barrier()
    l0 = _cabi.kernel_launch_count()
    # K steps enqueued back to back between ONE barrier + synchronize on either side (the bench contract); every step
    # has its own event pair behind the L2 flush, so the flush is not timed and no host round trip sits between
    # steps -- per-step host-launch skew between the ranks would otherwise be booked as "exchange".
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    clocks = ClockSampler(local_rank)
    with clocks:
        dist.barrier()
        torch.cuda.synchronize()
        for i in range(args.steps):
            flush.zero_()
            ev0[i].record()
            fused.batch_search_device(d_q, k, ef, stream)
            ev1[i].record()
        torch.cuda.synchronize()
        dist.barrier()
    times = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    launches = _cabi.kernel_launch_count() - l0
    stats = ix.stats()
    # the local search alone (no exchange, no wait): what the exchange adds is the difference
    d_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    for _ in range(warm):
        ix.batch_search_device(d_q.data_ptr(), nq, k, ef, d_i.data_ptr(), d_d.data_ptr(), True, stream.cuda_stream)
    dist.barrier()
    torch.cuda.synchronize()
    for i in range(args.steps):
        flush.zero_()
        ev0[i].record()
        ix.batch_search_device(d_q.data_ptr(), nq, k, ef, d_i.data_ptr(), d_d.data_ptr(), True, stream.cuda_stream)
        ev1[i].record()
    torch.cuda.synchronize()
    alone = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    del flush

    # ---- end to end: pinned host queries in, owned merged rows out on the host, copies inside the timed region ----
    qh = d_q.cpu().pin_memory()
    oh = torch.empty((max(len(own), 1), k), dtype=torch.int32).pin_memory()
    od = torch.empty((max(len(own), 1), k), dtype=torch.float32).pin_memory()
    d_q2 = torch.empty_like(d_q)

    def e2e_step():
        d_q2.copy_(qh, non_blocking=True)
        fused.batch_search_device(d_q2, k, ef, stream, out_ids=oh, out_dist=od)
        stream.synchronize()

    for _ in range(2):
        e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    e2e_hits = _hits(oh.numpy()[: len(own)], gt[own])

    red = torch.tensor([sum(times), sum(alone), e2e_s], dtype=torch.float64, device=dev)
    dist.all_reduce(red, op=dist.ReduceOp.MAX)
    h = torch.tensor([e2e_hits], dtype=torch.float64, device=dev)
    dist.all_reduce(h)
    total_ms, alone_ms, e2e_s = float(red[0]), float(red[1]), float(red[2])
    hbm, _, peak_src = load_peaks()
    k_ms = float(np.mean(alone))
    alg = algorithmic_bytes(cfg, stats, nq)
    ach = alg / (k_ms * 1e-3) / 1e9
    step_ms = total_ms / args.steps
    out = {
        "workload": cfg["workload"], "layout": f"{S} id-range shard(s) x {R} query replica(s)", "shards": S, "replicas": R,
        "rows_per_gpu": r1 - r0, "queries_per_step": nq * R,
        "ef": ef, "recall": rec, "recall_k": k, "recall_by_ef": hist[-5:],
        "value": nq * R * args.steps / (total_ms * 1e-3), "unit": "queries/s", "ms_per_step": step_ms,
        "search_alone_ms": alone_ms / args.steps, "exchange_ms": max(0.0, step_ms - alone_ms / args.steps),
        "exchange_frac": max(0.0, 1.0 - alone_ms / total_ms) if S > 1 else 0.0,
        "exchange": fused.exchange_kind if S > 1 else "none (replicas only)", "exchange_parity": parity,
        "e2e": {"value": nq * R * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": len(own) * k * 8,
                "api": "ShardedIndex.batch_search_device on pinned host buffers: H2D of the rank's query batch, search + fused exchange, owned merged rows written to pinned host memory",
                "recall": float(h.item()) / (nq * k * R)},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None, "kernel": "beam_search_kernel",
                     "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src, "note": "rank 0's shard, search alone"},
        "gpu_launches": int(launches), "build_s": build_s, "clocks": clocks.summary(),
    }
    del fused, coll, ix
    torch.cuda.empty_cache()
    return out


def ours_multi(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from alayalite_b200 import _cabi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _cabi.check(_cabi.lib().ab_set_device(local_rank))
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    t_start = time.time()

    if args.layout:
        S, R = [int(x) for x in args.layout.lower().split("x")]
        assert S * R == world, "--layout SxR must multiply to the number of GPUs"
    else:
        S, R = 1, world
    layouts = {(S, R)}
    blocks = [b for b in args.blocks.split(",") if b]
    want_sharded = (S, R) != (world, 1) and "nosharded" not in blocks
    want_c5 = world >= 2 and "noc5" not in blocks
    if want_sharded or want_c5:
        layouts.add((world, 1))
    # process groups of every layout (new_group is collective: all ranks create all groups in the same order)
    groups = {}
    for (s_, r_) in sorted(layouts):
        groups[(s_, r_)] = [dist.new_group(list(range(j * s_, (j + 1) * s_))) if s_ > 1 else None for j in range(r_)]

    cfg = CONFIGS[args.config]
    main = run_grid(cfg, S, R, args, torch, dist, dev, local_rank, stream, rank, world, groups, check_parity=True)
    extra = {}
    if want_sharded:
        try:
            extra["sharded"] = run_grid(cfg, world, 1, args, torch, dist, dev, local_rank, stream, rank, world, groups, check_parity=True)
        except Exception as e:  # noqa: BLE001
            extra["sharded"] = {"error": repr(e)[:300]}
    if want_c5:
        try:
            extra["c5"] = run_grid(CONFIGS["c5"], world, 1, args, torch, dist, dev, local_rank, stream, rank, world, groups, check_parity=True)
        except Exception as e:  # noqa: BLE001
            extra["c5"] = {"error": repr(e)[:300]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "layout": main["layout"], "ef": main["ef"], "recall_at_10": main["recall"], "vectors_total": cfg["n"],
                       "shards": S, "replicas": R, "queries_per_step": main["queries_per_step"],
                       "unit_note": "value = end-user queries/s: every query counted once (each replica group answers its own 10k-query batch over the whole index)",
                       "timing": "K steps enqueued back to back between one barrier + synchronize on both sides; a CUDA event pair per step on the launching stream behind a 256 MB L2 flush (not timed); max over ranks of the summed step times",
                       "build_s": main["build_s"], "exchange": main["exchange"], "exchange_parity": main["exchange_parity"]},
            "merged_qps": main["value"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"], "roofline": main["roofline"],
            "exchange_frac": main["exchange_frac"],
        }
        line.update(extra)
        line["bench_s"] = time.time() - t_start
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()

"""Seeded synthetic datasets shaped like BASELINE.json's configs.

I.i.d. uniform data is pathologically hard for graph search (distance concentration), so the graph
configs use a low-rank latent model  x = z A + 0.05 eps,  z ~ N(0, I_L),  A ~ N(0, 1/L)^(L x d),
which the survey calibrated to SIFT-like difficulty (L=16 at d=128).
"""

from __future__ import annotations

import numpy as np

__all__ = ["low_rank", "low_rank_device", "uniform", "integer_grid", "CONFIGS", "BLOCK_ROWS"]

# name -> (n, dim, nq, latent L, metric, quant, topk)
CONFIGS = {
    "quickstart": (1000, 128, 10, 0, "l2", "none", 10),
    "sift1m": (1_000_000, 128, 10_000, 16, "l2", "none", 10),
    "gist960": (1_000_000, 960, 10_000, 48, "l2", "sq8", 10),
    "rag768": (5_000_000, 768, 10_000, 48, "ip", "sq4", 100),
    "deep100m": (100_000_000, 96, 10_000, 24, "l2", "none", 10),
}


def low_rank(n: int, dim: int, nq: int, latent: int = 16, seed: int = 42, noise: float = 0.05, normalize: bool = False, chunk: int = 1 << 18,
             shard: int = 0):
    """Base [n, dim] and queries [nq, dim], float32, from one seeded stream (A, then base, then queries).
    shard > 0 draws another n rows of the SAME distribution (same mixing matrix A) from an independent
    stream: the rows of shard `shard` of a larger sharded index. Queries always come from the seed's
    own stream, so every shard sees the same queries."""
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal((latent, dim)) / np.sqrt(latent)).astype(np.float32)

    def gen(m, r):
        out = np.empty((m, dim), np.float32)
        for s in range(0, m, chunk):
            e = min(m, s + chunk)
            z = r.standard_normal((e - s, latent), dtype=np.float32)
            out[s:e] = z @ a
            out[s:e] += noise * r.standard_normal((e - s, dim), dtype=np.float32)
        if normalize:
            out /= np.linalg.norm(out, axis=1, keepdims=True)
        return out

    if shard == 0:
        return gen(n, rng), gen(nq, rng)
    base = gen(n, np.random.default_rng([seed, shard]))
    gen(n, rng)  # advance the seed's stream past shard 0's rows so the queries are the shared ones
    return base, gen(nq, rng)


def uniform(n: int, dim: int, nq: int, seed: int = 0):
    """README quickstart shape: np.random.rand-like data (python/tests/test_index_types.py:32-40)."""
    rng = np.random.default_rng(seed)
    return rng.random((n, dim), dtype=np.float32), rng.random((nq, dim), dtype=np.float32)


def integer_grid(n: int, dim: int, nq: int, seed: int = 7, hi: int = 16):
    """Small-integer coordinates stored as float32: every fp32 sum on the path is exact, so results
    are independent of summation order and must match the reference bit for bit, ties included."""
    rng = np.random.default_rng(seed)
    return (rng.integers(0, hi, (n, dim)).astype(np.float32), rng.integers(0, hi, (nq, dim)).astype(np.float32))


BLOCK_ROWS = 250_000


def low_rank_device(n: int, dim: int, nq: int, latent: int = 16, seed: int = 42, noise: float = 0.05, normalize: bool = False,
                    device="cuda", first_block: int = 0):
    """The same latent model generated on the GPU (torch Philox streams; plumbing, not a kernel of the path) for the
    benchmark's large configurations, where numpy on the host would take longer than the measurement.

    The base set is a sequence of BLOCK_ROWS-row blocks, block b drawn from its own stream seeded (seed, b), so a shard
    that owns rows [first_block * BLOCK_ROWS, first_block * BLOCK_ROWS + n) of a larger set generates exactly those
    rows, whatever the number of shards. Queries come from the stream (seed, 'q') and are the same for everybody.
    Returns (base [n, dim], queries [nq, dim]) as float32 tensors on `device`. Reproducible for a given GPU model and
    torch build -- which is all the two arms of bench.py, run on the same box, need."""
    import math

    import torch

    dev = torch.device(device)

    def stream(tag):
        g = torch.Generator(device=dev)
        g.manual_seed((seed * 1_000_003 + tag) & 0x7FFFFFFFFFFFFFFF)
        return g

    a = torch.randn((latent, dim), generator=stream(1 << 40), device=dev) / math.sqrt(latent)

    def draw(m, g, out):
        z = torch.randn((m, latent), generator=g, device=dev)
        torch.matmul(z, a, out=out)
        out.add_(torch.randn((m, dim), generator=g, device=dev), alpha=noise)
        if normalize:
            out.div_(out.norm(dim=1, keepdim=True))

    base = torch.empty((n, dim), dtype=torch.float32, device=dev)
    done, b = 0, first_block
    while done < n:
        g = stream(b)
        m = min(BLOCK_ROWS, n - done)
        if m == BLOCK_ROWS:
            draw(m, g, base[done:done + m])
        else:  # a partial last block is the head of the full block (keeps rows independent of n)
            tmp = torch.empty((BLOCK_ROWS, dim), dtype=torch.float32, device=dev)
            draw(BLOCK_ROWS, g, tmp)
            base[done:done + m] = tmp[:m]
        done += m
        b += 1
    q = torch.empty((nq, dim), dtype=torch.float32, device=dev)
    draw(nq, stream((1 << 40) + 1), q)
    return base, q

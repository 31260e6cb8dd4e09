"""Generates tests/golden/ref_small.npz from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile). Run in the build container, where the reference is mounted:

    python tests/golden/make_golden.py

The fixtures travel to the GPU box, where /root/reference does not exist. One-thread builds are
deterministic in the reference, so the graphs are reproducible."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from alayalite_b200.synthetic import low_rank  # noqa: E402
from oracle.oracle import Checker  # noqa: E402

CONFIGS = [("l2", "none"), ("ip", "none"), ("cos", "none"), ("l2", "sq8"), ("ip", "sq8"), ("l2", "sq4"), ("cos", "sq4")]
SEARCHES = [(10, 32), (5, 50), (10, 10)]  # (topk, ef)


def main():
    ref = Checker("ref")
    base, q = low_rank(600, 24, 40, latent=6, seed=11)
    out = {"base": base, "queries": q, "simd": np.array(ref.simd_level())}
    for metric, quant in CONFIGS:
        tag = f"{metric}_{quant}"
        ix = ref.index(metric, quant, 640, 24)
        ix.fit(base.copy(), 40, 1)
        g = ix.graph()
        out[f"{tag}/adj"], out[f"{tag}/levels"], out[f"{tag}/lists"], out[f"{tag}/ep"] = g["adj"], g["levels"], g["lists_flat"], np.array(g["ep"])
        out[f"{tag}/raw"] = ix.raw_rows()
        if quant != "none":
            mn, mx = ix.minmax()
            out[f"{tag}/min"], out[f"{tag}/max"], out[f"{tag}/codes"] = mn, mx, ix.codes()
        for k, ef in SEARCHES:
            out[f"{tag}/ids_k{k}_ef{ef}"] = ix.batch_search(q, k, ef, 2)
        ids, dist = ix.batch_search_with_distance(q, 10, 32, 2)
        out[f"{tag}/wd_ids"], out[f"{tag}/wd_dist"] = ids, dist
        pairs = [(i, (i * 7 + 3) % 600) for i in range(0, 600, 13)]
        out[f"{tag}/pairs"] = np.array(pairs, np.uint32)
        out[f"{tag}/pair_dist"] = np.array([ix.space_distance(i, j) for i, j in pairs], np.float32)
        # incremental update: three inserts then one remove, then search again
        ins = [ix.insert(q[i], 32) for i in range(3)]
        ix.remove(5)
        out[f"{tag}/insert_ids"] = np.array(ins, np.int64)
        out[f"{tag}/ids_after_update"] = ix.batch_search(q, 10, 32, 1)
        out[f"{tag}/adj_after_update"] = ix.graph()["adj"]
    out["gt_k10"] = ref.find_exact_gt(q, base, 10, 2)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

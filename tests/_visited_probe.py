"""Helper for test_visited_table_formats_agree: load a saved index and search it under whatever AB_* environment
the parent set (the knobs are read once per process).  python tests/_visited_probe.py folder out.npz ef..."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi  # noqa: E402

folder, out = sys.argv[1], sys.argv[2]
efs = [int(x) for x in sys.argv[3:]]
q = np.load(os.path.join(folder, "queries.npy"))
n = int(np.load(os.path.join(folder, "n.npy")))
ix = _cabi.NativeIndex("l2", "none", n)
ix.load(os.path.join(folder, "g.index"), os.path.join(folder, "g.data"))
res = {}
for ef in efs:
    ids, dist = ix.batch_search_with_distance(q, 10, ef)
    res[f"ids{ef}"] = ids
    res[f"dist{ef}"] = dist
    res[f"bitmap{ef}"] = np.array(ix.stats()["spilled"])  # queries that used their overflow table (level 2)
    res[f"big{ef}"] = np.array(ix.stats()["bitmap"])  # queries that went on to a pooled bitmap (level 3)
np.savez(out, **res)

"""Developer probe: pool distances of the packed SQ kernels against the oracle's distance function on one shape."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank
from oracle.oracle import Checker
seed, quant, metric, n, dim = int(sys.argv[1]), sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
orc = Checker("orc")
base, q = low_rank(n, dim, 48, latent=6, seed=seed)
o = orc.index(metric, quant, n + 64, dim); o.fit(base.copy(), 40); g = o.graph()
ix = _cabi.NativeIndex(metric, quant, n + 64); ix.fit_spaces(base); ix.set_graph(g["adj"], g["levels"], g["lists_flat"], g["ep"])
mn, mx, codes = ix.codes()
ocodes = o.codes()
print("codes equal", np.array_equal(codes, ocodes))
ids, dist, sizes = ix.traverse(q, 70)
name = ("l2_sqr" if metric == "l2" else "ip_sqr") + "_" + quant
bits = 8 if quant == "sq8" else 4
worst = 0.0; same = 0
for i in range(len(q)):
    pi, pd, st = o.traverse(q[i], 70)
    same += int(np.array_equal(pi, ids[i, :sizes[i]]))
    # oracle distance for the GPU's pool entries
    qc = np.array([orc.quantize(bits, q[i][j], mn[j], mx[j]) for j in range(dim)], dtype=np.uint8)
    if bits == 4:
        pad = np.zeros(dim + (dim & 1), np.uint8); pad[:dim] = qc
        qc = ((pad[0::2] << 4) | pad[1::2]).astype(np.uint8)
    for j in range(min(5, sizes[i])):
        od = orc.dist(name, qc, codes[ids[i, j]][: len(qc)], mn, mx)
        worst = max(worst, abs(od - dist[i, j]) / max(abs(od), 1e-9))
    if i < 3:
        print("gpu", ids[i, :6], dist[i, :6]); print("orc", pi[:6], pd[:6])
print("max rel err of pool distances vs oracle function:", worst, " identical pools:", same, "/", len(q))

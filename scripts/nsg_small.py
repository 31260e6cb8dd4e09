import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank
base, q = low_rank(int(sys.argv[1]) if len(sys.argv) > 1 else 3000, 48, 50, latent=10, seed=21)
ix = _cabi.NativeIndex("l2", "none", len(base))
ix.fit_spaces(base)
ix.build_graph_nsg(int(sys.argv[2]) if len(sys.argv) > 2 else 64, int(sys.argv[3]) if len(sys.argv) > 3 else 4)
print("ok", ix.batch_search(q, 10, 32)[0])

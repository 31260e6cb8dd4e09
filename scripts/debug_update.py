import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alayalite_b200 import _cabi
from oracle.oracle import Checker
GOLD = np.load("tests/golden/ref_small.npz")
orc = Checker("orc")
q = GOLD["queries"]; base = GOLD["base"]
for metric, quant in [("l2", "none"), ("l2", "sq8"), ("ip", "sq8"), ("l2", "sq4")]:
    tag = f"{metric}_{quant}"
    ix = _cabi.NativeIndex(metric, quant, 640)
    ix.fit_spaces(base)
    ix.set_graph(GOLD[f"{tag}/adj"], GOLD[f"{tag}/levels"], GOLD[f"{tag}/lists"], int(GOLD[f"{tag}/ep"]))
    for k, ef in ((10, 32), (5, 50), (10, 10)):
        got = ix.batch_search(q, k, ef)
        print(tag, "before", k, ef, np.mean(np.all(got == GOLD[f"{tag}/ids_k{k}_ef{ef}"], axis=1)))
    o = orc.index(metric, quant, 640, 24); o.fit(base.copy(), 40)
    ins = [ix.insert(q[i], 32) for i in range(3)]; ix.remove(5)
    ins2 = [o.insert(q[i], 32) for i in range(3)]; o.remove(5)
    a, b = ix.graph()["adj"], GOLD[f"{tag}/adj_after_update"]
    bad = np.argwhere(a != b)
    print(tag, "adj diffs", len(bad), bad[:10].tolist())
    for r in sorted(set(bad[:, 0].tolist()))[:4]:
        print("  row", r, "gpu", a[r].tolist(), "\n        ref", b[r].tolist())
    got = ix.batch_search(q, 10, 32)
    want = GOLD[f"{tag}/ids_after_update"]
    same = np.all(got == want, axis=1)
    print(tag, "after", same.mean(), "oracle-after", np.mean(np.all(o.batch_search(q, 10, 32) == want, axis=1)))
    for i in np.nonzero(~same)[0][:3]:
        print("  q", i, "gpu", got[i].tolist(), "ref", want[i].tolist())
        pi, pd, _ = o.traverse(q[i], 32)
        gi, gd, gs = ix.traverse(q[i:i+1], 32)
        print("   orc pool", pi[:12].tolist(), np.round(pd[:12], 5).tolist())
        print("   gpu pool", gi[0][:12].tolist(), np.round(gd[0][:12], 5).tolist())

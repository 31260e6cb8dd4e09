import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alayalite_b200 import _cabi
from alayalite_b200.synthetic import low_rank
base, q = low_rank(1_000_000, 128, 10_000, 16, seed=42)
d_base, d_q = torch.from_numpy(base).cuda(), torch.from_numpy(q).cuda()
ids = torch.empty((10000, 10), dtype=torch.int32, device="cuda"); dist = torch.empty((10000, 10), dtype=torch.float32, device="cuda")
os.environ["AB_BRUTEFORCE_TC"] = "1"
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    _cabi.bruteforce_topk_device("l2", d_base.data_ptr(), 1_000_000, 128, d_q.data_ptr(), 10_000, 10, ids.data_ptr(), dist.data_ptr())
    torch.cuda.synchronize(); print("tc ms", (time.time() - t) * 1e3)

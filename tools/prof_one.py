"""One layer through the device path a few times (ncu target).  python tools/prof_one.py r n [passes]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleekit_b200 import codebook, ops  # noqa: E402
from sleekit_b200 import workloads as wl  # noqa: E402
from sleekit_b200.pipeline import LayerSetQuantizer  # noqa: E402

r, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (768, 3072)
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cb = codebook.UniformCodebook(8, -1, 1)
W, H, _ = wl.synthetic_layer(r, n, 0, samples=2048)
Wd, Hd = torch.from_numpy(W).cuda(), torch.from_numpy(H).cuda()
lsq = LayerSetQuantizer(cb, scaling_mode="diag", act_order="diag", damp=0.01, streams=1)
for _ in range(passes):
    q, sc, e = lsq._one(Wd, Hd)     # the per-layer hot path of the bench: scale search -> GPTQ -> layer error
torch.cuda.synchronize()
print("ok", r, n, float(e))

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
import torch
from b200clip import ops
dev = torch.device("cuda:0")
NP, D = 28, 512
g = torch.Generator().manual_seed(1234)
P = torch.nn.functional.normalize(torch.randn(NP, D, generator=g), dim=1).to(torch.bfloat16).to(dev)
for N in (16, 1000, 100000, 1000000):
    X = torch.randn(N, D, generator=g).to(torch.bfloat16).to(dev)
    for guard in (1e-9, 0.0):
        for cg in (False, True):
            try:
                o = ops.zeroshot_score(X, P, pair_mode=True, temperature=0.07, thresholds=[0.5], count_guard=cg, guard=guard)
                torch.cuda.synchronize()
                print("ok", N, guard, cg, int(o["mask"].sum()), flush=True)
            except Exception as e:
                print("FAIL", N, guard, cg, str(e)[:80], flush=True); sys.exit(0)

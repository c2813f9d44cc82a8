import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
import torch
from b200clip import ops
dev = torch.device("cuda:0")
N, NP, D = 1_000_000, 28, 512
g = torch.Generator().manual_seed(1234)
X = torch.randn(N, D, generator=g).to(torch.bfloat16).to(dev)
P = torch.nn.functional.normalize(torch.randn(NP, D, generator=g), dim=1).to(torch.bfloat16).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for guard in (None, 0.0, 1e-5, 1e-4):
    print("start", guard, flush=True)
    kw = {} if guard is None else {"guard": guard}
    for _ in range(3):
        o = ops.zeroshot_score(X, P, pair_mode=True, temperature=0.07, thresholds=[0.5], count_guard=True, **kw)
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = ops.zeroshot_score(X, P, pair_mode=True, temperature=0.07, thresholds=[0.5], count_guard=True, **kw); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"guard={guard}: median {ts[4]*1e3:.1f} us, flagged rows {int(o['guard_rows'].item()) }")

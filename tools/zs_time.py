import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
from b200clip import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
N = 1_000_000
D = int(sys.argv[1]) if len(sys.argv) > 1 else 512
X = torch.randn(N, D, generator=g).to(torch.bfloat16).to(dev)
P = torch.nn.functional.normalize(torch.randn(28, D, generator=g), dim=1).to(torch.bfloat16).to(dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for guard in (None, 0.0):
    for _ in range(3):
        out = ops.zeroshot_score(X, P, pair_mode=True, temperature=0.07, thresholds=[0.5], guard=guard, count_guard=True)
    tot = 0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ops.zeroshot_score(X, P, pair_mode=True, temperature=0.07, thresholds=[0.5], guard=guard, count_guard=True); e1.record()
        torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    print("D", D, "GB/s", round(N * (D * 2 + 3) / (tot / 10) / 1e6, 1), "guard", guard, "ms", tot / 10, "guard_rows", int(out["guard_rows"]))
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for guard in (None, 0.0):
        flush.zero_()
        ops.zeroshot_score(X, P, pair_mode=True, temperature=0.07, thresholds=[0.5], guard=guard, count_guard=True)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f} us  {e.name[:80]}")

"""Stand-alone timing of the two InfoNCE kernels:  python tools/nce_time.py [B] [D] [b_loc]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
import torch
from b200clip import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
b_loc = int(sys.argv[3]) if len(sys.argv) > 3 else B
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
T = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(torch.bfloat16)
I = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(torch.bfloat16)[:b_loc].contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=8):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


lib = _lib.load()
nb = lib.b200clip_infonce_workspace_bytes(b_loc, B)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
r, c = torch.empty(b_loc, device=dev), torch.empty(B, device=dev)
fwd = lambda: _lib.check(lib.b200clip_infonce_fwd_stats(_lib.ptr(I), _lib.ptr(T), D, b_loc, B, 0.07, _lib.ptr(r), _lib.ptr(c), _lib.ptr(ws), nb,
                                                        _lib.stream_ptr()), "fwd")
med, best = timeit(fwd)
fl = 2.0 * b_loc * B * D
print(f"fwd (stats + reduce): median {med:.3f} ms, best {best:.3f} ms -> {fl / best / 1e9:.0f} TFLOP/s at B={B} b_loc={b_loc} D={D}")
if os.environ.get("NCE_TIME_BWD", "1") == "1":
    loss, rinvh, cinvh = ops.infonce_forward(I, T, 0.07)
    splits = int(lib.b200clip_infonce_bwd_splits(b_loc, B)) if b_loc < B else 1
    bwd = lambda: ops.infonce_backward(I, T, 0.07, rinvh, cinvh, None, allow_splits=splits > 1)
    med, best = timeit(bwd)
    print(f"bwd: median {med:.3f} ms, best {best:.3f} ms -> algorithmic {2 * fl / best / 1e9:.0f} TFLOP/s, executed {4 * fl / best / 1e9:.0f} TFLOP/s"
          f"  (loss {float(loss):.6f})")

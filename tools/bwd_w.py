"""Rank-shaped InfoNCE backward on ONE GPU (the shapes a rank sees at world size W: b_loc = B / W rows of a B-row batch):
dT alone, dI alone (per split count), and both concurrently on two streams -- python tools/bwd_w.py [W] [B] [D]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
import torch
from b200clip import _lib, ops

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
D = int(sys.argv[3]) if len(sys.argv) > 3 else 512
b_loc = B // W
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
T = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(torch.bfloat16)
I = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(torch.bfloat16)[:b_loc].contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
loss, rinvh, cinvh = ops.infonce_forward(I, T, 0.07)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d_t = torch.empty(B, D, device=dev)
d_i = torch.empty(8, b_loc, D, device=dev)


def launch(directions, splits):
    _lib.check(lib.b200clip_infonce_bwd(_lib.ptr(I), _lib.ptr(T), D, b_loc, B, 0, 0.07, _lib.ptr(rinvh), _lib.ptr(cinvh), None,
                                        _lib.ptr(d_i) if directions & 1 else None, splits,
                                        _lib.ptr(d_t) if directions & 2 else None, directions, _lib.stream_ptr()), "bwd")


def both(splits):
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        launch(2, 1)
    with torch.cuda.stream(s2):
        launch(1, splits)
    cur.wait_stream(s1)
    cur.wait_stream(s2)


def timeit(fn, n=7):
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
    return ts[len(ts) // 2]


fl = 4.0 * b_loc * B * D            # algorithmic flops of ONE direction (S + dX)
print(f"W={W} B={B} D={D} b_loc={b_loc}; one direction = {fl / 1e12:.3f} TFLOP")
t = timeit(lambda: launch(2, 1))
print(f"dT alone            {t * 1e3:8.1f} us  {fl / t / 1e9:6.0f} TF/s algorithmic")
for s in (1, 2, 4, 8):
    t = timeit(lambda: launch(1, s))
    print(f"dI alone splits={s}    {t * 1e3:8.1f} us  {fl / t / 1e9:6.0f} TF/s")
for s in (1, 2, 4, 8):
    t = timeit(lambda: both(s))
    print(f"both, 2 streams s={s} {t * 1e3:8.1f} us  {2 * fl / t / 1e9:6.0f} TF/s")
for s in (1, 4, 8):
    t = timeit(lambda: launch(3, s))
    print(f"one launch, s={s}     {t * 1e3:8.1f} us  {2 * fl / t / 1e9:6.0f} TF/s")

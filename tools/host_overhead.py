"""CPU enqueue time vs GPU time of one head step (host overhead matters at small per-rank batches).
   python tools/host_overhead.py [B]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
sys.path.insert(0, ROOT)
import torch
import b200clip
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = dict(bench.CFG["cfg3"], B=B)
dev = torch.device("cuda:0")
torch.manual_seed(0)
head = b200clip.ClipHead(cfg["E_img"], cfg["E_txt"], cfg["D"], cfg["C"], 0.07, 1.0).to(dev)
x_img, x_txt, labels, class_text = bench.synth_inputs(cfg, B, 0, dev)
x_img.requires_grad_(True)
x_txt.requires_grad_(True)


def step():
    for p in head.parameters():
        p.grad = None
    x_img.grad = None
    x_txt.grad = None
    loss = head(x_img, x_txt, class_text, labels)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    step()
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"B={B}: CPU enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, GPU {e0.elapsed_time(e1) / n:.3f} ms/step")
if len(sys.argv) > 2:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)

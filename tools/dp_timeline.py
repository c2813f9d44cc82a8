"""Kernel timeline of one graphed head step on rank 0 (torch.profiler/CUPTI), for the multi-GPU overlap analysis.
   torchrun --nproc-per-node N tools/dp_timeline.py [global_B] [E_img] [D]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import b200clip
import bench

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    __import__("b200clip").dp.init_process_group(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cfg = dict(bench.CFG["cfg3"], B=B)
if len(sys.argv) > 2:
    cfg["E_img"] = int(sys.argv[2])
if len(sys.argv) > 3:
    cfg["D"] = int(sys.argv[3])
b_loc = B // world
torch.manual_seed(0)
head = b200clip.ClipHead(cfg["E_img"], cfg["E_txt"], cfg["D"], cfg["C"], 0.07, 1.0).to(dev)
x_img, x_txt, labels, class_text = bench.synth_inputs(cfg, b_loc, rank, dev)
g = b200clip.GraphedHeadStep(head, x_img, x_txt, class_text, labels)
for _ in range(5):
    g()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        g()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps: a step starts at each cast kernel following the final kernel of the previous one; use 4 equal chunks
    n = len(evs) // 4
    step = evs[2 * n:3 * n]
    t0 = step[0].time_range.start
    print(f"kernels/memops per step: {n}; step span {step[-1].time_range.end - t0:.1f} us")
    busy = 0.0
    last_end = t0
    for e in step:
        s, en = e.time_range.start - t0, e.time_range.end - t0
        gap = e.time_range.start - last_end
        print(f"{s:9.1f} {en - s:8.1f} us  gap {gap:7.1f}  {e.name[:90]}")
        last_end = max(last_end, e.time_range.end)
g.close()
if world > 1:
    dist.destroy_process_group()

"""2-rank check of GraphedHeadStep (NCCL collectives inside the captured graph). Launch with torchrun."""
import os
import sys
import faulthandler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist
import b200clip
import synth

faulthandler.dump_traceback_later(60, exit=True)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
B, E, D, C = 1024, 768, 512, 16
torch.manual_seed(0)
head = b200clip.ClipHead(E, E, D, C).to(dev)
x_img, x_txt = synth.randn(1, B, E).to(torch.bfloat16), synth.randn(2, B, E).to(torch.bfloat16)
labels, class_text = synth.labels(4, B, C), synth.unit_rows(3, C, D)
n = B // world
sl = slice(rank * n, (rank + 1) * n)
xi = x_img[sl].to(dev).requires_grad_(True)
xt = x_txt[sl].to(dev).requires_grad_(True)
ct, lab = class_text.to(dev), labels[sl].to(dev)
loss = head(xi, xt, ct, lab)
loss.backward()
torch.cuda.synchronize()
print(rank, "eager ok", loss.item(), flush=True)
for p in head.parameters():
    p.grad = None
g = b200clip.GraphedHeadStep(head, xi.detach(), xt.detach(), ct, lab)
torch.cuda.synchronize()
print(rank, "capture ok", flush=True)
for i in range(3):
    gl = g()
    torch.cuda.synchronize()
    print(rank, "replay", i, gl.item(), flush=True)
g.close()
dist.destroy_process_group()
print(rank, 'teardown ok', flush=True)

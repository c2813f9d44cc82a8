"""Three MultiModalAttention forward+backward steps at B = 32768 (profiling target for ncu): python tools/attn_step.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
import torch
import b200clip

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D, C = 512, 16
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
att = b200clip.MultiModalAttention().to(dev)
x = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
t = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=1).to(dev).requires_grad_(True)
gy = torch.randn(B, D, generator=g).to(dev)
for _ in range(3):
    out, w = att(x, t)
    out.backward(gy)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))

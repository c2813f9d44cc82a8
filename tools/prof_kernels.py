"""Small driver for ncu captures: runs each hot kernel a few times at a size that keeps the ~40 ncu replays short.
   python tools/prof_kernels.py [infonce|zeroshot|proj] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
import torch
import b200clip
from b200clip import ops

what = sys.argv[1] if len(sys.argv) > 1 else "infonce"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if what == "infonce":
    I = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1).to(dev)
    T = torch.nn.functional.normalize(0.5 * I.cpu() + 0.5 * torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1), dim=1).to(dev)
    for _ in range(3):
        Ig, Tg = I.clone().requires_grad_(True), T.clone().requires_grad_(True)
        loss = b200clip.contrastive_loss(Ig, Tg, 0.07)
        loss.backward()
    torch.cuda.synchronize()
    print("loss", loss.item())
elif what == "zeroshot":
    X = torch.randn(B, 512, generator=g).to(torch.bfloat16).to(dev)
    P = torch.nn.functional.normalize(torch.randn(28, 512, generator=g), dim=1).to(torch.bfloat16).to(dev)
    for _ in range(3):
        out = ops.zeroshot_score(X, P, pair_mode=True, temperature=0.07, thresholds=[0.5])
    torch.cuda.synchronize()
    print("argmax sum", int(out["argmax"].sum()))

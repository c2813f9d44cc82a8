"""Wait-cycle breakdown of the InfoNCE backward kernel (one cluster): where each warp role stalls.
   python tools/nce_prof.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
import torch
from b200clip import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
I = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1).to(dev).to(torch.bfloat16)
T = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1).to(dev).to(torch.bfloat16)
loss, rinvh, cinvh = ops.infonce_forward(I, T, 0.07)
for _ in range(3):
    ops.infonce_backward(I, T, 0.07, rinvh, cinvh, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.infonce_backward(I, T, 0.07, rinvh, cinvh, None)
e1.record()
torch.cuda.synchronize()
print(f"bwd kernel: {e0.elapsed_time(e1) / 5:.3f} ms at B={B}")
ops.KERNEL_EVENTS["infonce_fwd"] = []
for _ in range(6):
    loss = ops.infonce_forward(I, T, 0.07)[0]
torch.cuda.synchronize()
ts = [a.elapsed_time(b) for a, b in ops.KERNEL_EVENTS["infonce_fwd"]][1:]
ops.KERNEL_EVENTS["infonce_fwd"] = None
print(f"fwd stats kernels: {sum(ts) / len(ts):.3f} ms at B={B}  (loss {float(loss):.6f})")
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
_lib.load().b200clip_debug_set_nce_prof(_lib.ptr(buf))
ops.infonce_backward(I, T, 0.07, rinvh, cinvh, None)
torch.cuda.synchronize()
_lib.load().b200clip_debug_set_nce_prof(None)
v = buf.cpu()[:128].view(2, 8, 8)
tl = buf.cpu()[128:128 + 512].view(2, 32, 8)
names = {0: ("producer", ["b_empty", "a_empty", "-", "-"]), 1: ("S issuer", ["s_empty", "b_full", "a_full", "-"]),
         2: ("dX issuer", ["b_full", "g_full(own)", "g_full(peer)", "-"]), 3: ("epilogue wg0", ["s_full", "g_empty", "tmem_ld", "-"]),
         4: ("epilogue wg1", ["s_full", "g_empty", "tmem_ld", "-"])}
nt = (B + 31) // 32
for h in range(2):
    for role, (nm, ws) in names.items():
        tot = int(v[h, role, 0])
        parts = ", ".join(f"{w}={int(v[h, role, 1 + i])} ({100.0 * int(v[h, role, 1 + i]) / max(tot, 1):.0f}%)" for i, w in enumerate(ws) if w != "-")
        print(f"cta{h} {nm:13s} loop={tot} cyc ({tot / nt:.0f}/tile)  waits: {parts}")

print("timeline (cycles relative to tile 512's load issue on the same CTA); cols: load_issue, S_bfull, S_issued, E_sfull, E_sent, D_gfull, D_issued")
for h in range(2):
    t0 = int(tl[h, 0, 0])
    for i in range(16):
        row = [int(tl[h, i, k]) for k in range(7)]
        print(f"cta{h} tile {512 + i}: " + " ".join(f"{(x - t0) if x else -1:7d}" for x in row))

"""Timing of the SURVEY 8(f) rows built in round 1 (MultiViewFusion, MultiModalAttention, ASL) on one B200: forward + backward
through the public modules, CUDA events, 20 timed iterations after 5 warm-ups, next to the measured peaks.
   python tools/f_rows_bench.py [B]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
sys.path.insert(0, ROOT)
import torch
import b200clip
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D, C = 512, 16
dev = torch.device("cuda:0")
pk = bench.peaks()
g = torch.Generator().manual_seed(0)


def timed(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def timed_graph(fn, n=20, warm=3):
    """the same step captured once into a CUDA graph and replayed: kernel time without the ~25 Python launches"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def kernels(fn, title):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    print(f"# kernels of one {title} step (eager): {len(evs)} launches, {sum(e.time_range.end - e.time_range.start for e in evs):.0f} us busy")
    for e in evs:
        print(f"#   {e.time_range.end - e.time_range.start:7.1f} us  {e.name[:100]}")


def line(name, ms, flops, bytes_):
    tf, gbs = flops / (ms / 1e3) / 1e12, bytes_ / (ms / 1e3) / 1e9
    print(json.dumps({"op": name, "B": B, "ms": round(ms, 4), "rows_per_s": round(B / (ms / 1e3)), "algorithmic_tflops": round(tf, 1),
                      "frac_of_bf16_peak": round(tf / pk["tflops"], 4), "algorithmic_GBps": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / pk["hbm"], 4)}))


# ---- MultiViewFusion fwd+bwd: 6 B (2D*D + D*D) flop; bytes: views in 2*B*D*4, h bf16 w+r, y f32 out, dy in, dx out 2*B*D*4 (+ bf16 temporaries)
fus = b200clip.MultiViewFusion().to(dev).eval()
f = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
l = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
gy = torch.randn(B, D, generator=g).to(dev)


def fusion_step():
    for p in fus.parameters():
        p.grad = None
    f.grad = l.grad = None
    fus(f, l).backward(gy)


line("MultiViewFusion fwd+bwd", timed(fusion_step), 6.0 * B * (2 * D * D + D * D), B * D * (8 + 2 + 4 + 4 + 8) * 1.0)
line("MultiViewFusion fwd+bwd (CUDA graph)", timed_graph(fusion_step), 6.0 * B * (2 * D * D + D * D), B * D * (8 + 2 + 4 + 4 + 8) * 1.0)
kernels(fusion_step, "MultiViewFusion")

# ---- MultiModalAttention fwd+bwd: GEMMs 6 B (2 D*D) flop (+ 2 passes of B*C*D tanh); bytes: x, ip f32 w+r, e bf16 w+r, out, d_out, d_e w+r, d_ip w+r, dx
att = b200clip.MultiModalAttention().to(dev)
x = torch.randn(B, D, generator=g).to(dev).requires_grad_(True)
t = torch.nn.functional.normalize(torch.randn(C, D, generator=g), dim=1).to(dev).requires_grad_(True)


def attn_step():
    for p in att.parameters():
        p.grad = None
    x.grad = t.grad = None
    out, w = att(x, t)
    out.backward(gy)


line("MultiModalAttention fwd+bwd", timed(attn_step), 6.0 * B * (2 * D * D), B * D * (4 + 8 + 4 + 4 + 4 + 8 + 4 + 4) * 1.0)
line("MultiModalAttention fwd+bwd (CUDA graph)", timed_graph(attn_step), 6.0 * B * (2 * D * D), B * D * (4 + 8 + 4 + 4 + 4 + 8 + 4 + 4) * 1.0)
kernels(attn_step, "MultiModalAttention")

# ---- ASL fwd+bwd on [B, C] logits: HBM-bound, 4 * B*C*4 bytes (logits + targets read twice, gradient written)
lg = (torch.randn(B, C, generator=g) * 3).to(dev).requires_grad_(True)
tg = (torch.rand(B, C, generator=g) < 0.2).float().to(dev)


def asl_step():
    lg.grad = None
    b200clip.multilabel_asymmetric_loss(lg, tg).backward()


line("multilabel_asymmetric_loss fwd+bwd", timed(asl_step), 0.0, B * C * 4 * 5.0)
line("multilabel_asymmetric_loss fwd+bwd (CUDA graph)", timed_graph(asl_step), 0.0, B * C * 4 * 5.0)
kernels(asl_step, "ASL")

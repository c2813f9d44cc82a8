"""cfg 5 (BASELINE.json): batch sweep of the fused head step on one GPU, B = 1k .. 64k at D = 512 / 768, graphed step.
   python tools/sweep.py [E_img] [D]      -> one line per B: ms/step, pairs/s, algorithmic TFLOP/s, fraction of the bf16 peak"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
sys.path.insert(0, ROOT)
import torch
import b200clip
import bench

E_img = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
pk = bench.peaks()
print(f"# head step sweep (graphed step, dropout 0.1 on), D={D}, E_img={E_img}, E_txt=768, C=16; peak {pk['tflops']} TFLOP/s ({pk['src']})")
for B in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
    cfg = dict(B=B, D=D, E_img=E_img, E_txt=768, C=16)
    torch.manual_seed(0)
    head = b200clip.ClipHead(E_img, 768, D, 16, 0.07, 1.0, dropout_rate=0.1).to(dev).train()
    x_img, x_txt, labels, class_text = bench.synth_inputs(cfg, B, 0, dev)
    step = b200clip.GraphedHeadStep(head, x_img, x_txt, class_text, labels)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = bench.head_flops(B, D, E_img, 768, 16) / (ms / 1e3) / 1e12
    print(json.dumps({"B": B, "ms_per_step": round(ms, 4), "pairs_per_s": round(B / (ms / 1e3)), "algorithmic_tflops": round(tf, 1),
                      "frac_of_peak": round(tf / pk["tflops"], 4), "loss": round(float(loss), 5)}))
    step.close()
    del step, head
    torch.cuda.empty_cache()

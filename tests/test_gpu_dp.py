"""2-rank NCCL run of the fused head (one process per GPU) vs the single-GPU result on the same global batch.
Skipped when fewer than 2 GPUs are visible (the driver's 1-GPU box); run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest
import torch

from gpu_util import gpu

pytestmark = gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200clip, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
B, E, D, C = 1024, 768, 512, 16
torch.manual_seed(0)
head = b200clip.ClipHead(E, E, D, C).to(dev)
x_img, x_txt = synth.randn(1, B, E).to(torch.bfloat16), synth.randn(2, B, E).to(torch.bfloat16)
labels, class_text = synth.labels(4, B, C), synth.unit_rows(3, C, D)
n = B // world
sl = slice(rank * n, (rank + 1) * n)
xi = x_img[sl].to(dev).requires_grad_(True); xt = x_txt[sl].to(dev).requires_grad_(True)
loss = head(xi, xt, class_text.to(dev), labels[sl].to(dev)); loss.backward()
torch.cuda.synchronize()
out = {"loss": loss.item(), "dxi": xi.grad.float().cpu(), "dxt": xt.grad.float().cpu(),
       "gw": head.image_projector.fc.weight.grad.cpu(), "gt": head.text_projector.text_projection.weight.grad.cpu(),
       "gf": head.classifier.weight.grad.cpu()}
# the same step captured in a CUDA graph (collectives included) must agree with the eager step
for p in head.parameters(): p.grad = None
gstep = b200clip.GraphedHeadStep(head, xi.detach(), xt.detach(), class_text.to(dev), labels[sl].to(dev))
gl = gstep(xi.detach(), xt.detach(), class_text.to(dev), labels[sl].to(dev))
torch.cuda.synchronize()
out["g_loss"] = gl.item(); out["g_dxi"] = gstep.grad_image.float().cpu(); out["g_gw"] = head.image_projector.fc.weight.grad.cpu()
torch.save(out, os.path.join(OUT, f"dp_r{rank}.pt"))
gstep.close()
dist.destroy_process_group()
'''


def test_two_rank_head_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import b200clip
    import synth
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\nOUT = {str(tmp_path)!r}\n" + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                    "--master-port", "29533", str(script)], check=True, env=env, timeout=240)
    dev = torch.device("cuda:0")
    B, E, D, C = 1024, 768, 512, 16
    torch.manual_seed(0)
    head = b200clip.ClipHead(E, E, D, C).to(dev)
    xi = synth.randn(1, B, E).to(torch.bfloat16).to(dev).requires_grad_(True)
    xt = synth.randn(2, B, E).to(torch.bfloat16).to(dev).requires_grad_(True)
    loss = head(xi, xt, synth.unit_rows(3, C, D).to(dev), synth.labels(4, B, C).to(dev))
    loss.backward()
    outs = [torch.load(tmp_path / f"dp_r{r}.pt") for r in range(2)]
    n = B // 2

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())

    for r, o in enumerate(outs):
        assert abs(o["loss"] - loss.item()) <= 1e-5 * abs(loss.item())
        assert rel(o["dxi"], xi.grad.float().cpu()[r * n:(r + 1) * n]) < 1e-2
        assert rel(o["dxt"], xt.grad.float().cpu()[r * n:(r + 1) * n]) < 1e-2
        assert rel(o["gw"], head.image_projector.fc.weight.grad.cpu()) < 1e-2
        assert rel(o["gt"], head.text_projector.text_projection.weight.grad.cpu()) < 1e-2
        assert rel(o["gf"], head.classifier.weight.grad.cpu()) < 1e-2
        assert abs(o["g_loss"] - o["loss"]) <= 1e-6 * abs(o["loss"])
        assert rel(o["g_dxi"], o["dxi"]) < 1e-5
        assert rel(o["g_gw"], o["gw"]) < 1e-4

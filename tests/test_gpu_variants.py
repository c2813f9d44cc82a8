"""The A/B switches of the library select whole kernels: round 1's InfoNCE forward / backward
(B200CLIP_FWD_VARIANT=1, B200CLIP_BWD_VARIANT=4), the un-forked projection backward (B200CLIP_PROJ_BWD_FORK=0) and a forced
dI split count (B200CLIP_BWD_SPLITS).  They are read once per process, so a child process runs the same seeded head step and
rank-shaped InfoNCE backward with the switches set; the results must agree with this process's default kernels."""
import os
import subprocess
import sys

import torch

from gpu_util import dev, gpu, rel_l2

pytestmark = gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BODY = r'''
import os, sys, torch
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200clip, synth
from b200clip import ops


def run(dev):
    torch.manual_seed(0)
    B, E, D, C = 1024, 768, 512, 16
    head = b200clip.ClipHead(E, E, D, C).to(dev)
    xi = synth.randn(1, B, E).to(torch.bfloat16).to(dev).requires_grad_(True)
    xt = synth.randn(2, B, E).to(torch.bfloat16).to(dev).requires_grad_(True)
    loss = head(xi, xt, synth.unit_rows(3, C, D).to(dev), synth.labels(4, B, C).to(dev))
    loss.backward()
    out = {"loss": loss.detach().float().cpu(), "dxi": xi.grad.float().cpu(), "dxt": xt.grad.float().cpu(),
           "gw1": head.image_projector.image_projection.weight.grad.cpu(), "gw2": head.text_projector.fc.weight.grad.cpu()}
    # rank-shaped InfoNCE backward (local rows 256..511 of a 1024-row batch) with column splits
    T = synth.unit_rows(5, B, D).to(torch.bfloat16).to(dev)
    I = synth.unit_rows(6, B, D).to(torch.bfloat16).to(dev)[256:512].contiguous()
    _, rinvh, cinvh = ops.infonce_forward(I, T, 0.07, row0=256)
    d_i, d_t = ops.infonce_backward(I, T, 0.07, rinvh, cinvh, None, row0=256, allow_splits=True)
    out["nce_di"] = (d_i.sum(0) if d_i.dim() == 3 else d_i).cpu()
    out["nce_dt"] = d_t.cpu()
    torch.cuda.synchronize()
    return out
'''


def test_env_switched_kernels_agree_with_defaults(tmp_path):
    ns = {"ROOT": ROOT}
    exec(compile(BODY, "variants_body", "exec"), ns)
    ref = ns["run"](dev())
    script = tmp_path / "child.py"
    out = tmp_path / "child.pt"
    script.write_text(f"ROOT = {ROOT!r}\n" + BODY + f"\ntorch.save(run(torch.device('cuda:0')), {str(out)!r})\n")
    env = dict(os.environ, B200CLIP_FWD_VARIANT="1", B200CLIP_BWD_VARIANT="4", B200CLIP_PROJ_BWD_FORK="0", B200CLIP_BWD_SPLITS="2")
    subprocess.run([sys.executable, str(script)], check=True, env=env, timeout=300)
    got = torch.load(out)
    assert abs(float(got["loss"]) - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    for k in ("dxi", "dxt", "gw1", "gw2", "nce_di", "nce_dt"):
        assert rel_l2(got[k], ref[k]) < 2e-3, k

"""FusedAdamW against torch.optim.AdamW on the head's parameter shapes, several steps."""
import torch

from gpu_util import dev, gpu, rel_l2

pytestmark = gpu


def test_fused_adamw_matches_torch_adamw():
    import b200clip
    d = dev()
    torch.manual_seed(0)
    shapes = [(512, 768), (512,), (512, 512), (16, 512), (16,), (3,)]
    mine = [torch.nn.Parameter(torch.randn(s, device=d)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    kw = dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)           # the reference's optimizer settings
    om, orf = b200clip.FusedAdamW(mine, **kw), torch.optim.AdamW(ref, **kw)
    for it in range(6):
        for a, b in zip(mine, ref):
            g = torch.randn_like(a) * (10.0 ** (it - 3))
            a.grad, b.grad = g.clone(), g.clone()
        om.step()
        orf.step()
    torch.cuda.synchronize()
    for a, b in zip(mine, ref):
        assert rel_l2(a, b) < 1e-6
    # a parameter without a gradient is skipped, like torch
    mine[0].grad = None
    before = mine[0].detach().clone()
    om.step()
    assert torch.equal(mine[0].detach(), before)

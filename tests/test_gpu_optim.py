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


def test_fused_adamw_state_dict_round_trip_and_torch_checkpoint():
    """The reference checkpoints optimizer.state_dict() and resumes with load_state_dict (0426/train.py:846-860, :670): the step
    counter must travel with the state (bias correction restarts at t=1 otherwise), in both directions between FusedAdamW and
    torch.optim.AdamW."""
    import copy
    import b200clip
    d = dev()
    torch.manual_seed(1)
    shapes = [(512, 768), (512,), (16, 512)]
    kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    grads = [[torch.randn(s, device=d) for s in shapes] for _ in range(7)]

    def run(opt, params, its):
        for it in its:
            for p, g in zip(params, grads[it]):
                p.grad = g.clone()
            opt.step()

    init = [torch.randn(s, device=d) for s in shapes]
    ref = [torch.nn.Parameter(p.clone()) for p in init]
    oref = torch.optim.AdamW(ref, **kw)
    run(oref, ref, range(7))                                           # uninterrupted torch run: the truth

    # (1) FusedAdamW: 4 steps, checkpoint, fresh optimizer, resume for 3 steps
    a = [torch.nn.Parameter(p.clone()) for p in init]
    oa = b200clip.FusedAdamW(a, **kw)
    run(oa, a, range(4))
    ck = copy.deepcopy(oa.state_dict())
    assert all(float(s["step"]) == 4.0 for s in ck["state"].values())
    a2 = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa2 = b200clip.FusedAdamW(a2, **kw)
    oa2.load_state_dict(copy.deepcopy(ck))          # load_state_dict adopts the tensors it is given: keep ck pristine
    run(oa2, a2, range(4, 7))
    # (2) torch.optim.AdamW checkpoint after 4 steps resumed by FusedAdamW
    b = [torch.nn.Parameter(p.clone()) for p in init]
    ob = torch.optim.AdamW(b, **kw)
    run(ob, b, range(4))
    b2 = [torch.nn.Parameter(p.detach().clone()) for p in b]
    ob2 = b200clip.FusedAdamW(b2, **kw)
    ob2.load_state_dict(copy.deepcopy(ob.state_dict()))
    run(ob2, b2, range(4, 7))
    # (3) FusedAdamW checkpoint resumed by torch.optim.AdamW
    c2 = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oc2 = torch.optim.AdamW(c2, **kw)
    oc2.load_state_dict(copy.deepcopy(ck))
    run(oc2, c2, range(4, 7))
    torch.cuda.synchronize()
    for got in (a2, b2, c2):
        for x, y in zip(got, ref):
            assert rel_l2(x, y) < 1e-6
    assert float(oa2.state[a2[0]]["step"]) == 7.0

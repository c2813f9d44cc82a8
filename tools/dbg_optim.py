import copy, sys, os
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", "clip-for-dl_b200")]
import torch, b200clip
d = torch.device("cuda:0")
torch.manual_seed(1)
shapes = [(512, 768), (512,), (16, 512)]
kw = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
grads = [[torch.randn(s, device=d) for s in shapes] for _ in range(7)]
def run(opt, params, its):
    for it in its:
        for p, g in zip(params, grads[it]):
            p.grad = g.clone()
        opt.step()
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
init = [torch.randn(s, device=d) for s in shapes]
ref = [torch.nn.Parameter(p.clone()) for p in init]; oref = torch.optim.AdamW(ref, **kw); run(oref, ref, range(7))
f = [torch.nn.Parameter(p.clone()) for p in init]; of = b200clip.FusedAdamW(f, **kw); run(of, f, range(7))
print("uninterrupted fused vs torch", [rel(x, y) for x, y in zip(f, ref)])
a = [torch.nn.Parameter(p.clone()) for p in init]; oa = b200clip.FusedAdamW(a, **kw); run(oa, a, range(4))
r4 = [torch.nn.Parameter(p.clone()) for p in init]; o4 = torch.optim.AdamW(r4, **kw); run(o4, r4, range(4))
print("4 steps fused vs torch", [rel(x, y) for x, y in zip(a, r4)])
ck = copy.deepcopy(oa.state_dict())
print("ck steps", [(float(s["step"]), s["step"].device) for s in ck["state"].values()])
a2 = [torch.nn.Parameter(p.detach().clone()) for p in a]; oa2 = b200clip.FusedAdamW(a2, **kw); oa2.load_state_dict(ck)
print("loaded steps", [(float(s["step"])) for s in oa2.state.values()], "exp_avg equal", [torch.equal(oa2.state[x]["exp_avg"], oa.state[y]["exp_avg"]) for x, y in zip(a2, a)])
run(oa2, a2, range(4, 7))
print("after resume step", float(oa2._step), "case1", [rel(x, y) for x, y in zip(a2, ref)])
b2 = [torch.nn.Parameter(p.detach().clone()) for p in r4]; ob2 = b200clip.FusedAdamW(b2, **kw); ob2.load_state_dict(copy.deepcopy(o4.state_dict())); run(ob2, b2, range(4, 7))
print("case2", [rel(x, y) for x, y in zip(b2, ref)])
c2 = [torch.nn.Parameter(p.detach().clone()) for p in a]; oc2 = torch.optim.AdamW(c2, **kw); oc2.load_state_dict(ck); run(oc2, c2, range(4, 7))
print("case3", [rel(x, y) for x, y in zip(c2, ref)], [float(s["step"]) for s in oc2.state.values()])

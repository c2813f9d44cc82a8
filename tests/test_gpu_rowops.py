"""Row kernels (rowops.cu) vs torch fp32: F.normalize fwd/bwd, LayerNorm fwd/bwd (+fused L2 copy), column sums."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import dev, gpu, rel_l2
import synth

pytestmark = gpu


@pytest.mark.parametrize("rows,D", [(1, 512), (37, 512), (4096, 512), (100, 768), (64, 1024), (9, 128)])
def test_l2norm_fwd_bwd(rows, D):
    from b200clip import ops
    x = synth.randn(1, rows, D).to(dev())
    x[0] *= 1e-3
    xr = x.clone().requires_grad_(True)
    y = F.normalize(xr, dim=-1)
    g = synth.randn(2, rows, D).to(dev())
    y.backward(g)
    yb, yf, inv = ops.l2norm_fwd(x, want_bf16=True, want_f32=True)
    assert rel_l2(yf, y) < 1e-6
    assert rel_l2(yb.float(), y) < 4e-3
    assert rel_l2(ops.l2norm_bwd(g, x, inv), xr.grad) < 1e-5
    # dy given as 3 partial sums + a scaled addend (what the fused head feeds it)
    parts = torch.stack([0.5 * g, 0.25 * g, 0.25 * g])
    add, sc = synth.randn(3, rows, D).to(dev()), torch.full((), 1.5, device=dev())
    assert rel_l2(ops.l2norm_bwd(parts, x, inv, addend=add, addend_scale=sc), xr.grad + 1.5 * add) < 1e-5
    # autograd wrapper + bf16 input path
    xa = x.clone().requires_grad_(True)
    ops.normalize(xa).backward(g)
    assert rel_l2(xa.grad, xr.grad) < 1e-5
    yb2, _, _ = ops.l2norm_fwd(x.to(torch.bfloat16))
    assert rel_l2(yb2.float(), F.normalize(x.to(torch.bfloat16).float(), dim=-1)) < 4e-3


def test_l2norm_zero_row():
    from b200clip import ops
    x = torch.zeros(3, 512, device=dev())
    _, yf, inv = ops.l2norm_fwd(x, want_bf16=False, want_f32=True)
    assert torch.equal(yf, torch.zeros_like(yf))          # x / max(0, eps) = 0, as F.normalize


@pytest.mark.parametrize("rows,D", [(5, 512), (1000, 512), (333, 768), (64, 1024)])
def test_layernorm_fwd_bwd(rows, D):
    from b200clip import _lib, ops
    lib = _lib.load()
    d = dev()
    z = (synth.randn(3, rows, D) * 2 + 0.5).to(d)
    gamma = (1 + 0.1 * synth.randn(4, D)).to(d)
    beta = (0.1 * synth.randn(5, D)).to(d)
    zr, gr, br = z.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yref = F.layer_norm(zr, (D,), gr, br, 1e-5)
    dy = synth.randn(6, rows, D).to(d)
    yref.backward(dy)
    y = torch.empty_like(z)
    yhat = torch.empty(rows, D, dtype=torch.bfloat16, device=d)
    mean, rstd, inv = (torch.empty(rows, device=d) for _ in range(3))
    _lib.check(lib.b200clip_layernorm_fwd(_lib.ptr(z), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(y), _lib.ptr(yhat),
                                          _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(inv), rows, D, 1e-5, 1e-12,
                                          _lib.stream_ptr()), "ln")
    assert rel_l2(y, yref) < 1e-6
    assert rel_l2(yhat.float(), F.normalize(yref, dim=-1)) < 4e-3
    assert rel_l2(inv, 1.0 / yref.norm(dim=-1)) < 1e-6
    dz = torch.empty_like(z)
    dzb = torch.empty(rows, D, dtype=torch.bfloat16, device=d)
    dg, db, dzs = torch.empty(D, device=d), torch.empty(D, device=d), torch.empty(D, device=d)
    nb = lib.b200clip_layernorm_bwd_workspace_bytes(rows, D)
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    _lib.check(lib.b200clip_layernorm_bwd(_lib.ptr(dy), _lib.ptr(z), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
                                          _lib.ptr(dz), _lib.ptr(dzb), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(dzs), 0, rows, D, 0.0, 0, None, _lib.ptr(ws), nb,
                                          _lib.stream_ptr()), "lnb")
    assert rel_l2(dz, zr.grad) < 1e-5
    assert rel_l2(dzb.float(), zr.grad) < 4e-3
    assert rel_l2(dg, gr.grad) < 1e-5
    assert rel_l2(db, br.grad) < 1e-5
    assert rel_l2(dzs, zr.grad.sum(0)) < 1e-4


@pytest.mark.parametrize("rows,D,parts,with_addend", [(1000, 512, 1, False), (777, 512, 3, True), (130, 1024, 2, True)])
def test_layernorm_backward_with_fused_l2norm_backward(rows, D, parts, with_addend):
    """b200clip_layernorm_l2_bwd == autograd through  normalize(LayerNorm(z))  (+ a gradient that reaches y directly),
    with the normalised features given in bf16 and d/dyhat given as partial sums."""
    from b200clip import _lib
    lib = _lib.load()
    d = dev()
    z = (synth.randn(3, rows, D) * 2 + 0.5).to(d)
    gamma = (1 + 0.1 * synth.randn(4, D)).to(d)
    beta = (0.1 * synth.randn(5, D)).to(d)
    zr, gr, br = z.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.layer_norm(zr, (D,), gr, br, 1e-5)
    nrm = y.norm(dim=-1, keepdim=True)
    yhat_b = (y / nrm).detach().to(torch.bfloat16)
    # reference: the kernel sees yhat only in bf16, so differentiate  y -> y / ||y||  at the bf16-rounded direction
    g = synth.randn(6, rows, D).to(d)
    add = synth.randn(7, rows, D).to(d) if with_addend else None
    sc = torch.full((), 0.75, device=d)
    yh = yhat_b.float()
    dy = (g - yh * (yh * g).sum(-1, keepdim=True)) / nrm.detach()
    if add is not None:
        dy = dy + 0.75 * add
    y.backward(dy)
    gp = torch.stack([g / parts] * parts).contiguous()
    mean, rstd = z.mean(-1), 1.0 / torch.sqrt(z.var(-1, unbiased=False) + 1e-5)
    inv = (1.0 / nrm.detach().squeeze(-1)).contiguous()
    dz = torch.empty_like(z)
    dzb = torch.empty(rows, D, dtype=torch.bfloat16, device=d)
    dg, db, dzs = torch.empty(D, device=d), torch.empty(D, device=d), torch.empty(D, device=d)
    nb = lib.b200clip_layernorm_bwd_workspace_bytes(rows, D)
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    _lib.check(lib.b200clip_layernorm_l2_bwd(_lib.ptr(gp), parts, _lib.ptr(yhat_b), _lib.ptr(inv), 1e-12, _lib.ptr(add),
                                             _lib.ptr(sc if add is not None else None), _lib.ptr(z), _lib.ptr(mean.contiguous()),
                                             _lib.ptr(rstd.contiguous()), _lib.ptr(gamma), _lib.ptr(dz), _lib.ptr(dzb), _lib.ptr(dg),
                                             _lib.ptr(db), _lib.ptr(dzs), 0, rows, D, 0.0, 0, None, _lib.ptr(ws), nb, _lib.stream_ptr()),
               "ln_l2_bwd")
    assert rel_l2(dz, zr.grad) < 2e-5
    assert rel_l2(dzb.float(), zr.grad) < 4e-3
    assert rel_l2(dg, gr.grad) < 2e-5
    assert rel_l2(db, br.grad) < 2e-5
    assert rel_l2(dzs, zr.grad.sum(0)) < 1e-4


@pytest.mark.parametrize("rows,N,bf16", [(1000, 512, False), (257, 768, True), (5, 128, False), (40000, 512, True), (300, 2048, False)])
def test_colsum(rows, N, bf16):
    from b200clip import _lib
    lib = _lib.load()
    d = dev()
    a = synth.randn(7, rows, N).to(d)
    if bf16:
        a = a.to(torch.bfloat16)
    out = torch.empty(N, device=d)
    nb = lib.b200clip_colsum_workspace_bytes(rows, N)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=d)
    _lib.check(lib.b200clip_colsum(_lib.ptr(a), int(bf16), N, rows, N, _lib.ptr(out), 0, _lib.ptr(ws), nb, _lib.stream_ptr()), "cs")
    assert rel_l2(out, a.float().sum(0)) < 1e-5

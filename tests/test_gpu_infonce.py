"""Fused InfoNCE kernels (infonce.cu) through the C ABI vs the oracle (reference contrastive_loss, 0426/train.py:154-176)
on the same bf16-rounded, L2-normalised inputs.  Tolerances from BASELINE.json north_star: loss 1e-3 relative,
gradients 2e-2 relative L2."""
import pytest
import torch

from gpu_util import dev, gpu, rel_l2
import ref_head as R
import synth

pytestmark = gpu
LOSS_TOL = 1e-3
GRAD_TOL = 2e-2


def _inputs(B, seed=11, corr=0.5, D=512):
    """unit rows, bf16-rounded; text correlated with image so the diagonal carries signal like trained CLIP."""
    I = synth.unit_rows(seed, B, D)
    T = synth.unit_rows(seed + 1, B, D)
    T = R.l2_normalize(corr * I + (1 - corr) * T)
    return synth.bf16_round(I), synth.bf16_round(T)


def _oracle(I, T, tau):
    I = I.clone().requires_grad_(True)
    T = T.clone().requires_grad_(True)
    loss = R.contrastive_loss(I, T, tau)
    loss.backward()
    return loss.detach(), I.grad, T.grad


# D = 768 (BASELINE.json configs[4], 0426/config.py:30 `shared_embedding_size`): 3-CTA clusters in the backward pass
@pytest.mark.parametrize("B,tau,D", [(128, 0.07, 512), (256, 0.07, 512), (200, 0.07, 512), (1000, 0.07, 512), (384, 1.0, 512),
                                     (2048, 0.07, 512), (33, 0.5, 512), (128, 0.07, 768), (200, 0.07, 768), (1000, 0.07, 768),
                                     (2048, 0.07, 768), (97, 1.0, 768)])
def test_loss_and_grads_match_reference(B, tau, D):
    import b200clip
    I, T = _inputs(B, D=D)
    loss_ref, dI_ref, dT_ref = _oracle(I, T, tau)
    Ig = I.to(dev()).requires_grad_(True)
    Tg = T.to(dev()).requires_grad_(True)
    loss = b200clip.contrastive_loss(Ig, Tg, tau)
    loss.backward()
    # + 2e-5: small batches with a dominant diagonal have losses ~ 1e-2, where the bf16 logits' absolute error shows
    assert abs(loss.item() - loss_ref.item()) <= LOSS_TOL * abs(loss_ref.item()) + 2e-5, (loss.item(), loss_ref.item())
    assert rel_l2(Ig.grad, dI_ref) < GRAD_TOL
    assert rel_l2(Tg.grad, dT_ref) < GRAD_TOL


# D = 512: stationary-X CTA-pair kernel; other widths (768 = BASELINE.json configs[4], 256, 1024, 64): streamed-X variant.
# b_loc < b_glob: a data-parallel rank's row block against all columns; 1: a single row; ragged sizes exercise the masks.
@pytest.mark.parametrize("b_loc,b_glob,D", [(640, 640, 512), (300, 300, 512), (256, 1024, 512), (1, 130, 512), (4096, 4096, 512),
                                            (640, 640, 768), (300, 1000, 768), (384, 384, 256), (200, 200, 1024), (129, 257, 64)])
def test_statistics_match_flash_oracle(b_loc, b_glob, D):
    from b200clip import _lib, ops
    lib = _lib.load()
    tau = 0.07
    I, T = _inputs(b_glob, D=D)
    row0 = (b_glob - b_loc) // 2 // 64 * 64
    I = I[row0:row0 + b_loc]
    r_ref, c_ref, diag_ref, m = R.contrastive_loss_flash(I.double(), T.double(), tau, row0=row0)
    d = dev()
    ib, tb = I.to(d).to(torch.bfloat16).contiguous(), T.to(d).to(torch.bfloat16)
    nb = lib.b200clip_infonce_workspace_bytes(b_loc, b_glob)
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    r, c = torch.empty(b_loc, device=d), torch.empty(b_glob, device=d)
    _lib.check(lib.b200clip_infonce_fwd_stats(_lib.ptr(ib), _lib.ptr(tb), D, b_loc, b_glob, tau, _lib.ptr(r), _lib.ptr(c), _lib.ptr(ws),
                                              nb, _lib.stream_ptr()), "stats")
    assert rel_l2(r, r_ref) < 1e-4
    assert rel_l2(c, c_ref) < 1e-4
    # the loss numerators (diagonal included) for this width
    rinvh, cinvh = torch.empty_like(r), torch.empty_like(c)
    sums = torch.empty(3, dtype=torch.float64, device=d)
    _lib.check(lib.b200clip_infonce_loss(_lib.ptr(ib), _lib.ptr(tb), D, b_loc, b_glob, row0, tau, _lib.ptr(r), _lib.ptr(c), 0, b_glob,
                                         _lib.ptr(rinvh), _lib.ptr(cinvh), _lib.ptr(sums), None, _lib.ptr(ws), nb, _lib.stream_ptr()), "loss")
    assert abs(sums[2].item() - diag_ref.item()) <= 1e-4 * abs(diag_ref.item())
    assert abs(sums[0].item() - torch.log(r_ref).sum().item()) <= 1e-4 * abs(torch.log(r_ref).sum().item()) + 1e-3


def test_upstream_gradient_scale_and_determinism():
    import b200clip
    I, T = _inputs(512)
    outs = []
    for _ in range(2):
        Ig = I.to(dev()).requires_grad_(True)
        Tg = T.to(dev()).requires_grad_(True)
        (3.0 * b200clip.contrastive_loss(Ig, Tg, 0.07)).backward()
        outs.append((Ig.grad.clone(), Tg.grad.clone()))
    _, dI_ref, dT_ref = _oracle(I, T, 0.07)
    assert rel_l2(outs[0][0], 3.0 * dI_ref) < GRAD_TOL
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])     # bit-reproducible


@pytest.mark.parametrize("W,D", [(2, 512), (4, 512), (8, 512), (4, 768)])
def test_rank_partitioned_equals_monolithic(W, D):
    """Simulated data-parallel ranks on ONE GPU: each 'rank' runs the kernels on its row block against all columns;
    column sums are SUM-combined and dT partials summed (what all_reduce / reduce_scatter do across GPUs)."""
    from b200clip import _lib, ops
    lib = _lib.load()
    B, tau = 1024, 0.07
    n = B // W
    I, T = _inputs(B, D=D)
    loss_ref, dI_ref, dT_ref = _oracle(I, T, tau)
    d = dev()
    ib, tb = I.to(d).to(torch.bfloat16), T.to(d).to(torch.bfloat16)
    nb = lib.b200clip_infonce_workspace_bytes(n, B)
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    rs, cs = [], torch.zeros(B, device=d)
    for k in range(W):
        r, c = torch.empty(n, device=d), torch.empty(B, device=d)
        blk = ib[k * n:(k + 1) * n].contiguous()
        _lib.check(lib.b200clip_infonce_fwd_stats(_lib.ptr(blk), _lib.ptr(tb), D, n, B, tau, _lib.ptr(r), _lib.ptr(c),
                                                  _lib.ptr(ws), nb, _lib.stream_ptr()), "stats")
        rs.append(r)
        cs += c
    total = torch.zeros(3, dtype=torch.float64, device=d)
    dI, dT = [], torch.zeros(B, D, device=d)
    for k in range(W):
        blk = ib[k * n:(k + 1) * n].contiguous()
        rinvh, cinvh = torch.empty(n, device=d), torch.empty(B, device=d)
        sums = torch.empty(3, dtype=torch.float64, device=d)
        _lib.check(lib.b200clip_infonce_loss(_lib.ptr(blk), _lib.ptr(tb), D, n, B, k * n, tau, _lib.ptr(rs[k]), _lib.ptr(cs),
                                             k * n, (k + 1) * n, _lib.ptr(rinvh), _lib.ptr(cinvh), _lib.ptr(sums), None,
                                             _lib.ptr(ws), nb, _lib.stream_ptr()), "loss")
        total += sums
        # allow_splits: with b_loc << b_glob the kernel cuts direction 0's columns into ranges (partial d_i sums)
        d_i, d_t = ops.infonce_backward(blk, tb, tau, rinvh, cinvh, None, row0=k * n, allow_splits=True)
        assert (d_i.dim() == 3) == (W >= 2) and (d_i.dim() == 2 or d_i.shape[0] == min(W, 4)), d_i.shape
        # the two directions launched separately (the data-parallel step does that to overlap the reduce-scatter of d_t with
        # direction 0) are bit-identical to the combined launch
        d_i1, none_t = ops.infonce_backward(blk, tb, tau, rinvh, cinvh, None, row0=k * n, allow_splits=True, directions=1)
        none_i, d_t2 = ops.infonce_backward(blk, tb, tau, rinvh, cinvh, None, row0=k * n, allow_splits=True, directions=2)
        assert none_t is None and none_i is None and torch.equal(d_i1, d_i) and torch.equal(d_t2, d_t)
        dI.append(d_i.sum(0) if d_i.dim() == 3 else d_i)
        dT += d_t
    loss = 1.0 / tau + (total[0] + total[1]) / (2.0 * B) - total[2] / B
    assert abs(loss.item() - loss_ref.item()) <= LOSS_TOL * abs(loss_ref.item())
    assert rel_l2(torch.cat(dI), dI_ref) < GRAD_TOL
    assert rel_l2(dT, dT_ref) < GRAD_TOL


def test_properties_at_full_size():
    """Size-independent checks at the bench size (B=32768): rows of dI are orthogonal... no -- use exact identities:
    sum_j G_ij = 0 per row => sum over all gradients of <dI_i, anything> is not free, but
    (1) sum_i dI_i . I_i + ... ; we use: d/d(scale) of the loss under I -> sI equals sum_i <dI_i, I_i>, and by symmetry
    sum_i <dI_i, I_i> == sum_j <dT_j, T_j> exactly in exact arithmetic (both equal sum_ij G_ij S_ij)."""
    import b200clip
    B = 32768
    g = torch.Generator(device="cpu").manual_seed(5)
    I = torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1)
    T = torch.nn.functional.normalize(0.5 * I + 0.5 * torch.nn.functional.normalize(torch.randn(B, 512, generator=g), dim=1), dim=1)
    Ig = I.to(dev()).to(torch.bfloat16).float().requires_grad_(True)
    Tg = T.to(dev()).to(torch.bfloat16).float().requires_grad_(True)
    loss = b200clip.contrastive_loss(Ig, Tg, 0.07)
    loss.backward()
    assert torch.isfinite(loss)
    a = (Ig.grad.double() * Ig.detach().double()).sum()
    b = (Tg.grad.double() * Tg.detach().double()).sum()
    assert abs(a.item() - b.item()) <= 2e-2 * max(abs(a.item()), 1e-6)
    # loss upper bound log(B) + 2/tau and lower bound 0
    assert 0.0 <= loss.item() <= torch.log(torch.tensor(float(B))).item() + 2 / 0.07
    # spot-check 256 rows of the gradient against the oracle evaluated on those rows only (needs full column stats)
    rows = torch.arange(0, B, 128)
    S = (Ig.detach()[rows].double().cpu() @ Tg.detach().double().cpu().T) / 0.07
    m = 1 / 0.07
    E = torch.exp(S - m)
    r = E.sum(1)
    # column sums need all rows: compute on the GPU in fp32 blocks
    c = torch.zeros(B, dtype=torch.float64)
    for s0 in range(0, B, 4096):
        blk = (Ig.detach()[s0:s0 + 4096] @ Tg.detach().T) / 0.07
        c += torch.exp(blk.double() - m).sum(0).cpu()
    G = E * (1 / r[:, None] + 1 / c[None, :]) / (2 * B)
    G[torch.arange(len(rows)), rows] -= 1.0 / B
    dI_ref = (G @ Tg.detach().double().cpu()) / 0.07
    assert rel_l2(Ig.grad[rows], dI_ref) < 2e-2


def test_rejects_unsupported_shapes():
    import b200clip
    I = synth.unit_rows(5, 64, 200).to(dev())
    with pytest.raises(RuntimeError):
        b200clip.contrastive_loss(I, I, 0.07)              # unit rows -> flash path; D = 200 is not a multiple of 64: loud failure, no fallback
    with pytest.raises(RuntimeError):
        b200clip.contrastive_loss(torch.randn(64, 512, device=dev()), torch.randn(16, 512, device=dev()), 0.07)


# corr is kept small: with a dominant diagonal at these temperatures the softmax saturates (loss ~ 1e-12, G ~ -1e-12) and any fp32
# implementation -- the reference's F.cross_entropy included -- only carries rounding noise relative to the fp64 truth
@pytest.mark.parametrize("B,tau,corr", [(200, 0.01, 0.0), (1000, 0.01, 0.12), (512, 0.02, 0.0), (256, 0.008, 0.1)])
def test_small_temperature_does_not_underflow(B, tau, corr):
    """CLIP clamps tau at 0.01.  With the plain shift m = 1/tau every exponential of a row whose cosines are all below
    1 - 126 tau ln2 (0.125 at tau = 0.01: any untrained batch) flushes to zero -> r = 0 -> inf loss, NaN gradients.  The
    kernels lower the shift (csrc/host.cuh nce_k2); loss and gradients must match the reference's stable F.cross_entropy."""
    import b200clip
    I, T = _inputs(B, seed=21, corr=corr)
    loss_ref, dI_ref, dT_ref = _oracle(I.double(), T.double(), tau)
    Ig, Tg = I.to(dev()).requires_grad_(True), T.to(dev()).requires_grad_(True)
    loss = b200clip.contrastive_loss(Ig, Tg, tau)
    loss.backward()
    assert torch.isfinite(loss) and torch.isfinite(Ig.grad).all() and torch.isfinite(Tg.grad).all()
    assert abs(loss.item() - loss_ref.item()) <= LOSS_TOL * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    assert rel_l2(Ig.grad, dI_ref) < GRAD_TOL and rel_l2(Tg.grad, dT_ref) < GRAD_TOL


def test_temperature_below_supported_range_is_rejected():
    import b200clip
    I, T = _inputs(128)
    with pytest.raises(RuntimeError, match="temperature"):
        b200clip.contrastive_loss(I.to(dev()), T.to(dev()), 0.004)


def test_unnormalised_inputs_take_the_general_path(golden):
    """0426/train.py:154-176 accepts any inputs; the flash path needs unit rows.  The public function must detect the
    difference and stay correct: reference golden (un-normalised randn, tau 1.0), then LayerNorm-scale inputs at tau 0.07
    (|logit| ~ 10^3) and mildly off-unit rows against the fp64 oracle, loss AND gradients."""
    import b200clip
    from b200clip import ops
    I, T = synth.randn(13, 20, 32), synth.randn(14, 20, 32)
    assert not ops.rows_are_unit(I.to(dev()), T.to(dev()))
    loss = b200clip.contrastive_loss(I.to(dev()), T.to(dev()), 1.0)
    assert abs(loss.item() - float(golden["nce_loss_unnorm"])) <= 1e-5 * abs(float(golden["nce_loss_unnorm"]))
    for B, D, scale, tau in ((300, 512, 22.6, 0.07), (64, 512, 1.3, 0.07), (1000, 128, 4.0, 0.5)):
        I, T = synth.unit_rows(31, B, D) * scale, synth.unit_rows(32, B, D) * scale
        loss_ref, dI_ref, dT_ref = _oracle(I.double(), T.double(), tau)
        Ig, Tg = I.to(dev()).requires_grad_(True), T.to(dev()).requires_grad_(True)
        loss = b200clip.contrastive_loss(Ig, Tg, tau)
        (loss * 3.0).backward()
        assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item()), (B, loss.item(), loss_ref.item())
        assert rel_l2(Ig.grad, 3.0 * dI_ref) < 1e-3 and rel_l2(Tg.grad, 3.0 * dT_ref) < 1e-3
    # unit rows keep the flash path; an explicit (wrong) inputs_normalized=True is the caller's contract
    Iu, Tu = _inputs(256)
    assert ops.rows_are_unit(Iu.to(dev()), Tu.to(dev()))
    big = torch.randn(8200, 512, device=dev())
    with pytest.raises(RuntimeError, match="not L2-normalised"):
        b200clip.contrastive_loss(big, big, 0.07)

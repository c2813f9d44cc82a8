"""Thin functional layer over the C ABI: tensor allocation + argument marshalling, and the autograd Functions
that give the reference's modules/losses their backward passes.  PyTorch is used for device memory, streams
and autograd bookkeeping only; all arithmetic happens in libb200clip.so.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib, dp
from ._lib import check, load, ptr, require_cuda, stream_ptr

LN_EPS = 1e-5          # nn.LayerNorm default (0426/train.py:82)
L2_EPS = 1e-12         # F.normalize default

EPI_STORE_F32, EPI_STORE_BF16, EPI_BIAS_GELU, EPI_BIAS_RESID_F32, EPI_ATOMIC_F32, EPI_GELU_BWD, EPI_RELU_BF16 = range(7)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 copy through b200clip_cast_f32_bf16 (bf16 inputs pass through)."""
    require_cuda(x)
    if x.dtype == torch.bfloat16:
        return x.contiguous()
    x = _f32c(x)
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    n = x.numel()
    if n % 4 == 0:
        check(load().b200clip_cast_f32_bf16(ptr(x), ptr(out), n, stream_ptr()), "cast_f32_bf16")
    else:                                    # ragged tail: pad to a multiple of 4 in a scratch copy
        pad = (-n) % 4
        xin = torch.zeros(n + pad, dtype=torch.float32, device=x.device)
        xin[:n] = x.reshape(-1)
        o = torch.empty(n + pad, dtype=torch.bfloat16, device=x.device)
        check(load().b200clip_cast_f32_bf16(ptr(xin), ptr(o), n + pad, stream_ptr()), "cast_f32_bf16")
        out = o[:n].reshape(x.shape).contiguous()
    return out


# --------------------------------------------------------------------------------------------------------------
# generic GEMM (tests + projection internals)
# --------------------------------------------------------------------------------------------------------------
def gemm_bf16(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, epilogue: int = EPI_STORE_F32,
              alpha: float = 1.0, bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
              aux: Optional[torch.Tensor] = None, split_k: int = 1, out: Optional[torch.Tensor] = None):
    require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    a, b = a.contiguous(), b.contiguous()
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N = b.shape[1] if b_mn else b.shape[0]
    assert (b.shape[0] if b_mn else b.shape[1]) == K, "inner dimensions differ"
    out1 = None
    if epilogue in (EPI_STORE_F32, EPI_BIAS_RESID_F32, EPI_ATOMIC_F32):
        if out is None:
            out = (torch.zeros if epilogue == EPI_ATOMIC_F32 else torch.empty)((M, N), dtype=torch.float32, device=a.device)
    else:
        if out is None:
            out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
        if epilogue == EPI_BIAS_GELU:
            out1 = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    rc = load().b200clip_gemm_bf16(ptr(a), ptr(b), int(a_mn), int(b_mn), M, N, K, a.shape[1], b.shape[1], epilogue, alpha,
                                   ptr(out), N, ptr(out1), N, ptr(bias), ptr(resid), N, ptr(aux), N, split_k, stream_ptr())
    check(rc, "gemm_bf16")
    return (out, out1) if out1 is not None else out


# --------------------------------------------------------------------------------------------------------------
# a-L2
# --------------------------------------------------------------------------------------------------------------
def l2norm_fwd(x: torch.Tensor, want_bf16: bool = True, want_f32: bool = False, eps: float = L2_EPS):
    require_cuda(x)
    assert x.dim() == 2
    if x.dtype != torch.bfloat16:
        x = _f32c(x)
    x = x.contiguous()
    rows, D = x.shape
    yb = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    yf = torch.empty((rows, D), dtype=torch.float32, device=x.device) if want_f32 else None
    inv = torch.empty((rows,), dtype=torch.float32, device=x.device)
    check(load().b200clip_l2norm_fwd(ptr(x), int(x.dtype == torch.bfloat16), D, ptr(yb), ptr(yf), ptr(inv), rows, D, eps,
                                     stream_ptr()), "l2norm_fwd")
    return yb, yf, inv


def l2norm_bwd(dy: torch.Tensor, x: torch.Tensor, inv: torch.Tensor, eps: float = L2_EPS, out: Optional[torch.Tensor] = None,
               accumulate: bool = False, addend: Optional[torch.Tensor] = None,
               addend_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx = d normalize(x) . dy  [+ addend * addend_scale]  (addend [rows, D] f32, addend_scale a device scalar).
    dy may be [S, rows, D]: S partial sums (the InfoNCE backward's column splits), added on the fly."""
    dy = _f32c(dy)
    rows, D = x.shape
    parts = dy.shape[0] if dy.dim() == 3 else 1
    if out is None:
        out = torch.empty((rows, D), dtype=torch.float32, device=x.device)
    check(load().b200clip_l2norm_bwd(ptr(dy), parts, ptr(x), int(x.dtype == torch.bfloat16), D, ptr(inv), ptr(out), int(accumulate),
                                     rows, D, eps, ptr(addend), ptr(addend_scale), stream_ptr()), "l2norm_bwd")
    return out


class L2NormalizeFn(torch.autograd.Function):
    """F.normalize(x, dim=-1) on [rows, D] (0426/train.py:191-192)."""

    @staticmethod
    def forward(ctx, x):
        xc = _f32c(x)
        _, yf, inv = l2norm_fwd(xc, want_bf16=False, want_f32=True)
        ctx.save_for_backward(xc, inv)
        return yf

    @staticmethod
    def backward(ctx, dy):
        xc, inv = ctx.saved_tensors
        return l2norm_bwd(dy, xc, inv)


def normalize(x: torch.Tensor, dim: int = -1) -> torch.Tensor:
    if dim not in (-1, x.dim() - 1):
        raise ValueError("b200clip.normalize: only the last dimension is supported")
    shp = x.shape
    return L2NormalizeFn.apply(x.reshape(-1, shp[-1])).reshape(shp)


def dropout_mask(rows: int, cols: int, p: float, seed: int, device) -> torch.Tensor:
    """The scaled keep-mask (keep ? 1/(1-p) : 0) the fused kernels apply for (p, seed) -- for tests and debugging."""
    out = torch.empty((rows, cols), dtype=torch.float32, device=device)
    check(load().b200clip_dropout_mask(ptr(out), rows, cols, float(p), int(seed) & 0xFFFFFFFF, stream_ptr()), "dropout_mask")
    return out


def dropout_seed_advance(seed_dev: torch.Tensor) -> None:
    """Advance a device-resident dropout seed word in stream order (one node of GraphedHeadStep's graph)."""
    require_cuda(seed_dev)
    check(load().b200clip_dropout_seed_advance(ptr(seed_dev), stream_ptr()), "dropout_seed_advance")


def new_dropout_seed() -> int:
    """A fresh 32-bit seed from torch's CPU generator (honours torch.manual_seed; no device sync)."""
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())


# --------------------------------------------------------------------------------------------------------------
# a-P1 / a-P2 projection block
# --------------------------------------------------------------------------------------------------------------
def proj_fwd(x_bf16, w1_bf16, b1, w2_bf16, b2, gamma, beta, want_yhat: bool, drop_p: float = 0.0, drop_seed: int = 0,
             drop_seed_dev: Optional[torch.Tensor] = None, want_y: bool = True):
    """drop_seed_dev: optional device int32 word added to drop_seed inside the kernels (GraphedHeadStep's per-replay seed).
    want_y=False skips the fp32 LayerNorm output (the fused head only consumes the normalised bf16 copy and 1/||y||)."""
    B, E = x_bf16.shape
    D = w1_bf16.shape[0]
    dev = x_bf16.device
    p = torch.empty((B, D), dtype=torch.bfloat16, device=dev)
    h = torch.empty((B, D), dtype=torch.bfloat16, device=dev)
    z = torch.empty((B, D), dtype=torch.float32, device=dev)
    y = torch.empty((B, D), dtype=torch.float32, device=dev) if want_y else None
    mean = torch.empty((B,), dtype=torch.float32, device=dev)
    rstd = torch.empty((B,), dtype=torch.float32, device=dev)
    yhat = torch.empty((B, D), dtype=torch.bfloat16, device=dev) if want_yhat else None
    inv = torch.empty((B,), dtype=torch.float32, device=dev) if want_yhat else None
    check(load().b200clip_proj_fwd(ptr(x_bf16), B, E, D, ptr(w1_bf16), ptr(b1), ptr(w2_bf16), ptr(b2), ptr(gamma), ptr(beta),
                                   LN_EPS, float(drop_p), int(drop_seed) & 0xFFFFFFFF, ptr(drop_seed_dev), ptr(p), ptr(h), ptr(z), ptr(y), ptr(yhat),
                                   ptr(mean), ptr(rstd), ptr(inv),
                                   stream_ptr()), "proj_fwd")
    return y, yhat, inv, (p, h, z, mean, rstd)


def proj_bwd(dy, x_bf16, w1_bf16, w2_bf16, gamma, saved, need_dx: bool, dx_dtype=torch.float32, drop_p: float = 0.0,
             drop_seed: int = 0, l2=None, drop_seed_dev: Optional[torch.Tensor] = None):
    """dy: gradient w.r.t. the block's output y.  Fused form: dy=None and l2=(dyhat, yhat_bf16, inv_norm, addend,
    addend_scale) -- the gradient w.r.t. the L2-normalised output ([S,]B,D partial sums allowed), the normalised bf16
    output, 1/||y|| and an optional f32 addend times a device scalar; the L2-norm backward runs inside LayerNorm's."""
    p, h, z, mean, rstd = saved
    B, E = x_bf16.shape
    D = w1_bf16.shape[0]
    dev = x_bf16.device
    dyhat = yhat = inv = addend = ascale = None
    parts = 1
    if dy is not None:
        dy = _f32c(dy)
    else:
        dyhat, yhat, inv, addend, ascale = l2
        dyhat = _f32c(dyhat)
        parts = dyhat.shape[0] if dyhat.dim() == 3 else 1
    dx_bf = need_dx and dx_dtype == torch.bfloat16
    dx = torch.empty((B, E), dtype=torch.bfloat16 if dx_bf else torch.float32, device=dev) if need_dx else None
    dw1 = torch.empty((D, E), dtype=torch.float32, device=dev)
    dw2 = torch.empty((D, D), dtype=torch.float32, device=dev)
    db1, db2, dg, dbeta = (torch.empty((D,), dtype=torch.float32, device=dev) for _ in range(4))
    nb = load().b200clip_proj_bwd_workspace_bytes(B, E, D)
    ws = _ws(nb, dev)
    check(load().b200clip_proj_bwd(ptr(dy), ptr(dyhat), parts, ptr(yhat), ptr(inv), ptr(addend), ptr(ascale), ptr(x_bf16), B, E, D, ptr(w1_bf16), ptr(w2_bf16), ptr(gamma), ptr(p), ptr(h), ptr(z),
                                   ptr(mean), ptr(rstd), float(drop_p), int(drop_seed) & 0xFFFFFFFF, ptr(drop_seed_dev), ptr(None if dx_bf else dx), ptr(dx if dx_bf else None), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ptr(dg), ptr(dbeta),
                                   ptr(ws), ws.numel(), stream_ptr()), "proj_bwd")
    return dx, dw1, db1, dw2, db2, dg, dbeta


class ProjectionFn(torch.autograd.Function):
    """ImageProjection/TextProjection forward+backward (0426/train.py:84-96); dropout (p, seed) is fused into the second
    GEMM's epilogue and regenerated in the backward pass from the same counter-based hash."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, drop_p=0.0, drop_seed=0):
        require_cuda(x, w1)
        xb = cast_bf16(x)
        w1b, w2b = cast_bf16(w1), cast_bf16(w2)
        b1, b2, gamma, beta = _f32c(b1), _f32c(b2), _f32c(gamma), _f32c(beta)
        y, _, _, saved = proj_fwd(xb, w1b, b1, w2b, b2, gamma, beta, want_yhat=False, drop_p=drop_p, drop_seed=drop_seed)
        ctx.save_for_backward(xb, w1b, w2b, gamma, *saved)
        ctx.drop = (float(drop_p), int(drop_seed))
        ctx.need_dx = x.requires_grad
        ctx.x_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, w1b, w2b, gamma, p, h, z, mean, rstd = ctx.saved_tensors
        dx, dw1, db1, dw2, db2, dg, dbeta = proj_bwd(dy, xb, w1b, w2b, gamma, (p, h, z, mean, rstd), ctx.need_dx,
                                                     torch.bfloat16 if ctx.x_dtype == torch.bfloat16 else torch.float32,
                                                     drop_p=ctx.drop[0], drop_seed=ctx.drop[1])
        if dx is not None and dx.dtype != ctx.x_dtype:
            dx = dx.to(ctx.x_dtype)
        return dx, dw1, db1, dw2, db2, dg, dbeta, None, None


# --------------------------------------------------------------------------------------------------------------
# MultiViewFusion (0426/train.py:988-1000)
# --------------------------------------------------------------------------------------------------------------
class FusionFn(torch.autograd.Function):
    """cat[frontal, lateral] -> Linear(2D, D) -> ReLU -> Dropout -> Linear(D, D); dropout (p, seed) is fused into the first
    GEMM's epilogue, the saved activations carry the mask into the backward pass."""

    @staticmethod
    def forward(ctx, frontal, lateral, w0, b0, w3, b3, drop_p=0.0, drop_seed=0):
        require_cuda(frontal, lateral, w0, w3)
        lib = load()
        f, l = _f32c(frontal), _f32c(lateral)
        B, D = f.shape
        if l.shape != f.shape or w0.shape != (D, 2 * D) or w3.shape != (D, D):
            raise RuntimeError(f"b200clip.MultiViewFusion: expected two [B,{D}] views and weights [{D},{2 * D}], [{D},{D}]")
        dev = f.device
        x = torch.empty((B, 2 * D), dtype=torch.bfloat16, device=dev)            # the concatenation, written by two strided casts
        check(lib.b200clip_cast_f32_bf16_2d(ptr(f), D, ptr(x), 2 * D, B, D, stream_ptr()), "cast_2d")
        check(lib.b200clip_cast_f32_bf16_2d(ptr(l), D, C.c_void_p(x.data_ptr() + 2 * D), 2 * D, B, D, stream_ptr()), "cast_2d")
        w0b, w3b = cast_bf16(w0), cast_bf16(w3)
        h = torch.empty((B, D), dtype=torch.bfloat16, device=dev)
        y = torch.empty((B, D), dtype=torch.float32, device=dev)
        check(lib.b200clip_fusion_fwd(ptr(x), B, D, ptr(w0b), ptr(_f32c(b0)), ptr(w3b), ptr(_f32c(b3)), float(drop_p),
                                      int(drop_seed) & 0xFFFFFFFF, ptr(h), ptr(y), stream_ptr()), "fusion_fwd")
        ctx.save_for_backward(x, w0b, w3b, h)
        ctx.drop_p = float(drop_p)
        ctx.need_dx = frontal.requires_grad or lateral.requires_grad
        ctx.dtypes = (frontal.dtype, lateral.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = load()
        x, w0b, w3b, h = ctx.saved_tensors
        B, D2 = x.shape
        D = D2 // 2
        dev = x.device
        dy = _f32c(dy)
        dx = torch.empty((2, B, D), dtype=torch.float32, device=dev) if ctx.need_dx else None      # d frontal | d lateral
        dw0 = torch.empty((D, D2), dtype=torch.float32, device=dev)
        dw3 = torch.empty((D, D), dtype=torch.float32, device=dev)
        db0, db3 = torch.empty((D,), dtype=torch.float32, device=dev), torch.empty((D,), dtype=torch.float32, device=dev)
        ws = _ws(lib.b200clip_fusion_bwd_workspace_bytes(B, D), dev)
        check(lib.b200clip_fusion_bwd(ptr(dy), ptr(x), B, D, ptr(w0b), ptr(w3b), ptr(h), ctx.drop_p, ptr(dx), ptr(dw0), ptr(db0),
                                      ptr(dw3), ptr(db3), ptr(ws), ws.numel(), stream_ptr()), "fusion_bwd")
        df = dl = None
        if dx is not None:
            df, dl = dx[0].to(ctx.dtypes[0]), dx[1].to(ctx.dtypes[1])
        return df, dl, dw0, db0, dw3, db3, None, None


# --------------------------------------------------------------------------------------------------------------
# MultiModalAttention (multimodal_attention/train.py:1069-1110)
# --------------------------------------------------------------------------------------------------------------
class AttentionFn(torch.autograd.Function):
    """(enhanced_features [B,D], attn_weights [B,C]) = MultiModalAttention(image_features [B,D], text_features [C,D])."""

    @staticmethod
    def forward(ctx, image_features, text_features, wi, bi, wt, bt, wa, ba, wo, bo):
        require_cuda(image_features, text_features, wi, wo)
        lib = load()
        B, D = image_features.shape
        Cn = text_features.shape[0]
        dev = image_features.device
        x = cast_bf16(image_features)
        t = _f32c(text_features)
        wib, wob = cast_bf16(wi), cast_bf16(wo)
        wtf, waf = _f32c(wt), _f32c(wa).reshape(-1)
        ip = torch.empty((B, D), dtype=torch.float32, device=dev)
        tp = torch.empty((Cn, D), dtype=torch.float32, device=dev)
        w = torch.empty((B, Cn), dtype=torch.float32, device=dev)
        e = torch.empty((B, D), dtype=torch.bfloat16, device=dev)
        out = torch.empty((B, D), dtype=torch.float32, device=dev)
        check(lib.b200clip_attention_fwd(ptr(x), ptr(t), B, Cn, D, ptr(wib), ptr(_f32c(bi)), ptr(wtf), ptr(_f32c(bt)), ptr(waf),
                                         ptr(_f32c(ba)), ptr(wob), ptr(_f32c(bo)), ptr(ip), ptr(tp), ptr(w), ptr(e), ptr(out),
                                         stream_ptr()), "attention_fwd")
        ctx.save_for_backward(x, t, wib, wtf, waf, _f32c(ba), wob, ip, tp, w, e)
        ctx.need = (image_features.requires_grad, text_features.requires_grad)
        ctx.dtypes = (image_features.dtype, text_features.dtype)
        ctx.wa_shape = wa.shape
        return out, w

    @staticmethod
    def backward(ctx, d_out, d_w):
        lib = load()
        x, t, wib, wtf, waf, ba, wob, ip, tp, w, e = ctx.saved_tensors
        B, D = ip.shape
        Cn = tp.shape[0]
        dev = ip.device
        d_out = _f32c(d_out) if d_out is not None else torch.zeros((B, D), dtype=torch.float32, device=dev)
        d_w = _f32c(d_w) if d_w is not None else None
        f32 = dict(dtype=torch.float32, device=dev)
        dx = torch.empty((B, D), **f32) if ctx.need[0] else None
        dt = torch.empty((Cn, D), **f32) if ctx.need[1] else None
        dwi, dwt, dwo = torch.empty((D, D), **f32), torch.empty((D, D), **f32), torch.empty((D, D), **f32)
        dbi, dbt, dbo, dwa = (torch.empty((D,), **f32) for _ in range(4))
        dba = torch.empty((1,), **f32)
        ws = _ws(lib.b200clip_attention_bwd_workspace_bytes(B, Cn, D), dev)
        check(lib.b200clip_attention_bwd(ptr(d_out), ptr(d_w), ptr(x), ptr(t), B, Cn, D, ptr(wib), ptr(wtf), ptr(waf), ptr(ba), ptr(wob),
                                         ptr(ip), ptr(tp), ptr(w), ptr(e), ptr(dx), ptr(dt), ptr(dwi), ptr(dbi), ptr(dwt), ptr(dbt),
                                         ptr(dwa), ptr(dba), ptr(dwo), ptr(dbo), ptr(ws), ws.numel(), stream_ptr()), "attention_bwd")
        if dx is not None:
            dx = dx.to(ctx.dtypes[0])
        if dt is not None:
            dt = dt.to(ctx.dtypes[1])
        return dx, dt, dwi, dbi, dwt, dbt, dwa.reshape(ctx.wa_shape), dba, dwo, dbo


# --------------------------------------------------------------------------------------------------------------
# a-N symmetric InfoNCE
# --------------------------------------------------------------------------------------------------------------
def infonce_forward(i_hat: torch.Tensor, t_hat: torch.Tensor, temperature: float, row0: int = 0, group=None,
                    sums_out: Optional[torch.Tensor] = None, loss_stream: Optional[torch.cuda.Stream] = None,
                    keep: Optional[list] = None):
    """i_hat: local rows [b_loc, D] bf16; t_hat: all rows [b_glob, D] bf16.  Returns (loss, rinvh, cinvh).
    With sums_out (3 doubles) the rank-local loss numerators are written there and loss is None (the fused head sums
    them over ranks together with the BCE numerators and finalises once).  With loss_stream (needs sums_out and `keep`) the
    loss numerators are computed on that stream, off the critical path: the current stream only runs the half-inverse
    statistics the backward pass needs; the caller joins loss_stream and must hold `keep` (the tensors that stream reads)
    until then."""
    lib = load()
    b_loc, D = i_hat.shape
    b_glob = t_hat.shape[0]
    dev = i_hat.device
    nb = lib.b200clip_infonce_workspace_bytes(b_loc, b_glob)
    ws = _ws(nb, dev)
    r = torch.empty((b_loc,), dtype=torch.float32, device=dev)
    c = torch.empty((b_glob,), dtype=torch.float32, device=dev)
    ev = KERNEL_EVENTS["infonce_fwd"]
    if ev is not None:
        e0, e1 = _timing_events()
        e0.record()
    check(lib.b200clip_infonce_fwd_stats(ptr(i_hat), ptr(t_hat), D, b_loc, b_glob, temperature, ptr(r), ptr(c), ptr(ws),
                                         ws.numel(), stream_ptr()), "infonce_fwd_stats")
    if ev is not None:
        e1.record()
        ev.append((e0, e1))
    world = dp.world(group) if (group is not None or b_loc != b_glob) else 1
    if world > 1:
        dp.sum_across(c, group)                               # partial column sums -> global
    rinvh = torch.empty_like(r)
    cinvh = torch.empty_like(c)
    deferred = sums_out is not None
    sums = sums_out if deferred else torch.empty((3,), dtype=torch.float64, device=dev)
    loss = None if deferred else torch.empty((), dtype=torch.float32, device=dev)
    c_lo, c_hi = (row0, row0 + b_loc) if world > 1 else (0, b_glob)
    if loss_stream is not None:
        assert deferred and keep is not None
        check(lib.b200clip_infonce_inv_stats(ptr(r), b_loc, ptr(c), b_glob, ptr(rinvh), ptr(cinvh), stream_ptr()), "infonce_inv_stats")
        loss_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(loss_stream):
            check(lib.b200clip_infonce_loss(ptr(i_hat), ptr(t_hat), D, b_loc, b_glob, row0, temperature, ptr(r), ptr(c), c_lo, c_hi,
                                            None, None, ptr(sums), None, ptr(ws), ws.numel(), stream_ptr()), "infonce_loss")
        keep.extend((ws, r, c, sums, i_hat, t_hat))
        return None, rinvh, cinvh
    check(lib.b200clip_infonce_loss(ptr(i_hat), ptr(t_hat), D, b_loc, b_glob, row0, temperature, ptr(r), ptr(c), c_lo, c_hi,
                                    ptr(rinvh), ptr(cinvh), ptr(sums), ptr(loss) if (world == 1 and not deferred) else None,
                                    ptr(ws), ws.numel(), stream_ptr()), "infonce_loss")
    if world > 1 and not deferred:
        loss = dp.infonce_loss_from_sums(dp.sum_across(sums, group), temperature, b_glob)
    return loss, rinvh, cinvh


# bench.py sets this to a list to collect (start, end) CUDA-event pairs around the dominant kernel (roofline leg)
KERNEL_EVENTS = {"infonce_bwd": None, "infonce_fwd": None}


def _timing_events():
    """Timing event pair; inside a CUDA-graph capture they become external event-record nodes, re-recorded by every
    replay, so the kernel between them can be timed on the launching stream while the graph runs."""
    ext = torch.cuda.is_current_stream_capturing()
    return (torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext))


def infonce_backward(i_hat, t_hat, temperature, rinvh, cinvh, grad_scale: Optional[torch.Tensor], row0: int = 0,
                     allow_splits: bool = False, directions: int = 3):
    """Returns d_i [b_loc, D] f32 and d_t_partial [b_glob, D] f32 (this rank's contribution).  With allow_splits (data
    parallel, b_loc << b_glob) d_i may come back as [S, b_loc, D] partial sums over column ranges (l2norm_bwd adds them).
    directions: 3 = both gradients in one launch; 1 = d_i only (d_t is None); 2 = d_t only (d_i is None)."""
    b_loc, D = i_hat.shape
    b_glob = t_hat.shape[0]
    dev = i_hat.device
    splits = int(load().b200clip_infonce_bwd_splits(b_loc, b_glob)) if allow_splits else 1
    d_i = torch.empty((splits, b_loc, D) if splits > 1 else (b_loc, D), dtype=torch.float32, device=dev) if directions & 1 else None
    d_t = torch.empty((b_glob, D), dtype=torch.float32, device=dev) if directions & 2 else None
    gs = None
    if grad_scale is not None:
        gs = _f32c(grad_scale.reshape(()))
    ev = KERNEL_EVENTS["infonce_bwd"]
    if ev is not None:
        e0, e1 = _timing_events()
        e0.record()
    check(load().b200clip_infonce_bwd(ptr(i_hat), ptr(t_hat), D, b_loc, b_glob, row0, temperature, ptr(rinvh), ptr(cinvh),
                                      ptr(gs), ptr(d_i), splits, ptr(d_t), int(directions), stream_ptr()), "infonce_bwd")
    if ev is not None:
        e1.record()
        ev.append((e0, e1))
    return d_i, d_t


class InfoNCEFn(torch.autograd.Function):
    """contrastive_loss(image_features, text_features, temperature) -- 0426/train.py:154-176, single device.
    Inputs must be L2-normalised (the only call sites pass F.normalize outputs); they are rounded to bf16."""

    @staticmethod
    def forward(ctx, image_features, text_features, temperature):
        require_cuda(image_features, text_features)
        ib, tb = cast_bf16(image_features), cast_bf16(text_features)
        loss, rinvh, cinvh = infonce_forward(ib, tb, float(temperature))
        ctx.save_for_backward(ib, tb, rinvh, cinvh)
        ctx.temperature = float(temperature)
        ctx.dtypes = (image_features.dtype, text_features.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        ib, tb, rinvh, cinvh = ctx.saved_tensors
        d_i, d_t = infonce_backward(ib, tb, ctx.temperature, rinvh, cinvh, grad_out)
        return d_i.to(ctx.dtypes[0]), d_t.to(ctx.dtypes[1]), None


def rows_are_unit(*tensors: torch.Tensor, tol: float = 2e-2) -> bool:
    """True when every row of every tensor is a unit vector (| ||x||^2 - 1 | <= tol; bf16-rounded unit vectors are inside
    4e-3).  One tiny kernel per tensor + ONE host read of a device flag."""
    flag = torch.zeros((), dtype=torch.int32, device=tensors[0].device)
    for t in tensors:
        x = _f32c(t)
        check(load().b200clip_rows_unit_check(ptr(x), x.shape[0], x.shape[1], float(tol), ptr(flag), stream_ptr()), "rows_unit_check")
    return int(flag.item()) == 0


class InfoNCEGeneralFn(torch.autograd.Function):
    """contrastive_loss for inputs that are not unit vectors (0426/train.py:154-176 accepts anything): fp32 logits with true
    row / column maxima, n <= 8192 (b200clip_infonce_general_fwd_bwd).  Forward computes the gradients too (one pass)."""

    @staticmethod
    def forward(ctx, image_features, text_features, temperature):
        require_cuda(image_features, text_features)
        lib = load()
        i, t = _f32c(image_features), _f32c(text_features)
        n, D = i.shape
        dev = i.device
        need = image_features.requires_grad or text_features.requires_grad
        loss = torch.empty((), dtype=torch.float32, device=dev)
        di = torch.empty((n, D), dtype=torch.float32, device=dev) if need else None
        dt = torch.empty((n, D), dtype=torch.float32, device=dev) if need else None
        ws = _ws(lib.b200clip_softclip_workspace_bytes(n), dev)
        check(lib.b200clip_infonce_general_fwd_bwd(ptr(i), ptr(t), n, D, float(temperature), None, ptr(loss), ptr(di), ptr(dt), ptr(ws),
                                                   ws.numel(), stream_ptr()), "infonce_general_fwd_bwd")
        if need:
            ctx.save_for_backward(di, dt)
        ctx.dtypes = (image_features.dtype, text_features.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        di, dt = ctx.saved_tensors
        g = _f32c(g)
        return (di * g).to(ctx.dtypes[0]), (dt * g).to(ctx.dtypes[1]), None


# --------------------------------------------------------------------------------------------------------------
# a-B multi-label BCE on sigmoid(cos/tau)
# --------------------------------------------------------------------------------------------------------------
def _label_sum(labels: torch.Tensor) -> torch.Tensor:
    out = torch.empty((), dtype=torch.float32, device=labels.device)
    check(load().b200clip_sum_f32(ptr(labels), labels.numel(), ptr(out), stream_ptr()), "sum_f32")
    return out


def mlbce(image_features, text_features, labels, temperature, *, grad_scale=None, want_dx=False, want_coef=False,
          label_sum=None, total_elems=None, finalize=True, dx_accum: Optional[torch.Tensor] = None):
    lib = load()
    x = _f32c(image_features)
    t = _f32c(text_features)
    y = _f32c(labels)
    B, D = x.shape
    Cn = t.shape[0]
    dev = x.device
    if label_sum is None:
        label_sum = _label_sum(y)
    if total_elems is None:
        total_elems = float(B) * Cn
    dx = dx_accum if dx_accum is not None else (torch.empty((B, D), dtype=torch.float32, device=dev) if want_dx else None)
    coef = torch.empty((B, Cn), dtype=torch.float32, device=dev) if want_coef else None
    xinv = torch.empty((B,), dtype=torch.float32, device=dev) if want_coef else None
    sums = torch.empty((2,), dtype=torch.float64, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev) if finalize else None
    status = torch.zeros((), dtype=torch.int32, device=dev) if finalize else None
    ws = _ws(lib.b200clip_smallc_workspace_bytes(B, Cn, D), dev)
    gs = _f32c(grad_scale.reshape(())) if grad_scale is not None else None
    check(lib.b200clip_mlbce_fwd_bwd(ptr(x), D, ptr(t), ptr(y), y.shape[1], y.shape[1], B, Cn, D, float(temperature),
                                     ptr(label_sum), float(total_elems), ptr(gs), ptr(dx), int(dx_accum is not None), ptr(coef), ptr(xinv), ptr(sums),
                                     ptr(loss), ptr(status), ptr(ws), ws.numel(), stream_ptr()), "mlbce_fwd_bwd")
    return loss, status, sums, dx, coef, xinv, label_sum


def skinny_outer(coef, x, row_scale=None, want_bias=False, out_scale: Optional[torch.Tensor] = None):
    lib = load()
    B, Cn = coef.shape
    D = x.shape[1]
    dev = x.device
    out_w = torch.empty((Cn, D), dtype=torch.float32, device=dev)
    out_b = torch.empty((Cn,), dtype=torch.float32, device=dev) if want_bias else None
    ws = _ws(lib.b200clip_smallc_workspace_bytes(B, Cn, D), dev)
    check(lib.b200clip_skinny_outer(ptr(coef), Cn, ptr(x), x.stride(0), ptr(row_scale), B, D, ptr(out_w), ptr(out_b), 0,
                                    ptr(out_scale), ptr(ws), ws.numel(), stream_ptr()), "skinny_outer")
    return out_w, out_b


class MultilabelContrastiveFn(torch.autograd.Function):
    """multilabel_contrastive_loss (0426/train.py:178-230), non-fallback branch; the guard flag is returned."""

    @staticmethod
    def forward(ctx, image_features, text_features, labels, temperature):
        require_cuda(image_features, text_features, labels)
        loss, status, _, _, _, _, lsum = mlbce(image_features, text_features, labels, temperature)
        ctx.save_for_backward(image_features, text_features, labels, lsum)
        ctx.temperature = float(temperature)
        ctx.mark_non_differentiable(status)
        return loss, status

    @staticmethod
    def backward(ctx, grad_out, _grad_status):
        image_features, text_features, labels, lsum = ctx.saved_tensors
        need_t = ctx.needs_input_grad[1]
        _, _, _, dx, coef, xinv, _ = mlbce(image_features, text_features, labels, ctx.temperature, grad_scale=grad_out,
                                           want_dx=ctx.needs_input_grad[0], want_coef=need_t, label_sum=lsum)
        dt = None
        if need_t:
            # d t_hat[c] = sum_i coef[i,c] * x_hat[i] / tau, then back through the text normalisation
            x = _f32c(image_features)
            t = _f32c(text_features)
            dth, _ = skinny_outer(coef, x, row_scale=xinv)
            dth = dth * (1.0 / ctx.temperature)
            _, _, tinv = l2norm_fwd(t, want_bf16=False, want_f32=False)
            dt = l2norm_bwd(dth, t, tinv).to(text_features.dtype)
        if dx is not None:
            dx = dx.to(image_features.dtype)
        return dx, dt, None, None


# --------------------------------------------------------------------------------------------------------------
# a-A FC adapter + BCEWithLogits
# --------------------------------------------------------------------------------------------------------------
def fc_bce(x, weight, bias, labels=None, *, grad_scale=None, want_dx=False, want_coef=False, want_pred=False,
           want_logits=False, total_elems=None, threshold=0.5, finalize=True, dx_accum: Optional[torch.Tensor] = None):
    lib = load()
    x, weight = _f32c(x), _f32c(weight)
    bias = _f32c(bias) if bias is not None else None
    B, D = x.shape
    Cn = weight.shape[0]
    dev = x.device
    y = _f32c(labels) if labels is not None else torch.zeros((B, Cn), dtype=torch.float32, device=dev)
    if total_elems is None:
        total_elems = float(B) * Cn
    dx = dx_accum if dx_accum is not None else (torch.empty((B, D), dtype=torch.float32, device=dev) if want_dx else None)
    coef = torch.empty((B, Cn), dtype=torch.float32, device=dev) if want_coef else None
    pred = torch.empty((B, Cn), dtype=torch.float32, device=dev) if want_pred else None
    logits = torch.empty((B, Cn), dtype=torch.float32, device=dev) if want_logits else None
    sums = torch.empty((2,), dtype=torch.float64, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev) if finalize else None
    ws = _ws(lib.b200clip_smallc_workspace_bytes(B, Cn, D), dev)
    gs = _f32c(grad_scale.reshape(())) if grad_scale is not None else None
    check(lib.b200clip_fc_bce_fwd_bwd(ptr(x), x.stride(0), ptr(weight), ptr(bias), ptr(y), y.stride(0), B, Cn, D,
                                      float(total_elems), float(threshold), ptr(gs), ptr(dx), int(dx_accum is not None), ptr(coef), ptr(pred),
                                      ptr(logits), ptr(sums), ptr(loss), ptr(ws), ws.numel(), stream_ptr()), "fc_bce_fwd_bwd")
    return loss, sums, dx, coef, pred, logits


class FcBceFn(torch.autograd.Function):
    """BCEWithLogitsLoss()(Linear(512,16)(x), labels) fused (NB02 c28:50-52, c29:23-25)."""

    @staticmethod
    def forward(ctx, x, weight, bias, labels):
        require_cuda(x, weight, labels)
        loss, *_ = fc_bce(x, weight, bias, labels)
        ctx.save_for_backward(x, weight, bias, labels)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        x, weight, bias, labels = ctx.saved_tensors
        _, _, dx, coef, _, _ = fc_bce(x, weight, bias, labels, grad_scale=grad_out, want_dx=ctx.needs_input_grad[0],
                                      want_coef=True)
        dw, db = skinny_outer(coef, _f32c(x), want_bias=True)
        return dx, dw.to(weight.dtype), (db.to(bias.dtype) if bias is not None else None), None


class LinearSmallFn(torch.autograd.Function):
    """nn.Linear(D, C<=32) forward (logits) with backward -- the adapter used stand-alone (NB02 c30:42)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        require_cuda(x, weight)
        *_, logits = fc_bce(x, weight, bias, None, want_logits=True)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return logits

    @staticmethod
    def backward(ctx, dz):
        x, weight = ctx.saved_tensors
        dz = _f32c(dz)
        dw, db = skinny_outer(dz, _f32c(x), want_bias=True)
        dx = None
        if ctx.needs_input_grad[0]:
            raise RuntimeError("b200clip: input gradient of the stand-alone adapter Linear is not implemented; "
                               "use fc_adapter_bce (fused) or freeze the encoder as NB02 c28:54-62 does")
        return dx, dw, (db if ctx.has_bias else None)


def bce_heads(image_features, class_text, fc_weight, fc_bias, labels, temperature, *, label_sum, total_elems_text,
              total_elems_fc, grad_scale=None, dx_accum=None, dx_out=None, want_coef=False, finalize=True,
              sums_out: Optional[torch.Tensor] = None):
    """Both BCE heads (a-B on the class texts + a-A FC adapter) in one pass over the image features.
    dx_accum: input gradient is ADDED to this tensor; dx_out: input gradient overwrites this tensor."""
    assert dx_accum is None or dx_out is None
    dx_t = dx_accum if dx_accum is not None else dx_out
    lib = load()
    x, t, w = _f32c(image_features), _f32c(class_text), _f32c(fc_weight)
    b = _f32c(fc_bias) if fc_bias is not None else None
    y = _f32c(labels)
    B, D = x.shape
    c1, c2 = t.shape[0], w.shape[0]
    dev = x.device
    coef = torch.empty((B, c2), dtype=torch.float32, device=dev) if want_coef else None
    sums = sums_out if sums_out is not None else torch.empty((3,), dtype=torch.float64, device=dev)
    l_text = torch.empty((), dtype=torch.float32, device=dev) if finalize else None
    l_fc = torch.empty((), dtype=torch.float32, device=dev) if finalize else None
    status = torch.zeros((), dtype=torch.int32, device=dev) if finalize else None
    ws = _ws(lib.b200clip_smallc_workspace_bytes(B, c1 + c2, D), dev)
    gs = _f32c(grad_scale.reshape(())) if grad_scale is not None else None
    check(lib.b200clip_bce_heads_fwd_bwd(ptr(x), x.stride(0), ptr(t), c1, ptr(w), ptr(b), c2, ptr(y), y.shape[1], y.stride(0), B, D,
                                         float(temperature), ptr(label_sum), float(total_elems_text), float(total_elems_fc),
                                         ptr(gs), ptr(dx_t), int(dx_accum is not None), ptr(coef), ptr(sums), ptr(l_text),
                                         ptr(l_fc), ptr(status), ptr(ws), ws.numel(), stream_ptr()), "bce_heads_fwd_bwd")
    return l_text, l_fc, status, sums, coef


def heads_mma_supported(D: int, c1: int, c2: int) -> bool:
    return D in (512, 768) and c1 == 16 and c2 == 16


def bce_heads_mma(yhat_bf16, inv_norm, class_text, fc_weight, fc_bias, labels, temperature, *, label_sum, total_elems_text,
                  total_elems_fc, sums_out, want_grad=True):
    """Both BCE heads on tensor cores from the normalised bf16 features (head step fast path, D = 512 or 768, 16+16 classes).
    Returns (d_y [B,D] f32 for upstream gradient 1, coefn [B,16] bf16, db_raw [16] f32); sums_out receives 3 doubles."""
    lib = load()
    B, D = yhat_bf16.shape
    dev = yhat_bf16.device
    t, w = _f32c(class_text), _f32c(fc_weight)
    b = _f32c(fc_bias) if fc_bias is not None else None
    y = _f32c(labels)
    d_y = torch.empty((B, D), dtype=torch.float32, device=dev) if want_grad else None
    coefn = torch.empty((B, 16), dtype=torch.bfloat16, device=dev) if want_grad else None
    db = torch.empty((16,), dtype=torch.float32, device=dev) if want_grad else None
    ws = _ws(lib.b200clip_bce_heads_mma_workspace_bytes(B), dev)
    check(lib.b200clip_bce_heads_mma_fwd(ptr(yhat_bf16), ptr(inv_norm), B, D, ptr(t), t.shape[0], ptr(w), ptr(b), w.shape[0], ptr(y),
                                         y.shape[1], y.stride(0), float(temperature), ptr(label_sum), float(total_elems_text),
                                         float(total_elems_fc), ptr(d_y), ptr(coefn), ptr(db), ptr(sums_out), ptr(ws), ws.numel(),
                                         stream_ptr()), "bce_heads_mma_fwd")
    return d_y, coefn, db


def skinny_outer_mma(coefn, yhat_bf16, db_raw, out_scale=None):
    """dW_fc [16, D] = out_scale * coefn^T yhat and db = out_scale * db_raw."""
    lib = load()
    B, D = yhat_bf16.shape
    dev = yhat_bf16.device
    out_w = torch.empty((16, D), dtype=torch.float32, device=dev)
    out_b = torch.empty((16,), dtype=torch.float32, device=dev)
    ws = _ws(lib.b200clip_bce_heads_mma_workspace_bytes(B), dev)
    check(lib.b200clip_skinny_outer_mma(ptr(coefn), ptr(yhat_bf16), B, D, 16, ptr(out_scale), ptr(out_w), ptr(db_raw), ptr(out_b),
                                        ptr(ws), ws.numel(), stream_ptr()), "skinny_outer_mma")
    return out_w, out_b


# --------------------------------------------------------------------------------------------------------------
# multilabel_asymmetric_loss (ASL), multimodal_attention/train.py:233-268
# --------------------------------------------------------------------------------------------------------------
_ASL_RED = {"none": 0, "mean": 1, "sum": 2}


def _asl_call(logits, targets, gamma_pos, gamma_neg, clip, eps, red, grad_scale=None, grad_elem=None, want_elem=False, want_grad=False):
    lib = load()
    z, t = _f32c(logits), _f32c(targets)
    dev = z.device
    n = z.numel()
    loss_elem = torch.empty_like(z) if want_elem else None
    d_logits = torch.empty_like(z) if want_grad else None
    s = torch.empty((1,), dtype=torch.float64, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    ws = _ws(lib.b200clip_asl_workspace_bytes(), dev)
    check(lib.b200clip_asl_fwd_bwd(ptr(z), ptr(t), n, float(gamma_pos), float(gamma_neg), float(clip or 0.0), float(eps), red,
                                   ptr(grad_scale), ptr(grad_elem), ptr(loss_elem), ptr(d_logits), ptr(s), ptr(loss), ptr(ws),
                                   ws.numel(), stream_ptr()), "asl_fwd_bwd")
    return loss, loss_elem, d_logits


class AslFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, gamma_pos, gamma_neg, clip, eps, reduction):
        require_cuda(logits, targets)
        if logits.shape != targets.shape:
            raise RuntimeError("multilabel_asymmetric_loss: logits and targets must have the same shape")
        red = _ASL_RED[reduction]
        loss, loss_elem, _ = _asl_call(logits, targets, gamma_pos, gamma_neg, clip, eps, red, want_elem=(red == 0))
        ctx.save_for_backward(logits, targets)
        ctx.cfg = (gamma_pos, gamma_neg, clip, eps, red)
        return loss_elem.reshape(logits.shape) if red == 0 else loss

    @staticmethod
    def backward(ctx, g):
        logits, targets = ctx.saved_tensors
        gamma_pos, gamma_neg, clip, eps, red = ctx.cfg
        g = _f32c(g)
        _, _, d = _asl_call(logits, targets, gamma_pos, gamma_neg, clip, eps, red, grad_scale=None if red == 0 else g.reshape(()),
                            grad_elem=g if red == 0 else None, want_grad=True)
        return d.reshape(logits.shape).to(logits.dtype), None, None, None, None, None, None


# --------------------------------------------------------------------------------------------------------------
# a-S soft-target CLIP loss (0426/train.py:127-152)
# --------------------------------------------------------------------------------------------------------------
def softclip_logits(text, image, temperature):
    require_cuda(text, image)
    t, i = _f32c(text), _f32c(image)
    n, D = t.shape
    out = torch.empty((n, i.shape[0]), dtype=torch.float32, device=t.device)
    if i.shape != t.shape:
        raise RuntimeError("contrastive_clip_loss_function: text and image projections must have the same shape")
    check(load().b200clip_softclip_logits(ptr(t), ptr(i), n, D, float(temperature), ptr(out), stream_ptr()), "softclip_logits")
    return out


def _softclip_call(t, i, temperature, grad_scale, want_grad):
    lib = load()
    n, D = t.shape
    dev = t.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    dt = torch.empty((n, D), dtype=torch.float32, device=dev) if want_grad else None
    di = torch.empty((n, D), dtype=torch.float32, device=dev) if want_grad else None
    ws = _ws(lib.b200clip_softclip_workspace_bytes(n), dev)
    check(lib.b200clip_softclip_fwd_bwd(ptr(t), ptr(i), n, D, float(temperature), ptr(grad_scale), ptr(loss), ptr(dt), ptr(di),
                                        ptr(ws), ws.numel(), stream_ptr()), "softclip_fwd_bwd")
    return loss, dt, di


class SoftClipFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, text, image, temperature):
        require_cuda(text, image)
        if text.shape != image.shape or text.dim() != 2:
            raise RuntimeError("contrastive_clip_loss_function: text and image projections must both be [B, D]")
        t, i = _f32c(text), _f32c(image)
        loss, _, _ = _softclip_call(t, i, temperature, None, False)
        ctx.save_for_backward(t, i)
        ctx.temperature = float(temperature)
        ctx.dtypes = (text.dtype, image.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        t, i = ctx.saved_tensors
        _, dt, di = _softclip_call(t, i, ctx.temperature, _f32c(g).reshape(()), True)     # recomputes the n x n matrices
        return dt.to(ctx.dtypes[0]), di.to(ctx.dtypes[1]), None


def head_loss_finalize(sums6, label_sum, tau_nce, b_glob, total_text, total_fc):
    """loss, parts[3] (InfoNCE, text BCE, FC BCE), status from the six numerators (already summed over ranks)."""
    dev = sums6.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    parts = torch.empty((3,), dtype=torch.float32, device=dev)
    status = torch.empty((), dtype=torch.int32, device=dev)
    check(load().b200clip_head_loss_finalize(ptr(sums6), ptr(label_sum), float(tau_nce), float(b_glob), float(total_text),
                                             float(total_fc), ptr(loss), ptr(parts), ptr(status), stream_ptr()),
          "head_loss_finalize")
    return loss, parts, status


# --------------------------------------------------------------------------------------------------------------
# a-M
# --------------------------------------------------------------------------------------------------------------
def predict_multilabel_raw(image_features, text_features, threshold: float, temperature: float) -> torch.Tensor:
    require_cuda(image_features, text_features)
    x, t = _f32c(image_features), _f32c(text_features)
    B, D = x.shape
    Cn = t.shape[0]
    pred = torch.empty((B, Cn), dtype=torch.float32, device=x.device)
    check(load().b200clip_predict_multilabel(ptr(x), x.stride(0), ptr(t), B, Cn, D, float(temperature), float(threshold),
                                             ptr(pred), stream_ptr()), "predict_multilabel")
    return pred


# --------------------------------------------------------------------------------------------------------------
# a-Z
# --------------------------------------------------------------------------------------------------------------
def _logit(p: float) -> float:
    if p <= 0.0:
        return -math.inf
    if p >= 1.0:
        return math.inf
    return math.log(p / (1.0 - p))


def zeroshot_score(x_bf16: torch.Tensor, prompts_bf16: torch.Tensor, *, pair_mode: bool, temperature: float,
                   thresholds: Optional[Sequence[float]] = None, thr_inclusive: bool = False, normalize_x: bool = True,
                   topk: int = 0, value_mode: int = 0, want_argmax: bool = True, want_mask: bool = True,
                   want_scores: bool = False, guard: Optional[float] = None, count_guard: bool = False,
                   deferred_fixup: bool = True):
    """Scores N embeddings against <=32 prompts.  thresholds are PROBABILITY thresholds (one per label or a scalar
    list); a label passes when sigmoid(score) (>|>=) thr, evaluated exactly as score (>|>=) logit(thr)."""
    require_cuda(x_bf16, prompts_bf16)
    assert x_bf16.dtype == torch.bfloat16 and prompts_bf16.dtype == torch.bfloat16
    x_bf16, prompts_bf16 = x_bf16.contiguous(), prompts_bf16.contiguous()
    n, D = x_bf16.shape
    np_ = prompts_bf16.shape[0]
    L = np_ // 2 if pair_mode else np_
    dev = x_bf16.device
    if guard is None:
        # fp32 (HMMA) accumulation error of a cosine is ~1e-6; re-evaluate in fp64 inside 2e-5 (on the cosine scale)
        guard = 2e-5 / temperature
    thr_arr = None
    if want_mask:
        if thresholds is None:
            thresholds = [0.5] * L
        thresholds = list(thresholds)
        if len(thresholds) == 1:
            thresholds = thresholds * L
        assert len(thresholds) == L
        thr_arr = (C.c_float * L)(*[_logit(float(t)) for t in thresholds])
    argmax = torch.empty((n,), dtype=torch.uint8, device=dev) if want_argmax else None
    mask_u32 = L > 16
    mask = torch.empty((n,), dtype=torch.int32 if mask_u32 else torch.int16, device=dev) if want_mask else None
    tk_idx = torch.empty((n, topk), dtype=torch.uint8, device=dev) if topk > 0 else None
    tk_val = torch.empty((n, topk), dtype=torch.float32, device=dev) if topk > 0 else None
    scores = torch.empty((n, L), dtype=torch.float32, device=dev) if want_scores else None
    gcount = torch.zeros((1,), dtype=torch.int64, device=dev) if count_guard else None
    # flagged rows (decision margin inside the guard band) are listed in this workspace and re-evaluated by a second kernel
    ws = _ws(load().b200clip_zeroshot_workspace_bytes(n), dev) if deferred_fixup else None
    check(load().b200clip_zeroshot_score(ptr(x_bf16), D, n, ptr(prompts_bf16), np_, D, int(pair_mode), int(normalize_x),
                                         float(temperature), thr_arr, int(thr_inclusive), float(guard), topk, value_mode,
                                         ptr(argmax), ptr(mask), int(mask_u32), ptr(tk_idx), ptr(tk_val), ptr(scores),
                                         ptr(gcount), ptr(ws), ws.numel() if ws is not None else 0, stream_ptr()), "zeroshot_score")
    return {"argmax": argmax, "mask": mask, "topk_idx": tk_idx, "topk_val": tk_val, "scores": scores, "guard_rows": gcount}

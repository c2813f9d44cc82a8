"""The fused CLIP head: projections -> (LayerNorm + L2-norm) -> symmetric InfoNCE + multi-label BCE + FC-adapter BCE,
forward and backward, single- or multi-GPU (data-parallel over the batch).

This is the path bench.py times (BASELINE.json: "contrastive head fwd+bwd").  It composes the same C-ABI entry points
as the stand-alone modules/losses, but as ONE autograd node so that no fp32 copies of the normalised embeddings are
materialised between kernels and every collective has a fixed place:

  rank r owns rows [r*B/W, (r+1)*B/W) of images and texts (its local pairs)
  fwd : all_gather(T_hat bf16)  ->  local row block of logits against all columns (never materialised)
        all_reduce(column sum-exp partials [B])      all_reduce(3 loss scalars)
  bwd : dI_hat complete locally;  dT_hat partial [B, D]  ->  reduce_scatter(SUM)
        all_reduce(head parameter gradients)                          
Only T_hat is gathered: the local I rows are the stationary operand of direction 0 and the streamed operand of
direction 1, so I_hat never leaves its rank.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import dp, ops
from .modules import MODEL_CONFIG, ClassificationAdapter, ImageProjection, TextProjection

PARAM_ORDER = ("iw1", "ib1", "iw2", "ib2", "ig", "ibeta", "tw1", "tb1", "tw2", "tb2", "tg", "tbeta", "fw", "fb")


_world, _rank = dp.world, dp.rank

# per-rank batches up to this size run the text-side and image-side kernel chains on two streams (head_forward / head_backward)
TWO_STREAM_MAX_ROWS = 8192
_SIDE_STREAMS = {}


def _side_stream(device, which: int = 0) -> torch.cuda.Stream:
    key = (torch.device(device).index, which)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def head_forward(x_img, x_txt, class_text, labels, tau_nce, tau_bce, group, drop_p, drop_seed, params, need_grad,
                 need_dx=(True, True), defer_loss=False, drop_seed_dev=None):
    """Forward pass of the fused head on this rank's local pairs.  Returns (loss, parts, state); `state` is what
    head_backward needs (None when need_grad is False); parts = [InfoNCE, text BCE, FC BCE].  With defer_loss the loss is
    not finalised here: loss is None and parts is a callable returning (loss, parts, status), to be called after the
    backward kernels are enqueued.  Plain function: ClipHeadFn wraps it for autograd, GraphedHeadStep calls it directly
    while capturing."""
    (iw1, ib1, iw2, ib2, ig, ibeta, tw1, tb1, tw2, tb2, tg, tbeta, fw, fb) = params
    ops.require_cuda(x_img, x_txt, class_text, labels, iw1)
    W, rank = _world(group), _rank(group)
    b_loc = x_img.shape[0]
    b_glob = b_loc * W
    row0 = rank * b_loc
    f = ops._f32c
    labels_f = f(labels)
    # The text side (projection, then the all-gather of T_hat) and the image side (projection; the BCE heads below) are
    # independent until the logits.  With small per-rank batches every kernel of the two chains is a fraction of a wave
    # (a [4096 x 512] GEMM is 64-128 tiles for 148 SMs), so they run on two streams; at large batches each kernel fills the GPU
    # and one stream is used.  Independent dropout streams for the two projections (seed, seed + 1).
    # Two more branches at those sizes: the BCE heads (they need only the image projection) run beside the InfoNCE forward
    # kernel, and the loss VALUE (diagonal + log-sum numerators, their all-reduce) is computed beside the backward pass, which
    # needs only the half-inverse statistics.
    small = b_loc <= TWO_STREAM_MAX_ROWS
    side = _side_stream(x_img.device) if small else None
    heads_side = _side_stream(x_img.device, 2) if small else None
    loss_side = _side_stream(x_img.device, 3)
    main = torch.cuda.current_stream()
    if side is not None:
        side.wait_stream(main)
        heads_side.wait_stream(main)
    with torch.cuda.stream(heads_side if small else main):
        lsum = ops._label_sum(labels_f)
        lsum_work = dp.sum_across_async(lsum, group)      # global label count, needed only by the BCE heads
    with torch.cuda.stream(side if side is not None else main):
        xt = ops.cast_bf16(x_txt)
        tw1b, tw2b = ops.cast_bf16(tw1), ops.cast_bf16(tw2)
        # the fp32 LayerNorm output is never read on this path (InfoNCE and the heads consume y_hat bf16 + 1/||y||): not written
        y_txt, that_loc, inv_txt, saved_t = ops.proj_fwd(xt, tw1b, f(tb1), tw2b, f(tb2), f(tg), f(tbeta), want_yhat=True,
                                                         drop_p=drop_p, drop_seed=drop_seed + 1, drop_seed_dev=drop_seed_dev,
                                                         want_y=False)
        if W > 1 and side is not None:
            # data parallel: the image chain starts when the text chain has finished, so the text chain gets the whole GPU and
            # the all-gather of T_hat (81 us at 8 ranks, RING_LL) starts ~25 us earlier, hidden behind the image chain + heads
            text_done = torch.cuda.Event()
            text_done.record(side)
        that_all, work = dp.gather_rows(that_loc, group, async_op=True)
        if side is not None:
            dp.wait(work)                         # the side stream waits for NCCL; the main stream joins the side stream below
            work = None
    xi = ops.cast_bf16(x_img)
    iw1b, iw2b = ops.cast_bf16(iw1), ops.cast_bf16(iw2)
    if W > 1 and side is not None:
        main.wait_event(text_done)
    C = class_text.shape[0]
    Cf = fw.shape[0]
    fast_heads = ops.heads_mma_supported(iw2.shape[0], C, Cf)          # tensor-core heads read y_hat; the fp32 heads read y
    y_img, ihat, inv_img, saved_i = ops.proj_fwd(xi, iw1b, f(ib1), iw2b, f(ib2), f(ig), f(ibeta), want_yhat=True,
                                                 drop_p=drop_p, drop_seed=drop_seed, drop_seed_dev=drop_seed_dev,
                                                 want_y=not fast_heads)
    sums6 = torch.empty((6,), dtype=torch.float64, device=x_img.device)   # rank-local loss numerators (NCE 3 | BCE 3)
    # one pass over y_img serves both BCE heads, forward AND backward: the input gradient d_bce (for upstream grad 1)
    # and the FC coefficients are produced here; backward only scales them by the incoming gradient.  The heads do not
    # depend on the gathered texts, so they run while the all-gather is still in flight.
    if small:
        heads_side.wait_stream(main)                       # the heads branch: label count (above), then the image features
    with torch.cuda.stream(heads_side if small else main):
        dp.wait(lsum_work)
        if fast_heads:        # tensor-core path: reads the bf16 normalised features LayerNorm wrote for InfoNCE
            d_bce, coef, db_raw = ops.bce_heads_mma(ihat, inv_img, class_text, fw, fb, labels_f, tau_bce, label_sum=lsum,
                                                    total_elems_text=float(b_glob) * C, total_elems_fc=float(b_glob) * Cf,
                                                    sums_out=sums6[3:], want_grad=need_grad)
        else:
            db_raw = None
            d_bce = torch.empty_like(y_img) if need_grad else None
            *_, coef = ops.bce_heads(y_img, class_text, fw, fb, labels_f, tau_bce, label_sum=lsum,
                                     total_elems_text=float(b_glob) * C, total_elems_fc=float(b_glob) * Cf,
                                     dx_out=d_bce, want_coef=need_grad, finalize=False, sums_out=sums6[3:])
    dp.wait(work)
    if side is not None:
        main.wait_stream(side)
    keep = []
    if W == 1:
        # The loss VALUE (diagonal + log-sum numerators) is needed by nothing in the backward pass: it is computed on its own
        # branch beside the backward kernels; `finish` joins the branch and runs the one-thread finalisation.  The MAIN stream
        # joins the heads branch only in head_backward, after the InfoNCE backward launch (d_bce / coef are first read by the
        # image-side chain behind it).
        _, rinvh, cinvh = ops.infonce_forward(ihat, that_all, tau_nce, row0=row0, sums_out=sums6[:3], loss_stream=loss_side,
                                              keep=keep)
        if small:
            loss_side.wait_stream(heads_side)              # the BCE numerators in sums6[3:]
        sums_work = None
        keep.extend((sums6, lsum))
    else:
        # Data parallel: the numerators are summed over ranks, and NCCL runs a communicator's collectives in issue order -- a
        # loss kernel queued behind the backward kernels' CTAs would hold this all-reduce, and with it the reduce-scatter of
        # dT issued after it (measured at 2 ranks: reduce-scatter start 570 us late).  So the loss kernel stays on the main
        # stream in front of the backward pass (6-17 us) and only the all-reduce is asynchronous.
        _, rinvh, cinvh = ops.infonce_forward(ihat, that_all, tau_nce, row0=row0, group=group, sums_out=sums6[:3])
        if small:
            main.wait_stream(heads_side)                   # after the InfoNCE forward is enqueued: the heads run beside it
            heads_side = None
        sums_work = dp.sum_across_async(sums6, group)

    def finish():
        if sums_work is None:
            torch.cuda.current_stream().wait_stream(loss_side)
            keep.clear()
        dp.wait(sums_work)
        return ops.head_loss_finalize(sums6, lsum, tau_nce, b_glob, float(b_glob) * C, float(b_glob) * Cf)

    if defer_loss:
        loss, parts = None, finish
    else:
        loss, parts, _status = finish()
    if not need_grad:
        return loss, parts, None
    tensors = (xi, xt, iw1b, iw2b, tw1b, tw2b, f(ig), f(tg), y_img, y_txt, ihat, that_all, inv_img, inv_txt, rinvh, cinvh,
               d_bce, coef, db_raw, *saved_i, *saved_t)
    meta = dict(tau_nce=tau_nce, group=group, W=W, row0=row0, need_dx=tuple(need_dx), in_dtypes=(x_img.dtype, x_txt.dtype),
                drop=(float(drop_p), int(drop_seed)), drop_seed_dev=drop_seed_dev, has_fc_bias=fb is not None,
                heads_side=heads_side)
    return loss, parts, (tensors, meta)


def head_backward(tensors, meta, g):
    """Backward pass: returns (d_x_img, d_x_txt, [14 parameter gradients in PARAM_ORDER])."""
    (xi, xt, iw1b, iw2b, tw1b, tw2b, ig, tg, y_img, y_txt, ihat, that_all, inv_img, inv_txt, rinvh, cinvh, d_bce, coef,
     db_raw, *rest) = tensors
    saved_i, saved_t = tuple(rest[:5]), tuple(rest[5:])
    group, row0 = meta["group"], meta["row0"]
    need_dxi, need_dxt = meta["need_dx"]
    drop_p, drop_seed = meta["drop"]
    seed_dev = meta.get("drop_seed_dev")
    g = ops._f32c(g).reshape(())
    W = meta["W"]
    b_loc = ihat.shape[0]
    main = torch.cuda.current_stream()
    side = _side_stream(ihat.device) if (b_loc <= TWO_STREAM_MAX_ROWS or W > 1) else None
    nce_side = None
    if W > 1:
        # The two directions are separate launches on two streams: direction 1 (dT partial, all B rows of T against the local
        # columns) is enqueued first and its reduce-scatter (58 MB per rank at B = 32768, W = 8) starts the moment it ends;
        # direction 0 (dI) runs concurrently and fills the SMs direction 1 leaves idle in its last wave (launched back to back
        # on ONE stream the two kernels cost 16 % more than the combined launch: 256 long CTAs are 1.7 waves of 148 SMs).
        nce_side = _side_stream(ihat.device, 1)
        nce_side.wait_stream(main)
        _, d_that = ops.infonce_backward(ihat, that_all, meta["tau_nce"], rinvh, cinvh, g, row0=row0, directions=2)
        d_that_loc, work = dp.scatter_sum_rows(d_that, group, async_op=True)
        with torch.cuda.stream(nce_side):
            d_ihat, _ = ops.infonce_backward(ihat, that_all, meta["tau_nce"], rinvh, cinvh, g, row0=row0, allow_splits=True,
                                             directions=1)
    else:
        d_ihat, d_that = ops.infonce_backward(ihat, that_all, meta["tau_nce"], rinvh, cinvh, g, row0=row0, allow_splits=True)
        d_that_loc, work = d_that, None
    that_loc = that_all[row0:row0 + b_loc]                           # this rank's normalised text rows (bf16)
    if meta.get("heads_side") is not None:
        main.wait_stream(meta["heads_side"])         # the BCE-heads branch of the forward pass (d_bce, coef, db_raw)
    # text side (needs only the reduce-scattered dT) on the side stream, image side (needs dI) on the main stream: two chains of
    # ~12 kernels that are each a fraction of a wave at per-rank batch sizes
    if side is not None:
        side.wait_stream(main)                       # forks after the dT kernel / reduce-scatter enqueue, NOT after dI
    with torch.cuda.stream(side if side is not None else main):
        if side is not None:
            dp.wait(work)
            gt = ops.proj_bwd(None, xt, tw1b, tw2b, tg, saved_t, need_dxt, meta["in_dtypes"][1], drop_p=drop_p, drop_seed=drop_seed + 1,
                              l2=(d_that_loc, that_loc, inv_txt, None, None), drop_seed_dev=seed_dev)
            if W > 1:
                # The text chain starts when the reduce-scatter lands, i.e. while dI is still running, and ends ~60 us before
                # the image chain: its gradient bucket travels behind the image chain, only the image bucket is exposed.
                txt_grads = dp.allreduce_flat([gt[1], gt[2], gt[3], gt[4], gt[5], gt[6]], group)
    if nce_side is not None:
        main.wait_stream(nce_side)
    # image side: the L2-normalisation backward (+ g * the two BCE heads' input gradient from the forward pass) runs inside
    # the projection block's LayerNorm-backward kernel
    if db_raw is not None:
        dfw, dfb = ops.skinny_outer_mma(coef, ihat, db_raw, out_scale=g)
    else:
        dfw, dfb = ops.skinny_outer(coef, y_img, want_bias=True, out_scale=g)
    gi = ops.proj_bwd(None, xi, iw1b, iw2b, ig, saved_i, need_dxi, meta["in_dtypes"][0], drop_p=drop_p, drop_seed=drop_seed,
                      l2=(d_ihat, ihat, inv_img, d_bce, g), drop_seed_dev=seed_dev)
    # parameter gradients: SUM, not mean (every loss term is normalised by the GLOBAL batch)
    if side is None:
        # one stream: the image-side bucket travels while the text side computes
        img_grads, img_work = dp.allreduce_flat([gi[1], gi[2], gi[3], gi[4], gi[5], gi[6], dfw, dfb], group, async_op=True)
        dp.wait(work)
        gt = ops.proj_bwd(None, xt, tw1b, tw2b, tg, saved_t, need_dxt, meta["in_dtypes"][1], drop_p=drop_p, drop_seed=drop_seed + 1,
                          l2=(d_that_loc, that_loc, inv_txt, None, None), drop_seed_dev=seed_dev)
        txt_grads = dp.allreduce_flat([gt[1], gt[2], gt[3], gt[4], gt[5], gt[6]], group)
        dp.wait(img_work)
    else:
        img_grads = dp.allreduce_flat([gi[1], gi[2], gi[3], gi[4], gi[5], gi[6], dfw, dfb], group)
        main.wait_stream(side)
        if W == 1:
            txt_grads = [gt[1], gt[2], gt[3], gt[4], gt[5], gt[6]]
    grads = [*img_grads[:6], *txt_grads, img_grads[6], img_grads[7]]
    if not meta["has_fc_bias"]:
        grads[-1] = None
    return gi[0], gt[0], grads


class ClipHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_img, x_txt, class_text, labels, tau_nce, tau_bce, group, drop_p, drop_seed, *params):
        if class_text.requires_grad:
            # the reference computes the class prompts under no_grad (0426/train.py:333); a silent None gradient would be a bug
            raise RuntimeError("b200clip.ClipHead: class_text_features must not require grad (compute them under torch.no_grad() "
                               "as 0426/train.py:333 does, or use b200clip.multilabel_contrastive_loss, which returns d text)")
        need_grad = any(t is not None and t.requires_grad for t in (x_img, x_txt, *params))
        loss, parts, state = head_forward(x_img, x_txt, class_text, labels, tau_nce, tau_bce, group, drop_p, drop_seed, params,
                                          need_grad, need_dx=(x_img.requires_grad, x_txt.requires_grad))
        if state is not None:
            ctx.save_for_backward(*state[0])
            ctx.meta = state[1]
        ctx.parts = parts
        return loss

    @staticmethod
    def backward(ctx, g):
        dxi, dxt, grads = head_backward(ctx.saved_tensors, ctx.meta, g)
        return (dxi, dxt, None, None, None, None, None, None, None, *grads)


class ClipHead(nn.Module):
    """ImageProjection + TextProjection + ClassificationAdapter with the fused step.  Sub-modules keep the reference's
    state_dict keys (image_projector.*, text_projector.* as in the reference's `models` dict, 0426/train.py:888-928)."""

    def __init__(self, image_embedding_size=MODEL_CONFIG["image_embedding_size"],
                 text_embedding_size=MODEL_CONFIG["text_embedding_size"],
                 shared_embedding_size=MODEL_CONFIG["shared_embedding_size"], num_labels=MODEL_CONFIG["num_labels"],
                 tau_nce: float = MODEL_CONFIG["temperature"], tau_bce: float = 1.0, group=None,
                 dropout_rate: float = 0.0):
        super().__init__()
        self.image_projector = ImageProjection(image_embedding_size, shared_embedding_size, dropout_rate=dropout_rate)
        self.text_projector = TextProjection(text_embedding_size, shared_embedding_size, dropout_rate=dropout_rate)
        self.dropout_rate = dropout_rate
        self.classifier = ClassificationAdapter(shared_embedding_size, num_labels)
        self.tau_nce, self.tau_bce, self.group = tau_nce, tau_bce, group

    def params(self):
        ip, tp, c = self.image_projector, self.text_projector, self.classifier
        return (ip.image_projection.weight, ip.image_projection.bias, ip.fc.weight, ip.fc.bias, ip.layer_norm.weight,
                ip.layer_norm.bias, tp.text_projection.weight, tp.text_projection.bias, tp.fc.weight, tp.fc.bias,
                tp.layer_norm.weight, tp.layer_norm.bias, c.weight, c.bias)

    def forward(self, image_embeddings, text_embeddings, class_text_features, labels):
        """One head step on this rank's local pairs; returns the global loss
        contrastive_loss(tau_nce) + multilabel_contrastive_loss(tau_bce) + BCEWithLogits(classifier)."""
        p = float(self.dropout_rate) if self.training else 0.0
        seed = ops.new_dropout_seed() if p > 0 else 0
        self.last_dropout_seed = seed
        return ClipHeadFn.apply(image_embeddings, text_embeddings, class_text_features, labels, self.tau_nce, self.tau_bce,
                                self.group, p, seed, *self.params())


class GraphedHeadStep:
    """One head step (forward + backward, collectives included) captured ONCE into a CUDA graph and replayed.

    At per-rank batches of a few thousand pairs the step is ~40 kernels of 5-100 us each: launched one by one from Python
    the host, not the GPU, sets the step time (~0.9 ms of enqueue work per step, tools/host_overhead.py).  Replaying the
    captured graph costs one launch.  Shapes are fixed at capture; inputs are copied into static device buffers (host or
    device tensors accepted), gradients land in the parameters' .grad and in `grad_image` / `grad_text`:

        step = GraphedHeadStep(head, x_img, x_txt, class_text, labels)
        loss = step(x_img, x_txt, class_text, labels)      # same semantics as head(...) followed by loss.backward()

    Dropout (head.train() with dropout_rate > 0, the reference's nn.Dropout(0.1), 0426/train.py:81,93) is captured too: the
    keep-mask seed lives in a device word (`seed_dev`) that a one-thread kernel at the head of the graph advances on every
    replay; the projection kernels add it to their (frozen) seed argument.

    Parameter gradients are OVERWRITTEN by each replay (equivalent to zero_grad + backward); every call re-binds the
    parameters' .grad to the graph's output tensors, so `opt.zero_grad(); step(...); opt.step()` works.  The returned loss
    ALIASES a static buffer that the next replay overwrites -- `.clone()` it to keep a history.
    """

    def __init__(self, head: "ClipHead", image_embeddings, text_embeddings, class_text_features, labels, warmup: int = 3,
                 input_grads: bool = True):
        ops.require_cuda(image_embeddings, text_embeddings, class_text_features, labels)
        self.head = head
        self.drop_p = float(head.dropout_rate) if head.training else 0.0
        self.drop_seed0 = ops.new_dropout_seed() if self.drop_p > 0 else 0
        # device-resident seed word, advanced by the first node of the graph (read it after a replay to reconstruct the mask:
        # effective seeds are drop_seed0 + seed_dev for the image projection, drop_seed0 + 1 + seed_dev for the text projection)
        self.seed_dev = torch.zeros((), dtype=torch.int32, device=image_embeddings.device) if self.drop_p > 0 else None
        self.x_img = image_embeddings.detach().clone()
        self.x_txt = text_embeddings.detach().clone()
        self.class_text = class_text_features.detach().clone()
        self.labels = labels.detach().clone()
        self.input_grads = input_grads
        self._one = torch.ones((), dtype=torch.float32, device=self.x_img.device)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.x_img.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):               # lazy init (function attributes, NCCL channels) outside capture
                self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.loss, self.grad_image, self.grad_text, grads = self._run()
        self._param_grads = list(zip(self.head.params(), grads))
        self.bind_grads()

    @torch.no_grad()
    def _run(self):
        """forward + backward as plain calls (no autograd engine: its worker thread and AccumulateGrad streams do not
        belong in a capture)."""
        h = self.head
        if self.seed_dev is not None:
            ops.dropout_seed_advance(self.seed_dev)
        _, finish, (tensors, meta) = head_forward(self.x_img, self.x_txt, self.class_text, self.labels, h.tau_nce, h.tau_bce,
                                                  h.group, self.drop_p, self.drop_seed0, h.params(), True,
                                                  need_dx=(self.input_grads, self.input_grads), defer_loss=True,
                                                  drop_seed_dev=self.seed_dev)
        dxi, dxt, grads = head_backward(tensors, meta, self._one)
        loss, self.parts, self.status = finish()          # loss value: after the backward kernels, off the critical path
        return loss, dxi, dxt, grads

    def bind_grads(self):
        """Point every parameter's .grad at the tensor the graph writes (again, if something re-bound .grad since the
        capture: optimizer.zero_grad(set_to_none=True) or an eager backward)."""
        for p, g in self._param_grads:
            if p is not None:
                p.grad = g

    def close(self):
        """Release the captured graph.  Call before torch.distributed.destroy_process_group(): NCCL waits for every graph
        that holds its kernels, so tearing the communicator down with a live graph hangs."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def __call__(self, image_embeddings=None, text_embeddings=None, class_text_features=None, labels=None):
        if self.graph is None:
            raise RuntimeError("GraphedHeadStep: the graph was released by close()")
        with torch.no_grad():
            for name, dst, src in (("image_embeddings", self.x_img, image_embeddings), ("text_embeddings", self.x_txt, text_embeddings),
                                   ("class_text_features", self.class_text, class_text_features), ("labels", self.labels, labels)):
                if src is None or src.data_ptr() == dst.data_ptr():
                    continue
                if tuple(src.shape) != tuple(dst.shape):
                    raise RuntimeError(f"GraphedHeadStep: {name} has shape {tuple(src.shape)}, the graph was captured for "
                                       f"{tuple(dst.shape)} (capture a new GraphedHeadStep for a different batch size)")
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.bind_grads()               # optimizer.zero_grad(set_to_none=True) between steps must not detach the graph's outputs
        return self.loss

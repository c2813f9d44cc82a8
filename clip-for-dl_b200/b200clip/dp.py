"""Data-parallel choreography of the CLIP head: the collectives and the algebra that combines
per-rank partial results.  Device-agnostic on purpose: the same functions run over NCCL on B200s (head.py, ops.py) and
over gloo on CPU tensors in tests/test_dp_gloo.py, where the test supplies the per-rank arithmetic.

  rank r owns rows [r*B/W, (r+1)*B/W) of images and texts
  fwd: gather_rows(T_hat)  ->  local row block of logits vs all columns  ->  sum_across(column sum-exp partials)
       -> sum_across(3 loss scalars)
  bwd: dI_hat complete locally; dT_hat partial [B, D] -> scatter_sum_rows ; allreduce_flat(parameter gradients)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class _Solo:
    """Sentinel process group: "this rank alone" (world 1) even when torch.distributed is initialised -- lets a rank run the
    single-GPU step next to the data-parallel one (bench.py's parity leg) without creating single-rank NCCL communicators."""

    def __repr__(self):
        return "dp.SOLO"


SOLO = _Solo()


def world(group=None) -> int:
    if group is SOLO:
        return 1
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def rank(group=None) -> int:
    if group is SOLO:
        return 0
    return dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0


def init_process_group(device, **kw):
    """torch.distributed.init_process_group("nccl") with HIGH-PRIORITY NCCL streams.  The head's kernels fill every SM (one
    CTA of ~224 KB shared memory per SM), so a collective enqueued next to them only gets SMs as compute CTAs retire; with
    default priorities its CTAs queue BEHIND the compute kernel's pending CTAs and the collective starts when that kernel drains
    (measured at 8 ranks: the reduce-scatter launched after the dT kernel started 190 us late, after the dI kernel's last wave).
    High-priority streams let the block scheduler place the collective's CTAs first."""
    opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    return dist.init_process_group("nccl", device_id=device, pg_options=opts, **kw)


def gather_rows(local: torch.Tensor, group=None, async_op: bool = False):
    """all_gather along dim 0 (contiguous rank order).  Returns (full, work-or-None)."""
    W = world(group)
    if W == 1:
        return local, None
    full = torch.empty((local.shape[0] * W, *local.shape[1:]), dtype=local.dtype, device=local.device)
    work = dist.all_gather_into_tensor(full, local.contiguous(), group=group, async_op=async_op)
    return full, (work if async_op else None)


def sum_across(t: torch.Tensor, group=None) -> torch.Tensor:
    if world(group) > 1:
        dist.all_reduce(t, group=group)
    return t


def sum_across_async(t: torch.Tensor, group=None):
    """all_reduce(SUM) in place without making the compute stream wait; returns a work handle (None at world 1)."""
    if world(group) > 1:
        return dist.all_reduce(t, group=group, async_op=True)
    return None


def wait(work) -> None:
    if work is not None:
        work.wait()


def scatter_sum_rows(full: torch.Tensor, group=None, async_op: bool = False):
    """reduce_scatter(SUM) along dim 0: every rank contributes a [B, ...] partial and keeps its own row block."""
    W = world(group)
    if W == 1:
        return full, None
    n = full.shape[0] // W
    if dist.get_backend(group) == "gloo":                 # gloo has no reduce_scatter: all_reduce + slice (tests only)
        dist.all_reduce(full, group=group)
        r = rank(group)
        return full[r * n:(r + 1) * n].contiguous(), None
    out = torch.empty((n, *full.shape[1:]), dtype=full.dtype, device=full.device)
    work = dist.reduce_scatter_tensor(out, full.contiguous(), group=group, async_op=async_op)
    return out, (work if async_op else None)


def allreduce_flat(tensors: Sequence[torch.Tensor], group=None, async_op: bool = False):
    """One bucket of parameter gradients: SUM is exact because every loss term is already normalised by the GLOBAL
    batch.  Returns the reduced tensors (views of the bucket); with async_op=True returns (tensors, work)."""
    if world(group) == 1:
        return (list(tensors), None) if async_op else list(tensors)
    flat = torch.cat([t.reshape(-1) for t in tensors])
    work = dist.all_reduce(flat, group=group, async_op=async_op)
    out, o = [], 0
    for t in tensors:
        out.append(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
    return (out, work) if async_op else out


def infonce_loss_from_sums(sums: torch.Tensor, temperature: float, b_glob: int) -> torch.Tensor:
    """sums = [sum_i log r_i, sum_j log c_j, sum_i S_ii] (already summed over ranks); m = the kernels' fixed shift
    (b200clip_infonce_shift: 1/tau, lowered for tau < 0.036 so that exp2 cannot flush whole rows)."""
    return (infonce_shift(temperature) + (sums[0] + sums[1]) / (2.0 * b_glob) - sums[2] / b_glob).to(torch.float32)


def infonce_shift(temperature: float) -> float:
    """The shift m, restated for device-agnostic callers (the gloo tests): mirrors csrc/host.cuh nce_shift in float32."""
    import numpy as np
    log2e = np.float32(1.4426950408889634)
    k1 = log2e / np.float32(temperature)
    off = np.float32(min(max(float(k1) - 40.0, 0.0), 88.0))
    return float(np.float64(np.float32(k1 - off)) / np.float64(log2e))


def bce_losses_from_sums(sums: torch.Tensor, label_sum: torch.Tensor, total_text: float, total_fc: float):
    """sums = [text pos numerator, text neg numerator, FC BCE sum] summed over ranks (0426/train.py:218-221)."""
    P = label_sum.double()
    N = total_text - P
    l_text = (0.5 * (-sums[0] / (P + 1e-8) - sums[1] / (N + 1e-8))).float()
    l_fc = (sums[2] / total_fc).float()
    return l_text, l_fc

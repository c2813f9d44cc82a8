#!/usr/bin/env python
"""bench.py -- the CLIP contrastive head fwd+bwd benchmark (BASELINE.json metric), one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W            our arm (B200 kernels through the public b200clip API)
  python bench.py --impl reference ...                      the reference's CPU path (oracle port) on the host cores
  python bench.py --workload zeroshot                       cfg 4 (1M x 28 prompts zero-shot scoring), HBM roofline

Workload (head): BASELINE.json configs[2] at the size the metric is quoted on -- global batch B=32768, D=512,
E_img=E_txt=768 (ViT-B/16 + Bio_ClinicalBERT widths), C=16 labels, tau 0.07 (InfoNCE) / 1.0 (BCE); the global batch is
FIXED as N grows (strong scaling), rank r owns B/N pairs.  One step = projections + LayerNorm/L2 + symmetric InfoNCE
+ multi-label BCE + FC-adapter BCE, forward and backward (all parameter and input gradients), synthetic data.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "clip-for-dl_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

CFG = {
    "cfg3": dict(B=32768, D=512, E_img=768, E_txt=768, C=16, name="cfg3: ViT-width CLIP head, global batch 32768, D=512"),
    "cfg2": dict(B=4096, D=512, E_img=2048, E_txt=768, C=16, name="cfg2: ResNet-width CLIP head, batch 4096, D=512"),
}
TAU_NCE, TAU_BCE = 0.07, 1.0
# DRAM bytes of ONE launch from the ncu --set full captures under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum)
NCE_BWD_KERNEL = "nce_bwdc_kernel<2>"
NCE_BWD_TRAFFIC_BYTES = 264.24e6          # profiles/r2_ncu_nce_B32768.txt: 156.66 MB read + 107.58 MB written (B = 32768, D = 512, one GPU)
ZS_TRAFFIC_BYTES = 1.0649e9               # profiles/r2_ncu_zeroshot_16row.txt: 1.0313 GB read + 33.6 MB written (N = 1M, 28 prompts)
DROPOUT = 0.1            # nn.Dropout(0.1) of the projections (0426/config.py:27), ON in the timed step as in training


def head_flops(B, D, E_img, E_txt, C):
    return 6.0 * B * B * D + 6.0 * B * (E_img * D + D * D) + 6.0 * B * (E_txt * D + D * D) + 12.0 * B * D * C


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=d.get("bf16_tflops_sustained", 1410.9), tflops_burst=d.get("bf16_tflops", 1687.9),
                    hbm=d.get("hbm_gbs", 6546.9), src="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with NVML while the timed region runs (NVML is initialised in the constructor, so the
    first sample lands at the start of the region; one sample every 5 ms)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self._nv = self._h = self._get = None
        self._names = {}
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self._names = {
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            self._get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception as e:  # NVML missing: report that instead of failing the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        if self._nv is None:
            return
        try:
            while not self._stop_evt.is_set():
                self.samples.append(self._nv.nvmlDeviceGetClockInfo(self._h, self._nv.NVML_CLOCK_SM))
                mask = self._get(self._h)
                for bit, nm in self._names.items():
                    if mask & bit:
                        self.reasons.add(nm)
                time.sleep(0.005)
        except Exception as e:
            self.reasons.add(f"nvml_error:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


SYNTH_BLOCK = 512        # rows per seeded block: rank slices are whole blocks for every N in {1, 2, 4, 8}


def synth_rows(cfg, row0, nrows):
    """Rows [row0, row0 + nrows) of the GLOBAL synthetic batch.  Block b (SYNTH_BLOCK rows) is drawn from seed 1234 + b, so the
    global batch -- and therefore the loss -- is the same for every rank count N; a rank only generates its own blocks."""
    assert row0 % SYNTH_BLOCK == 0 and nrows % SYNTH_BLOCK == 0, "per-rank batches must be multiples of SYNTH_BLOCK"
    xi, xt, lab = [], [], []
    for b in range(row0 // SYNTH_BLOCK, (row0 + nrows) // SYNTH_BLOCK):
        g = torch.Generator().manual_seed(1234 + b)
        xi.append(torch.randn(SYNTH_BLOCK, cfg["E_img"], generator=g).to(torch.bfloat16))
        xt.append(torch.randn(SYNTH_BLOCK, cfg["E_txt"], generator=g).to(torch.bfloat16))
        lab.append((torch.rand(SYNTH_BLOCK, cfg["C"], generator=g) < 0.0524).float())
    return torch.cat(xi), torch.cat(xt), torch.cat(lab)


def synth_class_text(cfg):
    gc = torch.Generator().manual_seed(99)
    return torch.nn.functional.normalize(torch.randn(cfg["C"], cfg["D"], generator=gc), dim=1)


def synth_inputs(cfg, b_loc, rank, device, pinned=False):
    x_img, x_txt, labels = synth_rows(cfg, rank * b_loc, b_loc)
    class_text = synth_class_text(cfg)
    if pinned:
        return x_img.pin_memory(), x_txt.pin_memory(), labels.pin_memory(), class_text.to(device)
    return x_img.to(device), x_txt.to(device), labels.to(device), class_text.to(device)


def head_config(cfg, world, eager=False):
    """`config` of the JSON line -- shared by both arms (the reference arm reports on OUR arm's config)."""
    B = cfg["B"]
    return {"workload": cfg["name"], "global_batch": B, "per_gpu_batch": B // world, "D": cfg["D"], "E_img": cfg["E_img"],
            "E_txt": cfg["E_txt"], "C": cfg["C"], "tau": [TAU_NCE, TAU_BCE], "dropout": DROPOUT, "parallelism": f"dp{world}",
            "launch_mode": "eager (kernel by kernel)" if eager else "one CUDA graph per step (GraphedHeadStep)",
            "l2": "no explicit flush: per-step working set (activations + fp32 grads, >400 MB at B=32768) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference head on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_head_step_time(cfg, B, reps, warm):
    """Seconds per fwd+bwd step of the oracle port of the reference head (torch fp32, all host threads) on the first B rows of the
    bench's own global batch (dropout off: the port, like the parity tests, has no dropout)."""
    import ref_head as R
    import synth
    torch.set_num_threads(os.cpu_count() or 1)
    D, C = cfg["D"], cfg["C"]
    ip = {k: v.clone().requires_grad_(True) for k, v in synth.projection_params(100, cfg["E_img"], D).items()}
    tp = {k: v.clone().requires_grad_(True) for k, v in synth.projection_params(200, cfg["E_txt"], D).items()}
    fw = synth.uniform(31, -0.04, 0.04, C, D).requires_grad_(True)
    fb = synth.uniform(32, -0.04, 0.04, C).requires_grad_(True)
    xi, xt, labels = synth_rows(cfg, 0, B)
    x_img, x_txt = xi.float().requires_grad_(True), xt.float().requires_grad_(True)
    class_text = synth_class_text(cfg)
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        out = R.head_step(x_img, x_txt, class_text, labels, ip, tp, fw, fb, TAU_NCE, TAU_BCE)
        out["loss"].backward()
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
        for t in (x_img, x_txt, fw, fb, *ip.values(), *tp.values()):
            t.grad = None
    return times


def host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().total / 2 ** 30
    except Exception:
        return 0.0


def pick_cpu_sample(cfg, n_steps, budget_s):
    """Largest batch B in {4096, 8192, 16384, full} whose n_steps steps fit the time budget (CPU time grows ~B^2: fp32 B x B
    logits; ~5 live copies need 5 * 4 * B^2 bytes of RAM) -- calibrated with one step at B = 4096."""
    full = cfg["B"]
    t4096 = min(cpu_head_step_time(cfg, min(4096, full), reps=1, warm=1))
    best = min(4096, full)
    for B in (8192, 16384, 32768, 65536):
        if B > full:
            break
        est = t4096 * (B / 4096.0) ** 2
        ram_need = 6 * 4 * B * B / 2 ** 30 + 4
        if est * n_steps <= budget_s and ram_need <= 0.8 * host_ram_gb():
            best = B
    return best, t4096


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    Bs, _ = pick_cpu_sample(cfg, args.steps + args.warmup, budget_s=150.0)
    times = cpu_head_step_time(cfg, Bs, reps=args.steps, warm=args.warmup)
    ms = 1e3 * sum(times) / len(times)
    val = Bs / (ms / 1e3)
    full = Bs == cfg["B"]
    sample = (f"oracle port of the reference head (oracle/ref_head.py, pinned to the unmodified reference by tests/golden; torch fp32 "
              f"CPU, {cores} threads), each step = fwd+bwd on " + ("the full workload" if full else
              f"a bounded sample: the first B={Bs} rows of the B={cfg['B']} global batch; CPU time grows ~B^2, so at the full "
              f"batch the CPU would deliver ~{val * Bs / cfg['B']:.0f} pairs/s (extrapolated_full_config)"))
    line = {
        "impl": "reference", "metric": "clip_head_fwd_bwd_pairs_per_sec", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": head_config(cfg, max(world, 1)),
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample, "sample_batch": Bs,
                         "extrapolated_full_config": val * Bs / cfg["B"]},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def time_fwd_kernel(lib, b_loc, B, D, dev, reps=5):
    """Stand-alone launches of the InfoNCE forward (statistics) kernel at the step's shape, CUDA events, after a flush."""
    import b200clip  # noqa: F401
    from b200clip import _lib
    g = torch.Generator().manual_seed(7)
    T = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(torch.bfloat16)
    I = T[:b_loc].roll(1, 0).contiguous()
    nb = lib.b200clip_infonce_workspace_bytes(b_loc, B)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    r, c = torch.empty(b_loc, device=dev), torch.empty(B, device=dev)
    ts = []
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.b200clip_infonce_fwd_stats(_lib.ptr(I), _lib.ptr(T), D, b_loc, B, TAU_NCE, _lib.ptr(r), _lib.ptr(c), _lib.ptr(ws), nb,
                                                  _lib.stream_ptr()), "fwd")
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def measure_head(args, cfg, rank, world, dev, steps, warmup, eager=False, want_e2e=True, kernel_events="bwd"):
    """Times the head step of `cfg` on this rank (all ranks call it together).  Returns the measurement dict of this workload:
    device-timed value with inputs resident in HBM, the e2e figure fed from pinned host memory, kernel timings, launch count."""
    import b200clip
    from b200clip import _lib, ops
    lib = _lib.load()
    B = cfg["B"]
    assert B % world == 0
    b_loc = B // world
    torch.manual_seed(0)
    head = b200clip.ClipHead(cfg["E_img"], cfg["E_txt"], cfg["D"], cfg["C"], TAU_NCE, TAU_BCE, dropout_rate=DROPOUT).to(dev)
    head.train()                                   # nn.Dropout(0.1) active, as in the reference's training step
    x_img, x_txt, labels, class_text = synth_inputs(cfg, b_loc, rank, dev)
    x_img.requires_grad_(True)
    x_txt.requires_grad_(True)

    def eager_step(xi, xt, lab):
        for p in head.parameters():
            p.grad = None
        xi.grad = None
        xt.grad = None
        loss = head(xi, xt, class_text, lab)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The step runs as ONE CUDA graph (b200clip.GraphedHeadStep, public API): ~35 kernels + collectives per step would
    # otherwise be enqueued one by one from Python (~0.9 ms of host work per step, more than the GPU needs at N=8).
    # --eager times the same step launched kernel by kernel.  The event pairs around the two InfoNCE kernels are captured
    # as external event-record nodes, so every replay re-records them on the launching stream.  Such a node is not free: four of
    # them cost 47 us per replay at B = 4096 (0.332 vs 0.285 ms, tools/cfg2_probe.py) -- they cut the graph into segments.  So
    # the graph of the headline run carries only the pair around the DOMINANT kernel (the roofline leg); the forward kernel is
    # timed stand-alone after the region, and the cfg2 sub-record, which reports no per-kernel time, carries none.
    if eager:
        kernel_events = "both"
    ops.KERNEL_EVENTS["infonce_bwd"] = [] if kernel_events != "none" else None
    ops.KERNEL_EVENTS["infonce_fwd"] = [] if kernel_events == "both" else None
    if eager:
        def step(xi=None, xt=None, lab=None):
            return eager_step(x_img if xi is None else xi, x_txt if xt is None else xt, labels if lab is None else lab)
    else:
        gstep = b200clip.GraphedHeadStep(head, x_img, x_txt, class_text, labels, warmup=3)

        def step(xi=None, xt=None, lab=None):
            return gstep(xi, xt, None, lab)
    # Every event created while the graph was captured is referenced by one of its event-record nodes: they must stay alive as
    # long as the graph does.  A data-parallel step launches the backward kernel twice (dT direction, then dI): its time is the
    # sum of the two launches.
    n_bwd = 2 if world > 1 else 1
    keep_alive = (list(ops.KERNEL_EVENTS["infonce_bwd"] or []), list(ops.KERNEL_EVENTS["infonce_fwd"] or []))
    graph_events = (keep_alive[0][-n_bwd:], keep_alive[1][-1:])
    ops.KERNEL_EVENTS["infonce_bwd"] = ops.KERNEL_EVENTS["infonce_fwd"] = None

    for _ in range(max(warmup, 3)):
        loss = step()
    barrier()
    # ---- timed region: exactly K steps, device-timed, inputs resident in HBM ------------------------------------
    sampler = ClockSampler(dev.index)
    sampler.start()
    if eager:
        ops.KERNEL_EVENTS["infonce_bwd"], ops.KERNEL_EVENTS["infonce_fwd"] = [], []
    n0 = lib.b200clip_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if eager:
        launches = (lib.b200clip_launch_count() - n0) // steps
        ev = ops.KERNEL_EVENTS["infonce_bwd"]
        bwd_ms = [max(ev[i][0].elapsed_time(e1) for _, e1 in ev[i:i + n_bwd]) for i in range(0, len(ev), n_bwd)]
        fwd_ms = [a.elapsed_time(b) for a, b in ops.KERNEL_EVENTS["infonce_fwd"]]
        ops.KERNEL_EVENTS["infonce_bwd"] = ops.KERNEL_EVENTS["infonce_fwd"] = None
        kernel_timing = "CUDA events around every launch inside the timed region"
    else:
        # kernels per replay = kernels the library launched while the graph was captured (counted once, below)
        # data parallel: the two direction launches run concurrently on two streams -> span from the first start to the last end
        span = lambda evs: max(evs[0][0].elapsed_time(e1) for _, e1 in evs)
        bwd_ms, fwd_ms = [], []
        kernel_timing = "none (sub-record)"
        if graph_events[0]:
            last_bwd = [span(graph_events[0])]                            # the last timed step's launch(es)
            for _ in range(steps):                       # same replay, read back step by step (sync between steps)
                step()
                torch.cuda.synchronize()
                bwd_ms.append(span(graph_events[0]))
                fwd_ms += [a.elapsed_time(b) for a, b in graph_events[1]]
            kernel_timing = (f"external event-record nodes around the launch inside the step graph; mean of {steps} replays "
                             f"read back one by one right after the timed region (last timed step: {last_bwd[0]:.3f} ms)")
            if not graph_events[1]:
                fwd_ms = [time_fwd_kernel(lib, b_loc, B, cfg["D"], dev)]
                kernel_timing += "; fwd_kernel_ms: stand-alone launches after the region"
        n1 = lib.b200clip_launch_count()
        eager_step(x_img, x_txt, labels)                 # one eager step = the launches one replay contains (+1: seed advance)
        launches = lib.b200clip_launch_count() - n1 + (1 if DROPOUT > 0 else 0)
        gstep.bind_grads()
        torch.cuda.synchronize()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / steps
    out = {"B": B, "b_loc": b_loc, "ms_step": ms_step, "value": B / (ms_step / 1e3), "loss": float(loss.item()),
           "launches": int(launches), "clocks": clocks, "bwd_ms": sum(bwd_ms) / max(len(bwd_ms), 1),
           "fwd_ms": sum(fwd_ms) / max(len(fwd_ms), 1), "kernel_timing": kernel_timing}

    if want_e2e:
        # ---- e2e: same step through the public API with HOST buffers (pinned H2D in, loss D2H out, every step) ----------
        # Input pipeline as a training loop runs it: while step i computes, a copy stream moves step i+1's batch from pinned
        # host memory into the other of two device buffer sets.  Every step's H2D copy and its loss read-back are inside the
        # timed region.  The loss of EVERY step is copied to pinned host memory and read by the host, one step late: the host
        # enqueues step i, then waits for step i-1's loss (an event, not a stream synchronize), so the GPU never idles while
        # the host turns around (a logging loop that does not need the value before launching the next step).
        hx_img, hx_txt, hlab, _ = synth_inputs(cfg, b_loc, rank, dev, pinned=True)
        bufs = []
        for _ in range(2):
            bufs.append((torch.empty_like(hx_img, device=dev).requires_grad_(True),
                         torch.empty_like(hx_txt, device=dev).requires_grad_(True), torch.empty_like(hlab, device=dev)))
        host_loss = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
        seen = []
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(slot):
            with torch.cuda.stream(copy_stream), torch.no_grad():
                copy_stream.wait_event(consumed[slot])          # the step that last used this buffer set has finished
                d_img, d_txt, d_lab = bufs[slot]
                d_img.copy_(hx_img, non_blocking=True)
                d_txt.copy_(hx_txt, non_blocking=True)
                d_lab.copy_(hlab, non_blocking=True)
                copied[slot].record(copy_stream)

        def e2e_step(i):
            slot = i & 1
            cur = torch.cuda.current_stream()
            cur.wait_event(copied[slot])
            d_img, d_txt, d_lab = bufs[slot]
            l = step(d_img, d_txt, d_lab)                      # graph mode: D2D into the static inputs, then one replay
            consumed[slot].record(cur)
            host_loss[slot].copy_(l.detach(), non_blocking=True)
            loss_ready[slot].record(cur)
            prefetch(slot ^ 1)                                  # next step's batch: enqueued while this step computes
            if i > 0:
                read_loss(slot ^ 1)                             # step i-1's loss, while step i runs

        def read_loss(slot):
            loss_ready[slot].synchronize()
            seen.append(float(host_loss[slot]))

        for slot in range(2):
            consumed[slot].record(torch.cuda.current_stream())
        prefetch(0)
        for i in range(4):
            e2e_step(i)
        read_loss(3 & 1)
        barrier()
        seen.clear()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(steps):
            e2e_step(i)                                         # `steps` H2D batches are copied inside the region
        read_loss((steps - 1) & 1)                              # the last step's loss: inside the region too
        s1.record()
        barrier()
        assert len(seen) == steps and all(v == v for v in seen), "e2e: every step's loss must have been read"
        t2 = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": B / (t2.item() / steps / 1e3), "unit": "pairs/s",
                      "loss_readback": "every step's loss is copied to pinned host memory and read by the host inside the timed region, one step late (event wait, no stream synchronize)",
                      "h2d_bytes_per_step": world * (hx_img.numel() * 2 + hx_txt.numel() * 2 + hlab.numel() * 4),
                      "d2h_bytes_per_step": world * 4}

    # ---- parity leg: the data-parallel step against the single-GPU step on the SAME global batch, dropout off -------------
    # (the global batch is identical for every N, so SCALE runs double as a check of the collectives)
    head.eval()
    xi, xt = x_img.detach().clone().requires_grad_(True), x_txt.detach().clone().requires_grad_(True)
    for p in head.parameters():
        p.grad = None
    loss_n = head(xi, xt, class_text, labels)
    loss_n.backward()
    torch.cuda.synchronize()
    par = {"loss": float(loss_n.item()), "dropout": "off for this leg"}
    if world > 1:
        gw_n = head.image_projector.fc.weight.grad.detach().clone()
        if rank == 0:
            fx_img, fx_txt, flab = synth_rows(cfg, 0, B)
            fxi, fxt = fx_img.to(dev).requires_grad_(True), fx_txt.to(dev).requires_grad_(True)
            for p in head.parameters():
                p.grad = None
            from b200clip import dp
            head.group = dp.SOLO                         # world 1 on this rank: the single-GPU step on the full global batch
            loss_1 = head(fxi, fxt, class_text, flab.to(dev))
            loss_1.backward()
            head.group = None
            torch.cuda.synchronize()
            rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
            par.update({"loss_1gpu": float(loss_1.item()),
                        "loss_rel": abs(float(loss_n.item()) - float(loss_1.item())) / abs(float(loss_1.item())),
                        "dx_img_rel_l2": rel(xi.grad.float(), fxi.grad[:b_loc].float()),
                        "dx_txt_rel_l2": rel(xt.grad.float(), fxt.grad[:b_loc].float()),
                        "dw_rel_l2": rel(gw_n, head.image_projector.fc.weight.grad)})
            # the input gradients are bf16 tensors (bf16 inputs): fp32-level differences of the collectives' summation order
            # (measured: d_ihat 2e-7, reduce-scattered d_that 3e-5, tools/dbg_dp_parity.py) flip a fraction of the bf16
            # roundings along the projection backward, hence 5e-3 there and 1e-3 on the fp32 weight gradient
            par["ok"] = bool(par["loss_rel"] <= 1e-5 and par["dx_img_rel_l2"] <= 5e-3 and par["dx_txt_rel_l2"] <= 5e-3 and
                             par["dw_rel_l2"] <= 1e-3)
        dist.barrier()
    out["parity"] = par
    if not eager:
        gstep.close()
    del keep_alive                                    # a live graph holding NCCL kernels would block communicator teardown
    return out


def run_head(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b200clip arm has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's version banner must not land on stdout (one JSON line)
        from b200clip import dp
        dp.init_process_group(dev)                                   # NCCL with high-priority streams (dp.py)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    m = measure_head(args, cfg, rank, world, dev, args.steps, args.warmup, eager=args.eager)
    B, b_loc, ms_step = m["B"], m["b_loc"], m["ms_step"]
    pk = peaks()
    # dominant kernel: the InfoNCE backward.  Algorithmic flops per launch on one rank = 4 * b_loc * B * D (dI and dT products;
    # the recomputed logits are NOT counted), DESIGN.md "Kernels".
    achieved = (4.0 * b_loc * B * cfg["D"]) / (m["bwd_ms"] / 1e3) / 1e12 if m["bwd_ms"] > 0 else 0.0
    flops = head_flops(B, cfg["D"], cfg["E_img"], cfg["E_txt"], cfg["C"]) / world
    # DRAM bytes of one launch from the ncu --set full capture of this kernel at this size (profiles/); no capture exists for
    # other sizes / rank counts
    traffic = NCE_BWD_TRAFFIC_BYTES if (world == 1 and B == 32768 and cfg["D"] == 512) else None
    roofline = {"bound": "tensor", "kernel": NCE_BWD_KERNEL, "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "peak_source": pk["src"] + ", sustained bf16",
                "executed_tflops": 2.0 * achieved,      # the kernel also recomputes the logits once per direction (8 B^2 D executed)
                "launch_ms": m["bwd_ms"], "kernel_timing": m["kernel_timing"], "share_of_step": m["bwd_ms"] / ms_step,
                "fwd_kernel_ms": m["fwd_ms"], "step_algorithmic_tflops": flops / (ms_step / 1e3) / 1e12,
                "step_frac_of_peak": flops / (ms_step / 1e3) / 1e12 / pk["tflops"]}
    line = {
        "metric": "clip_head_fwd_bwd_pairs_per_sec", "value": m["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": head_config(cfg, world, args.eager),
        "e2e": m["e2e"], "gpu_launches": m["launches"], "clocks": m["clocks"], "roofline": roofline, "loss": m["loss"],
        "parity": m["parity"],
    }
    if rank == 0:
        if world == 1 and not args.no_extras and args.config == "cfg3":
            # the other single-GPU BASELINE.json configs, measured by the same code in the same run (driver-reproducible)
            extra = {}
            c2 = CFG["cfg2"]
            m2 = measure_head(args, c2, 0, 1, dev, args.steps, args.warmup, want_e2e=False, kernel_events="none")
            f2 = head_flops(c2["B"], c2["D"], c2["E_img"], c2["E_txt"], c2["C"])
            extra["cfg2"] = {"workload": c2["name"], "value": m2["value"], "unit": "pairs/s", "ms_per_step": m2["ms_step"],
                             "gpu_launches": m2["launches"], "loss": m2["loss"],
                             "step_algorithmic_tflops": f2 / (m2["ms_step"] / 1e3) / 1e12,
                             "step_frac_of_peak": f2 / (m2["ms_step"] / 1e3) / 1e12 / pk["tflops"]}
            z = zeroshot_record(args)
            extra["cfg4_zeroshot"] = {k: z[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches", "roofline", "config")}
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            Bs, _ = pick_cpu_sample(cfg, 2, budget_s=45.0)      # one warm-up-free pair of steps inside ~45 s
            times = cpu_head_step_time(cfg, Bs, reps=2 if Bs < cfg["B"] else 1, warm=0)
            best = min(times)
            line["cpu_baseline"] = {
                "value": Bs / best, "unit": "pairs/s", "cores": cores, "kind": "port", "sample_batch": Bs,
                "extrapolated_full_config": (Bs / best) * Bs / B,
                "sample": (f"oracle port of the reference head (torch fp32 CPU, {cores} threads), best step at B={Bs} of the "
                           f"B={B} global batch" + ("" if Bs == B else "; CPU time grows ~B^2: extrapolated_full_config = "
                                                    "pairs/s the CPU would deliver at the full batch"))}
        emit(line)
    if world > 1:
        # never let teardown hang the job: the result line is already out
        guard = threading.Timer(20.0, lambda: os._exit(0))
        guard.daemon = True
        guard.start()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def zeroshot_record(args):
    """cfg 4: 1M embeddings x 28 (pos,neg) prompts, D=512 -- HBM roofline."""
    from b200clip import _lib, ops
    dev = torch.device("cuda", torch.cuda.current_device())
    N, NP, D = 1_000_000, 28, 512
    g = torch.Generator().manual_seed(1234)
    X = torch.randn(N, D, generator=g).to(torch.bfloat16)
    P = torch.nn.functional.normalize(torch.randn(NP, D, generator=g), dim=1).to(torch.bfloat16)
    Xd, Pd = X.to(dev), P.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lib = _lib.load()

    def run():
        return ops.zeroshot_score(Xd, Pd, pair_mode=True, temperature=0.07, thresholds=[0.5])

    for _ in range(max(args.warmup, 3)):
        run()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    sampler.start()
    n0 = lib.b200clip_launch_count()
    tot = 0.0
    for _ in range(args.steps):
        flush.zero_()                                     # evict X from L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = (lib.b200clip_launch_count() - n0) // args.steps
    ms = tot / args.steps
    bytes_alg = N * D * 2 + NP * D * 2 + N * 3
    pk = peaks()
    # e2e: pinned host embeddings in, label sets out, as a chunked two-stream pipeline: while chunk i is scored, chunk i+1
    # travels host->device on a copy stream and chunk i-1's results travel back on a third; every byte of X crosses PCIe
    # inside the timed region
    NCH = 8
    rows = N // NCH
    hX = X.pin_memory()
    h_am = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_mask = torch.empty(N, dtype=torch.int16).pin_memory()
    dbuf = [torch.empty((rows, D), dtype=torch.bfloat16, device=dev) for _ in range(2)]
    cs, ds = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream()

    def e2e_pass():
        h2d = [torch.cuda.Event() for _ in range(NCH)]
        done = [torch.cuda.Event() for _ in range(NCH)]
        outs = []
        for c in range(NCH):
            with torch.cuda.stream(cs):
                if c >= 2:
                    cs.wait_event(done[c - 2])               # the kernel that read this buffer two chunks ago has finished
                dbuf[c & 1].copy_(hX[c * rows:(c + 1) * rows], non_blocking=True)
                h2d[c].record(cs)
            cur.wait_event(h2d[c])
            o = ops.zeroshot_score(dbuf[c & 1], Pd, pair_mode=True, temperature=0.07, thresholds=[0.5])
            done[c].record(cur)
            outs.append(o)
            with torch.cuda.stream(ds):
                ds.wait_event(done[c])
                h_am[c * rows:(c + 1) * rows].copy_(o["argmax"], non_blocking=True)
                h_mask[c * rows:(c + 1) * rows].copy_(o["mask"], non_blocking=True)
        ds.synchronize()
        cur.synchronize()
        return outs

    e2e_pass()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e_pass()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / 3
    achieved = bytes_alg / (ms / 1e3) / 1e9
    return {"metric": "zeroshot_embeddings_per_sec", "value": N / (ms / 1e3), "unit": "embeddings/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg4: zero-shot 1M embeddings x 28 (pos,neg) prompts, D=512", "l2": "256 MB flush between iterations"},
            "e2e": {"value": N / (e2e_ms / 1e3), "unit": "embeddings/s", "h2d_bytes_per_step": N * D * 2, "d2h_bytes_per_step": N * 3,
                    "pipeline": f"{NCH} chunks, H2D / kernel / D2H on three streams, wall clock over 3 passes (host-synchronised)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "zeroshot_kernel", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": achieved / pk["hbm"], "traffic": ZS_TRAFFIC_BYTES, "peak_source": pk["src"]}}


def run_zeroshot(args):
    torch.cuda.set_device(0)
    emit(zeroshot_record(args))


_RESULT_OUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's real stdout; fd 1 itself is pointed at stderr for the rest of the run so that
    banners printed by libraries (NCCL prints its version to stdout) cannot end up next to it."""
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200clip", choices=["b200clip", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CFG))
    ap.add_argument("--workload", default="head", choices=["head", "zeroshot"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg2 / cfg4 sub-records of the default single-GPU line")
    ap.add_argument("--eager", action="store_true", help="launch the step kernel by kernel instead of replaying one CUDA graph")
    args = ap.parse_args()
    cfg = CFG[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.workload == "zeroshot":
        run_zeroshot(args)
    else:
        run_head(args, cfg)


if __name__ == "__main__":
    main()

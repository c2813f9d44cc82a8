"""Why bench's cfg2 sub-record differs from tools/sweep.py at the same shape: measure_head as is / without the clock sampler /
without the external event-record nodes.   python tools/cfg2_probe.py"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
sys.path.insert(0, ROOT)
import torch
import bench
from b200clip import ops

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
args = types.SimpleNamespace(steps=20, warmup=5)
cfg = bench.CFG["cfg2"]


class NoSampler:
    def __init__(self, *a, **k): pass
    def start(self): pass
    def stop(self): return {"sm_mhz": 0, "sm_max_mhz": 0, "reasons": [], "samples": 0}


def run(tag):
    m = bench.measure_head(args, cfg, 0, 1, dev, 20, 5, want_e2e=False)
    print(tag, round(m["ms_per_step"], 4) if "ms_per_step" in m else {k: m[k] for k in list(m)[:4]})


run("as is          ")
real = bench.ClockSampler
bench.ClockSampler = NoSampler
run("no clock sampler")
bench.ClockSampler = real
orig = ops._timing_events


class _NullEvent:
    def record(self, *a, **k): pass
    def elapsed_time(self, other): return 0.0
    def synchronize(self): pass


ops._timing_events = lambda: (_NullEvent(), _NullEvent())
run("no event nodes  ")

"""torchrun --nproc-per-node N tools/dp_step_time.py [global_B] : device-timed graphed head step (dropout on, like bench.py),
for A/B runs under environment switches (B200CLIP_BWD_SPLITS, B200CLIP_TWO_STREAM_ROWS, B200CLIP_NCCL_PRIO)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import b200clip, bench
from b200clip import dp, head as H
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
if world > 1:
    if os.environ.get("B200CLIP_NCCL_PRIO", "1") == "1": dp.init_process_group(dev)
    else: dist.init_process_group("nccl", device_id=dev)
if "B200CLIP_TWO_STREAM_ROWS" in os.environ: H.TWO_STREAM_MAX_ROWS = int(os.environ["B200CLIP_TWO_STREAM_ROWS"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cfg = dict(bench.CFG["cfg3"], B=B)
torch.manual_seed(0)
head = b200clip.ClipHead(cfg["E_img"], cfg["E_txt"], cfg["D"], cfg["C"], 0.07, 1.0, dropout_rate=0.1).to(dev).train()
x_img, x_txt, labels, class_text = bench.synth_inputs(cfg, B // world, rank, dev)
g = b200clip.GraphedHeadStep(head, x_img, x_txt, class_text, labels)
for _ in range(10): g()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40): loss = g()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 40], device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"N={world} splits={os.environ.get('B200CLIP_BWD_SPLITS','auto')} two_stream_rows={H.TWO_STREAM_MAX_ROWS} "
          f"nccl_prio={os.environ.get('B200CLIP_NCCL_PRIO','1')}: {t.item()*1e3:.1f} us/step  loss {float(loss):.5f}", flush=True)
g.close()
if world > 1:
    import threading; threading.Timer(15.0, lambda: os._exit(0)).start()
    dist.destroy_process_group()

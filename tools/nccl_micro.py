"""torchrun --nproc-per-node N tools/nccl_micro.py : device-timed NCCL collectives at the head step's message sizes."""
import os, sys, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
hp = os.environ.get("HP", "0") == "1"
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=hp)
dist.init_process_group("nccl", device_id=dev, pg_options=opts)
B, D = 32768, 512
n = B // world
def t(fn, it=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
loc = torch.randn(n, D, device=dev).bfloat16(); full = torch.empty(B, D, device=dev, dtype=torch.bfloat16)
part32 = torch.randn(B, D, device=dev); out32 = torch.empty(n, D, device=dev)
part16 = part32.bfloat16(); out16 = torch.empty(n, D, device=dev, dtype=torch.bfloat16)
c = torch.randn(B, device=dev); g4 = torch.randn(1_050_000, device=dev); g8 = torch.randn(2_000_000, device=dev); s6 = torch.randn(6, device=dev, dtype=torch.float64)
res = {
 "all_gather bf16 [B,512] (33.5 MB total)": t(lambda: dist.all_gather_into_tensor(full, loc)),
 "reduce_scatter f32 [B,512] (67 MB in)": t(lambda: dist.reduce_scatter_tensor(out32, part32)),
 "reduce_scatter bf16 [B,512] (33.5 MB in)": t(lambda: dist.reduce_scatter_tensor(out16, part16)),
 "all_reduce f32 [B] (128 KB)": t(lambda: dist.all_reduce(c)),
 "all_reduce f32 4.2 MB": t(lambda: dist.all_reduce(g4)),
 "all_reduce f32 8 MB": t(lambda: dist.all_reduce(g8)),
 "all_reduce f64 [6]": t(lambda: dist.all_reduce(s6)),
}
if rank == 0:
    print(f"# world {world} NCCL_PROTO={os.environ.get('NCCL_PROTO')} NCCL_ALGO={os.environ.get('NCCL_ALGO')} high_priority={hp}")
    for k, v in res.items(): print(f"{v:8.1f} us  {k}")
dist.destroy_process_group()

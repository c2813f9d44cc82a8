"""torchrun --nproc-per-node 2 tools/dbg_dp_parity.py : stage-by-stage difference between the data-parallel head step and the
single-GPU step on the same global batch (rank 0 prints)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import b200clip, bench
from b200clip import dp, ops, head as H
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
__import__("b200clip").dp.init_process_group(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cfg = dict(bench.CFG["cfg3"], B=B)
b_loc = B // world
torch.manual_seed(0)
head = b200clip.ClipHead(cfg["E_img"], cfg["E_txt"], cfg["D"], cfg["C"], 0.07, 1.0).to(dev).eval()
x_img, x_txt, labels, class_text = bench.synth_inputs(cfg, b_loc, rank, dev)
one = torch.ones((), device=dev)
def run(xi, xt, lab, group):
    _, finish, (tensors, meta) = H.head_forward(xi, xt, class_text, lab, 0.07, 1.0, group, 0.0, 0, head.params(), True, defer_loss=True)
    (xi_, xt_, iw1b, iw2b, tw1b, tw2b, ig, tg, y_img, y_txt, ihat, that_all, inv_img, inv_txt, rinvh, cinvh, d_bce, coef, db_raw, *rest) = tensors
    d_ihat, d_that = ops.infonce_backward(ihat, that_all, 0.07, rinvh, cinvh, one, row0=meta["row0"], allow_splits=True)
    dxi, dxt, grads = H.head_backward(tensors, meta, one)
    loss, parts, _ = finish()
    return dict(ihat=ihat.float(), that_all=that_all.float(), rinvh=rinvh, cinvh=cinvh, d_ihat=d_ihat if d_ihat.dim() == 2 else d_ihat.sum(0),
                d_that=d_that, dxi=dxi.float(), dxt=dxt.float(), loss=loss, d_bce=d_bce)
a = run(x_img, x_txt, labels, None)
d_that_sum = a["d_that"].clone(); dist.all_reduce(d_that_sum)
torch.cuda.synchronize()
if rank == 0:
    fi, ft, fl = bench.synth_rows(cfg, 0, B)
    b = run(fi.to(dev), ft.to(dev), fl.to(dev), dp.SOLO)
    rel = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm().clamp_min(1e-30))
    n = b_loc
    print("loss", float(a["loss"]), float(b["loss"]))
    print("ihat", rel(a["ihat"], b["ihat"][:n]), "that_all", rel(a["that_all"], b["that_all"]))
    print("rinvh", rel(a["rinvh"], b["rinvh"][:n]), "cinvh", rel(a["cinvh"], b["cinvh"]))
    print("d_bce", rel(a["d_bce"], b["d_bce"][:n]))
    print("d_ihat", rel(a["d_ihat"], b["d_ihat"][:n]), "d_that(sum over ranks)", rel(d_that_sum, b["d_that"]))
    print("dxi", rel(a["dxi"], b["dxi"][:n]), "dxt", rel(a["dxt"], b["dxt"][:n]))
dist.barrier()
dist.destroy_process_group()

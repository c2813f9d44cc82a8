"""world_size-2 (and 4) gloo test of the data-parallel choreography in b200clip/dp.py on CPU tensors: the oracle supplies the
per-rank arithmetic the CUDA kernels perform on a B200, dp.py supplies the collectives + combination algebra that head.py
uses over NCCL.  The result must equal the monolithic oracle (reference contrastive_loss, 0426/train.py:154-176)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, tau, out_dir):
    for p in (os.path.join(ROOT, "clip-for-dl_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import ref_head as R
    import synth
    from b200clip import dp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    n = B // world
    I = synth.unit_rows(11, B, 64).double()
    T = synth.unit_rows(12, B, 64).double()
    I_loc, T_loc = I[rank * n:(rank + 1) * n], T[rank * n:(rank + 1) * n]
    assert dp.world() == world and dp.rank() == rank
    # forward: gather text rows, local row block vs all columns, combine column sums and loss scalars
    T_all, work = dp.gather_rows(T_loc.contiguous(), async_op=True)
    work.wait()
    assert torch.equal(T_all, T)
    r, c_part, diag, m = R.contrastive_loss_flash(I_loc, T_all, tau, row0=rank * n)
    # global label count: asynchronous, waited on only before the BCE heads (head.py)
    lsum = torch.tensor(float(rank + 1))
    lsum_work = dp.sum_across_async(lsum)
    c = dp.sum_across(c_part.clone())                       # the one collective on the forward critical path
    # six loss numerators (3 InfoNCE | 3 BCE) in ONE all-reduce that nothing in the backward pass waits for
    sums6 = torch.zeros(6, dtype=torch.float64)
    sums6[:3] = torch.stack([torch.log(r).sum(), torch.log(c[rank * n:(rank + 1) * n]).sum(), diag])
    sums6[3:] = torch.tensor([1.0, 2.0, 3.0]) * (rank + 1)
    sums_work = dp.sum_across_async(sums6)
    # backward: dI local, dT partial -> reduce-scatter; two gradient buckets, the first one asynchronous
    dI, dT_part = R.contrastive_grads_flash(I_loc, T_all, tau, r, c, row0=rank * n)
    dT_loc, _ = dp.scatter_sum_rows(dT_part.clone())
    g_img, g_work = dp.allreduce_flat([torch.full((3,), float(rank + 1)), torch.full((2, 2), 10.0 * (rank + 1))], async_op=True)
    g_txt = dp.allreduce_flat([torch.full((5,), 100.0 * (rank + 1))])
    dp.wait(g_work)
    g = [g_img[0], g_img[1]]
    dp.wait(sums_work)
    dp.wait(lsum_work)
    loss = dp.infonce_loss_from_sums(sums6[:3], tau, B)
    tri = world * (world + 1) / 2
    assert torch.allclose(sums6[3:], torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64) * tri)
    assert torch.allclose(g_txt[0], torch.full((5,), 100.0 * tri))
    torch.save({"loss": loss, "dI": dI, "dT": dT_loc, "g0": g[0], "g1": g[1], "lsum": lsum}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_dp_head_matches_monolithic(tmp_path, world):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_head as R
    import synth
    B, tau = 48, 0.07
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, B, tau, str(tmp_path)), nprocs=world, join=True)
    I = synth.unit_rows(11, B, 64).double().requires_grad_(True)
    T = synth.unit_rows(12, B, 64).double().requires_grad_(True)
    ref = R.contrastive_loss(I, T, tau)
    ref.backward()
    outs = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world)]
    n = B // world
    tri = world * (world + 1) / 2
    for r, o in enumerate(outs):
        assert abs(o["loss"].item() - ref.item()) < 1e-5 * abs(ref.item())
        assert torch.allclose(o["dI"], I.grad[r * n:(r + 1) * n], rtol=1e-9, atol=1e-12)
        assert torch.allclose(o["dT"], T.grad[r * n:(r + 1) * n], rtol=1e-9, atol=1e-12)
        assert torch.equal(o["g0"], torch.full((3,), tri)) and torch.equal(o["g1"], torch.full((2, 2), 10.0 * tri))
        assert o["lsum"].item() == tri


def test_single_process_is_identity():
    sys.path.insert(0, os.path.join(ROOT, "clip-for-dl_b200"))
    from b200clip import dp
    x = torch.arange(6.0).reshape(3, 2)
    full, work = dp.gather_rows(x)
    assert full is x and work is None
    out, work = dp.scatter_sum_rows(x)
    assert out is x and work is None
    assert dp.allreduce_flat([x])[0] is x
    assert dp.world() == 1 and dp.rank() == 0

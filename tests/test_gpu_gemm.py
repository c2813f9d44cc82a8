"""tcgen05 GEMM (gemm.cuh) vs fp32 matmul on the same bf16-rounded operands: all operand majors, epilogues, ragged M,
K tails (TMA zero fill) and split-K."""
import pytest
import torch

from gpu_util import dev, gpu, rel_l2
import synth

pytestmark = gpu


def _mk(seed, *shape):
    return synth.randn(seed, *shape).to(dev()).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 128), (256, 512, 512), (200, 512, 768), (32, 768, 2048),
                                   (1000, 2048, 512), (16, 512, 768)])
def test_nt_store_f32(M, N, K):
    from b200clip import ops
    a, b = _mk(1, M, K), _mk(2, N, K)
    out = ops.gemm_bf16(a, b)
    ref = a.float() @ b.float().T
    assert rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 512), (256, 768, 512), (64, 2048, 512)])
def test_nn_b_mn_major(M, N, K):
    from b200clip import ops
    a, b = _mk(3, M, K), _mk(4, K, N)                     # b stored [K][N]
    out = ops.gemm_bf16(a, b, b_mn=True)
    assert rel_l2(out, a.float() @ b.float()) < 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (512, 512, 1000), (512, 768, 4096), (512, 2048, 300)])
def test_tn_both_mn_major_and_split_k(M, N, K):
    from b200clip import ops
    a, b = _mk(5, K, M), _mk(6, K, N)                     # a stored [K][M], b stored [K][N]
    ref = a.float().T @ b.float()
    assert rel_l2(ops.gemm_bf16(a, b, a_mn=True, b_mn=True), ref) < 1e-5
    out = ops.gemm_bf16(a, b, a_mn=True, b_mn=True, epilogue=ops.EPI_ATOMIC_F32, split_k=5)
    assert rel_l2(out, ref) < 1e-5


def test_a_mn_b_k():
    from b200clip import ops
    a, b = _mk(7, 320, 256), _mk(8, 512, 320)             # a[K][M], b[N][K]
    assert rel_l2(ops.gemm_bf16(a, b, a_mn=True), a.float().T @ b.float().T) < 1e-5


def test_epilogues():
    from b200clip import ops
    M, N, K = 300, 512, 768
    a, b = _mk(9, M, K), _mk(10, N, K) * 0.05
    bias = synth.randn(11, N).to(dev())
    acc = a.float() @ b.float().T
    # +bias, GELU
    p, h = ops.gemm_bf16(a, b, epilogue=ops.EPI_BIAS_GELU, bias=bias)
    pref = (acc + bias).to(torch.bfloat16)
    assert rel_l2(p.float(), pref.float()) < 2e-3
    assert rel_l2(h.float(), torch.nn.functional.gelu(p.float())) < 4e-3
    # +bias +resid -> f32
    z = ops.gemm_bf16(a, b, epilogue=ops.EPI_BIAS_RESID_F32, bias=bias, resid=p)
    assert rel_l2(z, acc + bias + p.float()) < 1e-5
    # bf16 store and relu
    assert rel_l2(ops.gemm_bf16(a, b, epilogue=ops.EPI_STORE_BF16, bias=bias).float(), acc + bias) < 4e-3
    assert rel_l2(ops.gemm_bf16(a, b, epilogue=ops.EPI_RELU_BF16, bias=bias).float(), torch.relu(acc + bias)) < 4e-3
    # gelu backward epilogue: acc * gelu'(p) + aux
    w = _mk(12, N, N) * 0.05                              # [K=N][N] read MN-major
    dz = _mk(13, M, N)
    aux = synth.randn(14, M, N).to(dev())
    out = ops.gemm_bf16(dz, w, b_mn=True, epilogue=ops.EPI_GELU_BWD, resid=p, aux=aux)
    pf = p.float().requires_grad_(True)
    torch.nn.functional.gelu(pf).backward(dz.float() @ w.float())
    assert rel_l2(out.float(), pf.grad + aux) < 4e-3


def test_rejects_bad_arguments():
    from b200clip import ops
    a, b = _mk(1, 128, 64), _mk(2, 40, 64)
    with pytest.raises(RuntimeError):
        ops.gemm_bf16(a, b)                                # N not a multiple of 32

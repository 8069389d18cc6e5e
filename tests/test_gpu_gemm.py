"""GPU parity: the tcgen05 GEMM family (through the C ABI) vs a plain fp32 torch reference and vs the
CUDA-core restatement (clipn_gemm_ref). Tolerances: bf16 outputs rel-L2 <= 4e-3 (one bf16 rounding of an
fp32-accumulated result), fp32 outputs rel-L2 <= 2e-5 (summation-order noise)."""
import math

import pytest
import torch

from open_clip_b200 import _lib as L
from open_clip_b200 import ops
from gpu_util import BF16, F32, max_err, randn, rel_err

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (256, 256, 128), (400, 384, 192), (136, 128, 520), (1024, 768, 3072), (2048, 2304, 768)]


def _operands(M, N, K, a_mn, b_mn, seed):
    a = randn(M, K, seed=seed, scale=0.5)
    b = randn(N, K, seed=seed + 1, scale=0.5)
    ref = a.float() @ b.float().T
    A = a.T.contiguous() if a_mn else a
    Bm = b.T.contiguous() if b_mn else b
    return A, Bm, ref


@pytest.mark.parametrize("ref_path", [True, False], ids=["cuda_core_ref", "tcgen05"])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("shape", SHAPES[:4] + [SHAPES[4]])
def test_store_layouts(shape, a_mn, b_mn, ref_path):
    M, N, K = shape
    if ref_path and M * N * K > 64 * 1024 * 1024:
        pytest.skip("restatement kernel is for small cases")
    A, Bm, ref = _operands(M, N, K, a_mn, b_mn, 1)
    out = ops.gemm(A, Bm, a_mn=a_mn, b_mn=b_mn, ref=ref_path)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 4e-3, (shape, a_mn, b_mn, rel_err(out, ref))


@pytest.mark.parametrize("shape", SHAPES)
def test_bias_and_f32_store(shape):
    M, N, K = shape
    A, Bm, ref = _operands(M, N, K, False, False, 2)
    bias = randn(N, seed=5)
    out = ops.gemm(A, Bm, bias=bias, alpha=0.5)
    out32 = ops.gemm(A, Bm, bias=bias, alpha=0.5, epilogue=L.EPI_STORE_F32)
    want = 0.5 * ref + bias.float()
    assert rel_err(out, want) < 4e-3
    assert rel_err(out32, want) < 2e-5


@pytest.mark.parametrize("shape", SHAPES[1:5])
def test_bias_gelu_and_resid(shape):
    M, N, K = shape
    A, Bm, ref = _operands(M, N, K, False, False, 3)
    bias = randn(N, seed=6)
    h = torch.empty((M, N), dtype=BF16, device="cuda")
    g = torch.empty((M, N), dtype=BF16, device="cuda")
    ops.gemm(A, Bm, bias=bias, epilogue=L.EPI_BIAS_GELU, out=h, out2=g)
    want_h = (ref + bias.float()).to(BF16)
    assert rel_err(h, want_h) < 4e-3
    assert rel_err(g, torch.nn.functional.gelu(want_h.float())) < 6e-3
    res = randn(M, N, seed=7)
    y = ops.gemm(A, Bm, bias=bias, aux=res, epilogue=L.EPI_BIAS_RESID)
    assert rel_err(y, want_h.float() + res.float()) < 4e-3


@pytest.mark.parametrize("shape", SHAPES[1:5])
def test_dgelu(shape):
    M, N, K = shape
    A, Bm, ref = _operands(M, N, K, False, True, 4)
    hpre = randn(M, N, seed=8)
    dh = torch.empty((M, N), dtype=BF16, device="cuda")
    g = torch.empty((M, N), dtype=BF16, device="cuda")
    csum = torch.zeros(N, dtype=F32, device="cuda")
    ops.gemm(A, Bm, b_mn=True, epilogue=L.EPI_DGELU, aux=hpre, out=dh, out2=g, col_sum=csum)
    x = hpre.float().requires_grad_(True)
    y = torch.nn.functional.gelu(x)
    y.backward(ref)
    assert rel_err(dh, x.grad) < 6e-3
    assert rel_err(g, y.detach()) < 4e-3
    assert rel_err(csum, x.grad.sum(0)) < 2e-3       # fused bias gradient (column sums of dh)
    dh2 = torch.empty_like(dh)                       # second output is optional
    ops.gemm(A, Bm, b_mn=True, epilogue=L.EPI_DGELU, aux=hpre, out=dh2)
    assert torch.equal(dh, dh2)
    csum2 = torch.zeros(N, dtype=F32, device="cuda")
    st = ops.gemm(A, Bm, b_mn=True, col_sum=csum2)   # STORE + fused column sums
    assert rel_err(csum2, ref.sum(0)) < 2e-3 and rel_err(st, ref) < 4e-3


@pytest.mark.parametrize("shape", SHAPES[1:5])
def test_gelu_grad_forward_and_mul_aux_backward(shape):
    """fast-mode MLP pair: forward keeps gelu'(h) and gelu(h); backward multiplies the dgrad by the saved derivative."""
    M, N, K = shape
    A, Bm, ref = _operands(M, N, K, False, False, 5)
    bias = randn(N, seed=6)
    gp = torch.empty((M, N), dtype=BF16, device="cuda")
    g = torch.empty((M, N), dtype=BF16, device="cuda")
    ops.gemm(A, Bm, bias=bias, epilogue=L.EPI_BIAS_GELU_GRAD, out=gp, out2=g)
    h = (ref + bias.float()).to(BF16).float().requires_grad_(True)
    y = torch.nn.functional.gelu(h)
    y.sum().backward()
    assert rel_err(g, y.detach()) < 6e-3
    assert rel_err(gp, h.grad) < 6e-3
    csum = torch.zeros(N, dtype=F32, device="cuda")
    dh = ops.gemm(A, Bm, epilogue=L.EPI_MUL_AUX, aux=gp, col_sum=csum)
    want = ref * gp.float()
    assert rel_err(dh, want) < 4e-3
    assert rel_err(csum, want.sum(0)) < 2e-3


@pytest.mark.parametrize("splits", [1, 3, 7])
@pytest.mark.parametrize("shape", [(256, 128, 4096), (768, 768, 6400), (2304, 768, 3200)])
def test_wgrad_split_k(shape, splits):
    M, N, K = shape
    A, Bm, ref = _operands(M, N, K, True, True, 9)
    out = torch.zeros((M, N), dtype=F32, device="cuda")
    ops.gemm(A, Bm, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=out, splits=splits)
    ops.gemm(A, Bm, a_mn=True, b_mn=True, epilogue=L.EPI_ACCUM_F32, out=out, splits=splits)  # accumulates
    assert rel_err(out, 2 * ref) < 2e-5


def test_tc_matches_cuda_core_restatement_bitwise_close():
    A, Bm, _ = _operands(384, 256, 256, False, False, 11)
    x = ops.gemm(A, Bm, epilogue=L.EPI_STORE_F32)
    y = ops.gemm(A, Bm, epilogue=L.EPI_STORE_F32, ref=True)
    assert max_err(x, y) < 1e-3


def test_bad_arguments_raise():
    a = randn(128, 64)
    with pytest.raises(L.ClipnError):
        ops.gemm(a, randn(100, 64))  # N not a multiple of 32
    with pytest.raises(L.ClipnError):
        ops.gemm(a, randn(128, 32))  # K mismatch

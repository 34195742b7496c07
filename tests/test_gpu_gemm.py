"""GPU: the GEMM kernels through the C ABI (gims_linear) — fp32 CUDA-core kernel and tcgen05 3xTF32 kernel
against an fp64 torch reference; ragged row counts, K-concatenated inputs, bias / residual / ReLU epilogues."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    # rows_max, rows_live, N, K0, K1, bias, residual, relu
    (4096, 4096, 768, 256, 0, True, False, False),
    (4096, 4001, 512, 256, 256, True, False, True),
    (4096, 4096, 256, 512, 0, True, True, False),
    (2048, 1999, 128, 128, 128, True, False, True),
    (300, 257, 256, 256, 0, False, False, False),
    (200, 200, 64, 32, 0, True, False, True),
    (129, 129, 32, 64, 0, True, False, False),
    (1, 1, 256, 128, 0, True, True, False),
]


def _run(mode, rows_max, rows, N, K0, K1, bias, resid, relu, seed=0):
    from gims_b200 import _lib
    L = _lib.lib()
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(seed)
    A0 = torch.randn(rows_max, K0, generator=g).to(dev)
    A1 = torch.randn(rows_max, K1, generator=g).to(dev) if K1 else None
    W = (torch.randn(N, K0 + K1, generator=g) / (K0 + K1) ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    R = torch.randn(rows_max, N, generator=g).to(dev) if resid else None
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.gims_split_tf32(_lib.ptr(W), _lib.ptr(hi), _lib.ptr(lo), W.numel(), st), 'split')
    assert torch.equal(hi + lo, W)
    Y = torch.full((rows_max, N), float('nan'), device=dev)
    nd = torch.tensor([rows], dtype=torch.int32, device=dev)
    _lib.check(L.gims_linear(_lib.ptr(A0), K0, K0, _lib.ptr(A1), K1, K1, _lib.ptr(W), _lib.ptr(hi), _lib.ptr(lo),
                             _lib.ptr(b), _lib.ptr(R), N, _lib.ptr(Y), N, N, int(relu), rows_max, _lib.ptr(nd), mode, st),
               'gims_linear')
    torch.cuda.synchronize()
    A = torch.cat([A0, A1], 1) if K1 else A0
    ref = A.double() @ W.double().t()
    if bias:
        ref = ref + b.double()
    if resid:
        ref = ref + R.double()
    if relu:
        ref = ref.relu()
    assert torch.isnan(Y[rows:]).all(), 'rows beyond the live count were written'
    err = (Y[:rows].double() - ref[:rows]).abs().max().item()
    scale = ref[:rows].abs().max().item()
    return err / scale


@pytest.mark.parametrize('shape', SHAPES)
def test_simt_gemm(shape):
    from gims_b200 import _lib
    e = _run(_lib.GEMM_SIMT, *shape)
    print('\n[simt gemm %s] rel err %.2e' % (shape[:5], e))
    assert e < 2e-6


@pytest.mark.parametrize('shape', SHAPES)
def test_tc_gemm(shape):
    from gims_b200 import _lib
    e = _run(_lib.GEMM_TC, *shape)
    print('\n[tc gemm %s] rel err %.2e' % (shape[:5], e))
    assert e < 4e-6

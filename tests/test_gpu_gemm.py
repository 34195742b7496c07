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
    # the persistent kernel's corners: one k-block per tile (a single issuer / accumulator), K concatenated from two
    # 64-wide buffers, and many tiles per CTA (628 work items on 148 CTAs) with ragged last rows
    (1000, 977, 256, 64, 0, True, True, True),
    (700, 700, 384, 64, 64, True, False, False),
    (20000, 19999, 512, 256, 0, True, True, True),
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
    from gims_b200.packing import split_f16
    h16, l16, sinv = [t.to(dev) for t in split_f16(W.cpu())]
    Y = torch.full((rows_max, N), float('nan'), device=dev)
    nd = torch.tensor([rows], dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.gims_linear(_lib.ptr(A0), K0, K0, _lib.ptr(A1), K1, K1, _lib.ptr(W), _lib.ptr(hi), _lib.ptr(lo),
                             _lib.ptr(h16), _lib.ptr(l16), _lib.ptr(sinv), _lib.ptr(b), _lib.ptr(R), N, _lib.ptr(Y), N, N,
                             int(relu), rows_max, _lib.ptr(nd), mode, _lib.ptr(status), st), 'gims_linear')
    torch.cuda.synchronize()
    assert int(status.cpu()) == 0
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
def test_f16_gemm(shape):
    """fp16 hi + lo operands (k-blocks of 64; shapes with K % 64 != 0 run the tf32 kernel)."""
    from gims_b200 import _lib
    e = _run(_lib.GEMM_TC_F16, *shape)
    print('\n[f16x2 gemm %s] rel err %.2e' % (shape[:5], e))
    assert e < 4e-6


def test_f16_gemm_range_flag():
    """An activation >= 32768 raises GIMS_STATUS_FP16_RANGE instead of silently overflowing."""
    from gims_b200 import _lib
    from gims_b200.packing import split_f16
    L = _lib.lib()
    dev = torch.device('cuda')
    A = torch.randn(256, 128, device=dev)
    A[77, 5] = 5e4
    W = torch.randn(64, 128) / 11
    hi, lo = torch.empty(64, 128, device=dev), torch.empty(64, 128, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    Wd = W.to(dev)
    _lib.check(L.gims_split_tf32(_lib.ptr(Wd), _lib.ptr(hi), _lib.ptr(lo), Wd.numel(), st), 'split')
    h16, l16, sinv = [t.to(dev) for t in split_f16(W)]
    Y = torch.empty(256, 64, device=dev)
    nd = torch.tensor([256], dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.gims_linear(_lib.ptr(A), 128, 128, None, 0, 0, _lib.ptr(Wd), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(h16),
                             _lib.ptr(l16), _lib.ptr(sinv), None, None, 0, _lib.ptr(Y), 64, 64, 0, 256, _lib.ptr(nd),
                             _lib.GEMM_TC_F16, _lib.ptr(status), st), 'gims_linear')
    torch.cuda.synchronize()
    assert int(status.cpu()) & _lib.STATUS_FP16_RANGE


@pytest.mark.parametrize('shape', SHAPES)
def test_tc_gemm(shape):
    from gims_b200 import _lib
    e = _run(_lib.GEMM_TC, *shape)
    print('\n[tc gemm %s] rel err %.2e' % (shape[:5], e))
    assert e < 4e-6


@pytest.mark.parametrize('n0,n1,layer', [(512, 470, 0), (512, 470, 1), (2048, 2048, 3), (130, 64, 1), (1000, 1337, 0)])
def test_attention_layer_tc_vs_simt(n0, n1, layer):
    """One AttentionalPropagation layer (QKV GEMM -> attention -> merge -> MLP, a-11/a-12) through the C ABI:
    tcgen05 path against the fp32 CUDA-core path and the oracle on the same descriptors, ragged live counts."""
    from gims_b200 import GMatcher, _lib
    from gims_b200.synth import make_state_dict
    from oracle import gims_oracle as orc
    L = _lib.lib()
    dev = torch.device('cuda')
    sd = make_state_dict(7)
    gm = GMatcher({})
    gm.load_state_dict(sd)
    gm = gm.cuda().eval()
    model = gm.handle()
    g = torch.Generator().manual_seed(n0 + n1 + layer)
    live0, live1 = n0 - 5, n1 - 3
    desc = torch.randn(n0 + n1, 256, generator=g)
    nd = torch.tensor([live0, live1], dtype=torch.int32, device=dev)
    scratch = torch.zeros(L.gims_attn_scratch_floats(n0 + n1), device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = {}
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    default_mode = L.gims_get_gemm_mode()
    for mode in (_lib.GEMM_SIMT, _lib.GEMM_TC, _lib.GEMM_TC_F16, _lib.GEMM_BF16):
        L.gims_set_gemm_mode(mode)
        d = desc.to(dev).clone()
        _lib.check(L.gims_attn_layer_forward(model, layer, _lib.ptr(d), n0, n1, _lib.ptr(nd), _lib.ptr(scratch),
                                             _lib.ptr(status), st), 'attn layer')
        torch.cuda.synchronize()
        outs[mode] = d.cpu()
    L.gims_set_gemm_mode(default_mode)
    assert int(status.cpu()) == 0
    x0, x1 = desc[:live0].t()[None], desc[n0:n0 + live1].t()[None]
    cross = layer % 2 == 1
    with torch.no_grad():
        d0 = orc.attn_propagation(sd, layer, x0, x1 if cross else x0)
        d1 = orc.attn_propagation(sd, layer, x1, x0 if cross else x1)
    want = torch.cat([(x0 + d0)[0].t(), (x1 + d1)[0].t()])
    # fp32-parity modes: CUDA-core fp32, 3xTF32, fp16 hi+lo attention operands; the bf16 variant is only reported
    for mode, name, bar in ((_lib.GEMM_SIMT, 'simt', 5e-6), (_lib.GEMM_TC, 'tf32x3', 5e-6), (_lib.GEMM_TC_F16, 'f16x2', 5e-6),
                            (_lib.GEMM_BF16, 'bf16', 2e-2)):
        got = torch.cat([outs[mode][:live0], outs[mode][n0:n0 + live1]])
        err = ((got - want).abs().max() / want.abs().max()).item()
        print('\n[attn layer %s n=(%d,%d) layer %d] rel err %.2e' % (name, n0, n1, layer, err))
        assert err < bar, name

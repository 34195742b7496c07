"""Warm launch time and the pipeline timeline of CTA (0,0) of the tensor-core GEMM (gims_debug_gemm_trace)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gims_b200 import _lib

L = _lib.lib()
dev = torch.device('cuda')
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
MODE = _lib.GEMM_MODES[sys.argv[2]] if len(sys.argv) > 2 else _lib.GEMM_TC_F16
from gims_b200.packing import split_f16


def run(K0, K1, N, reps=50, dump=True):
    A0 = torch.randn(rows, K0, device=dev)
    A1 = torch.randn(rows, K1, device=dev) if K1 else None
    W = torch.randn(N, K0 + K1, device=dev) / (K0 + K1) ** 0.5
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    _lib.check(L.gims_split_tf32(_lib.ptr(W), _lib.ptr(hi), _lib.ptr(lo), W.numel(), st), 'split')
    h16, l16, sinv = [t.to(dev) for t in split_f16(W.cpu())]
    b = torch.randn(N, device=dev)
    Y = torch.empty(rows, N, device=dev)
    nd = torch.tensor([rows], dtype=torch.int32, device=dev)

    def call():
        _lib.check(L.gims_linear(_lib.ptr(A0), K0, K0, _lib.ptr(A1), K1, K1, _lib.ptr(W), _lib.ptr(hi), _lib.ptr(lo),
                                 _lib.ptr(h16), _lib.ptr(l16), _lib.ptr(sinv), _lib.ptr(b), None, N, _lib.ptr(Y), N, N, 1, rows,
                                 _lib.ptr(nd), MODE, None, st), 'lin')
    for _ in range(5):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / reps
    trace = torch.zeros(64, dtype=torch.int64, device=dev)
    L.gims_debug_gemm_trace(C.c_void_p(trace.data_ptr()))
    call()
    torch.cuda.synchronize()
    L.gims_debug_gemm_trace(None)
    t = [int(x) for x in trace.cpu()]
    flops = 2.0 * rows * (K0 + K1) * N
    print('K=%d+%d N=%d: %.2f us/launch back-to-back (%.1f TFLOP/s useful fp32-equivalent)' % (K0, K1, N, us, flops / us / 1e6))
    if not dump:
        return
    z = t[0]
    if os.environ.get('GIMS_GEMM_PERSIST', '1') != '0' and MODE == _lib.GEMM_TC_F16 and N >= 256:
        # persistent kernel: 8 stamps per tile of CTA 0
        print('  CTA 0 exit %d clk.  tile: first operands | last A block split | MMAs retired | PARK free | fold done | first MMA || epilogue: PARK full  done' % (t[1] - z))
        for lt in range(4):
            r = t[8 + 8 * lt: 16 + 8 * lt]
            if r[0] == 0:
                break
            print('  %d %14d %14d %14d %10d %10d %10d || %10d %10d' % (lt, r[0] - z, r[1] - z, r[2] - z, r[3] - z, r[4] - z, r[7] - z, r[5] - z, r[6] - z))
        return
    nkb = (K0 + K1) // 32
    print('  entry 0 | prologue done %d | accum_full seen %d | epilogue done %d | exit %d | globaltimer span %d ns' %
          (t[1] - z, t[3] - z, t[4] - z, t[5] - z, t[7] - t[6]))
    print('  epilogue chunk 0: acc loaded %d, staged %d, stored %d | chunk 1: %d %d %d' % tuple(x - z for x in t[56:62]))
    print('  kb: full_seen  tmem_slot_free  split_done | mma_start  mma_issued  committed   (first 8 k-blocks)')
    for kb in range(min(nkb, 8)):
        print('  %2d %8d %8d %8d | %8d %8d %8d' % (kb, t[8 + kb] - z, t[48 + kb] - z, t[16 + kb] - z, t[24 + kb] - z,
                                               t[32 + kb] - z, t[40 + kb] - z))


for shape in [(256, 0, 256), (256, 256, 512), (512, 0, 256), (256, 0, 768)]:
    run(*shape)

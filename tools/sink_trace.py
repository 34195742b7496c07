"""Timeline of CTA 0 of the Sinkhorn kernel (gims_debug_sinkhorn_trace): `python tools/sink_trace.py [n] [offset]`
(n keypoints per image, default 2048; offset = mean score, 90 mimics trained weights)."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gims_b200 import _lib
L = _lib.lib()
dev = torch.device('cuda')
n0 = n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
offset = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
ld = L.gims_couplings_ld(n1)
coup = torch.randn(n0 + 1, ld, device=dev) * 2.6 + offset
coup[n0, :] = 1.0
coup[:, n1] = 1.0
nd = torch.tensor([n0, n1], dtype=torch.int32, device=dev)
ws = torch.empty(L.gims_sinkhorn_workspace_bytes(n0, n1), dtype=torch.uint8, device=dev)
uo, vo = torch.zeros(n0 + 1, device=dev), torch.zeros(n1 + 1, device=dev)
i0, i1 = torch.zeros(n0, dtype=torch.int32, device=dev), torch.zeros(n1, dtype=torch.int32, device=dev)
m0, m1 = torch.zeros(n0, dtype=torch.int64, device=dev), torch.zeros(n1, dtype=torch.int64, device=dev)
s0, s1 = torch.zeros(n0, device=dev), torch.zeros(n1, device=dev)
trace = torch.zeros(16 * 8 + 256 * 8, dtype=torch.int64, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    _lib.check(L.gims_sinkhorn_match(_lib.ptr(coup), ld, n0, n1, _lib.ptr(nd), 100, 0.2, _lib.ptr(ws), ws.numel(), _lib.ptr(uo), _lib.ptr(vo),
                                     _lib.ptr(i0), _lib.ptr(i1), _lib.ptr(m0), _lib.ptr(m1), _lib.ptr(s0), _lib.ptr(s1), None, st), 'sinkhorn')
run(); run(); torch.cuda.synchronize()
L.gims_debug_sinkhorn_trace(C.c_void_p(trace.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
L.gims_debug_sinkhorn_trace(None)
print('kernel+finalize time %.1f us' % (1000 * e0.elapsed_time(e1)))
tall = trace.cpu()
t = tall[:128].view(16, 8)
pc = tall[128:128 + 148 * 8].view(148, 8)
if n1 > 2051 or n0 > 2367:
    print('streamed kernel, CTA 0 (cycles): iter  rows(stream E)  red.add  grid hop  gather  barrier | total')
    for it in range(1, 8):
        r = [int(x) for x in t[it]]
        print('%3d %14d %9d %9d %8d %8d | %7d' % (it, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[5] - r[0]))
    sys.exit(0)
print('scaled-kernel path, CTA 0 (cycles):')
print('iter   row: fma+butterfly+sync  finish rows+sync | col: fma   red+leftover | grid hop | gather: loads+math  barrier | total')
for it in range(1, 12):
    r = [int(x) for x in t[it]]
    print('%3d   %14d %18d %14d %12d %12d %14d %10d %10d' % (it, r[7] - r[0], r[1] - r[7], r[6] - r[1], r[2] - r[6], r[3] - r[2],
                                                       r[4] - r[3], r[5] - r[4], r[5] - r[0]))

import numpy as np
pc = pc.numpy().astype(np.int64)
pc = pc[pc[:, 0] > 0]          # CTAs of this launch (the on-chip kernel uses half the SMs)
print('CTAs traced: %d' % len(pc))
base = pc[:, 0].min()
print('per-CTA globaltimer (ns) at iteration 5, relative to the earliest start:')
print('cta   start  row_done  col_done  hop_done  gather_done')
order = np.argsort(pc[:, 2])
for b in list(order[:4]) + list(order[-8:]):
    r = pc[b] - base
    print('%3d  %6d %8d %9d %9d %10d' % (b, r[0], r[1], r[2], r[3], r[5]))
print('col_done spread: min %d max %d median %d' % ((pc[:,2]-base).min(), (pc[:,2]-base).max(), int(np.median(pc[:,2]-base))))
print('start spread: min %d max %d' % ((pc[:,0]-base).min(), (pc[:,0]-base).max()))

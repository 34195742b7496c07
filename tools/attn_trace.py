"""Dump the pipeline timeline of one attention CTA (gims_debug_attention_trace) for a 2048x2048 layer."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gims_b200 import GMatcher, _lib
from gims_b200.synth import make_state_dict

L = _lib.lib()
dev = torch.device('cuda')
gm = GMatcher({}); gm.load_state_dict(make_state_dict(7)); gm = gm.cuda().eval()
model = gm.handle()
n0 = n1 = 2048
desc = torch.randn(n0 + n1, 256, device=dev)
nd = torch.tensor([n0, n1], dtype=torch.int32, device=dev)
scratch = torch.zeros(L.gims_attn_scratch_floats(n0 + n1), device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
trace = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
for it in range(3):
    _lib.check(L.gims_attn_layer_forward(model, 0, _lib.ptr(desc), n0, n1, _lib.ptr(nd), _lib.ptr(scratch), None, st), 'warm')
torch.cuda.synchronize()
L.gims_debug_attention_trace(C.c_void_p(trace.data_ptr()))
_lib.check(L.gims_attn_layer_forward(model, 0, _lib.ptr(desc), n0, n1, _lib.ptr(nd), _lib.ptr(scratch), None, st), 'trace')
torch.cuda.synchronize()
L.gims_debug_attention_trace(None)
t = trace.cpu().view(64, 8)
t0 = int(t[0, 0])
x = [int(v) for v in t[40]]
print('kernel entry -> first tile top: %d clk; last stamp -> exit: see below; entry -> exit %d clk = %d ns (globaltimer)' %
      (t0 - x[0], x[1] - x[0], x[3] - x[2]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(20):
    _lib.check(L.gims_attn_layer_forward(model, 0, _lib.ptr(desc), n0, n1, _lib.ptr(nd), _lib.ptr(scratch), None, st), 'time')
e1.record()
torch.cuda.synchronize()
print('whole layer (QKV GEMM + attention + MLP1 + MLP2), back to back: %.1f us' % (e0.elapsed_time(e1) * 1000 / 20))
print('tile  qk_issue  pv_ready  pv_issued | s_seen(w0)  p_given by warp 0,1,2,3   (cycles since first stamp)')
prev = None
for j in range(32):
    r = [int(x) - t0 if int(x) else -1 for x in t[j]]
    print('%3d  top %7d  k_ready %7d  qk_issued %7d  pv_ready %7d  pv_issued %7d | s_seen %7d p_given %7d' % (j, r[0], r[7], r[6], r[1], r[2], r[4], r[3]))
    prev = r

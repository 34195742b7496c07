"""Pipeline timeline of one k_attention_f16 CTA (gims_debug_attention_trace) for an n x n layer: python tools/attn16_trace.py [n]"""
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
n0 = n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
desc = torch.randn(n0 + n1, 256, device=dev)
nd = torch.tensor([n0, n1], dtype=torch.int32, device=dev)
scratch = torch.zeros(L.gims_attn_scratch_floats(n0 + n1), device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
trace = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
def layer(tag):
    _lib.check(L.gims_attn_layer_forward(model, 0, _lib.ptr(desc), n0, n1, _lib.ptr(nd), _lib.ptr(scratch), None, st), tag)
for it in range(3):
    layer('warm')
torch.cuda.synchronize()
L.gims_debug_attention_trace(C.c_void_p(trace.data_ptr()))
layer('trace')
torch.cuda.synchronize()
L.gims_debug_attention_trace(None)
for cls in ('attention', 'gemm'):
    L.gims_profile_begin(_lib.PROF[cls], 64)
    for it in range(10):
        layer('time')
    torch.cuda.synchronize()
    tot, cnt = C.c_double(0), C.c_int(0)
    L.gims_profile_end(C.byref(tot), C.byref(cnt))
    print('%s: %d launches, %.1f us each' % (cls, cnt.value, 1e3 * tot.value / max(1, cnt.value)))
t = trace.cpu().view(64, 16)
t0 = int(t[0, 0])
print('tile | qk_ready qk_issued | pv_ready pv_issued | t0: s_seen s_loaded max_known exp_done o_full p_given folded | t1: s_seen   (cycles since QK(0) ready)')
for j in range(min(40, (n1 + 63) // 64)):
    r = [int(x) - t0 if int(x) else -1 for x in t[j]]
    print('%3d | %7d %7d | %7d %7d | %7d %7d %7d %7d %7d %7d %7d | %7d' % (j, r[0], r[1], r[2], r[3], r[4], r[8], r[9], r[10], r[11], r[5], r[6], r[7]))

"""Debug: one attention layer in fp16x2 mode vs the fp32 CUDA-core path, small sizes, prints progress."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gims_b200 import GMatcher, _lib
from gims_b200.synth import make_state_dict
L = _lib.lib()
dev = torch.device('cuda')
gm = GMatcher({}); gm.load_state_dict(make_state_dict(7)); gm = gm.cuda().eval()
model = gm.handle()
for (n0, n1) in [(128, 64), (256, 256), (512, 470), (2048, 2048)]:
    desc = torch.randn(n0 + n1, 256, device=dev)
    nd = torch.tensor([n0, n1], dtype=torch.int32, device=dev)
    scratch = torch.zeros(L.gims_attn_scratch_floats(n0 + n1), device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = {}
    for mode in (_lib.GEMM_SIMT, _lib.GEMM_TC_F16):
        L.gims_set_gemm_mode(mode)
        d = desc.clone()
        print('launch', n0, n1, 'mode', mode, flush=True)
        _lib.check(L.gims_attn_layer_forward(model, 1, _lib.ptr(d), n0, n1, _lib.ptr(nd), _lib.ptr(scratch), None, st), 'layer')
        torch.cuda.synchronize()
        outs[mode] = d
    err = ((outs[0] - outs[2]).abs().max() / outs[0].abs().max()).item()
    print('n=(%d,%d) rel err f16x2 vs simt %.2e' % (n0, n1, err), flush=True)

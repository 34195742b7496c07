"""cProfile of the reference-facing call (host tensors in, dict out) on one thread: where the host time of a pair goes."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gims_b200 import Matching
from gims_b200.synth import make_pair, make_state_dict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = Matching({})
m.gmodel.load_state_dict(make_state_dict(0))
m = m.eval().to('cuda')
pairs = []
for s in range(8):
    d = make_pair(n, n, seed=100 + s)
    d = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in d.items()}
    d['device'] = 'cuda'
    pairs.append(d)
def one(d):
    with torch.no_grad():
        p = m(dict(d))
    return p['matches0'].cpu()
for d in pairs: one(d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(64): one(pairs[i % 8])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 64
print('single caller: %.3f ms per pair (%.1f pairs/s)' % (dt * 1e3, 1 / dt))
# host time only: enqueue without reading results back
t0 = time.perf_counter()
for i in range(64):
    with torch.no_grad():
        p = m(dict(pairs[i % 8]))
t1 = time.perf_counter()
torch.cuda.synchronize()
print('host time per call without the final read-back: %.3f ms (then %.3f ms to drain)' % ((t1 - t0) / 64 * 1e3, (time.perf_counter() - t1) * 1e3))
pr = cProfile.Profile()
pr.enable()
for i in range(64): one(pairs[i % 8])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28)
print(s.getvalue()[:6000])

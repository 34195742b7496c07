"""Debug: potentials of the CUDA path vs a golden fixture (which entries differ, common-mode or not)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from golden_util import inputs_for, load_golden, weights_for
from gims_b200 import GMatcher
name = sys.argv[1] if len(sys.argv) > 1 else 'fwd_n512_damped'
rec, g = load_golden(name)
data = inputs_for(rec)
m = GMatcher({'sinkhorn_iterations': rec['iters'], 'match_threshold': rec['match_threshold']})
m.load_state_dict(weights_for(rec)); m = m.cuda().eval()
dev = torch.device('cuda')
r = m.run_pair(data['keypoints0'][0].to(dev), data['descriptors0'][0].to(dev), data['scores0'][0].to(dev),
               data['keypoints1'][0].to(dev), data['descriptors1'][0].to(dev), data['scores1'][0].to(dev),
               data['image0'].shape, data['image1'].shape, rec['radius'], rec['percentile'], rec['min_size'], debug=True)
torch.cuda.synchronize()
cnt = r['n_kept_dev'].cpu().numpy(); n0, n1 = int(cnt[0]), int(cnt[1])
print('status %#x' % cnt[6])
u, v = r['u'].cpu().numpy()[:n0 + 1].astype(np.float64), r['v'].cpu().numpy()[:n1 + 1].astype(np.float64)
du, dv = u - g['u'], v - g['v']
for nm, d, ref in (('u', du, g['u']), ('v', dv, g['v'])):
    i = np.abs(d).argmax()
    print('%s: mean %.3e std %.3e min %.3e max %.3e | worst idx %d (of %d) value %.4f | last entry diff %.3e value %.4f' %
          (nm, d.mean(), d.std(), d.min(), d.max(), i, len(d), ref[i], d[-1], ref[-1]))
print('du[:-1] mean %.3e  dv[:-1] mean %.3e  sum of means %.3e' % (du[:-1].mean(), dv[:-1].mean(), du[:-1].mean() + dv[:-1].mean()))
# float64 Sinkhorn on OUR couplings: which of the two fp32 results is closer to exact arithmetic?
coup = r['couplings'].cpu().numpy()[:n0 + 1, :n1 + 1].astype(np.float64)
norm = -np.log(n0 + n1)
lmu = np.full(n0 + 1, norm); lmu[-1] = np.log(n1) + norm
lnu = np.full(n1 + 1, norm); lnu[-1] = np.log(n0) + norm
def lse(a, axis):
    mx = a.max(axis=axis, keepdims=True)
    return (mx + np.log(np.exp(a - mx).sum(axis=axis, keepdims=True))).squeeze(axis)
uu, vv = np.zeros(n0 + 1), np.zeros(n1 + 1)
for _ in range(rec['iters']):
    uu = lmu - lse(coup + vv[None, :], 1)
    vv = lnu - lse(coup + uu[:, None], 0)
for nm, a, b in (('ours-f64', u - uu, v - vv), ('ref-f64', g['u'] - uu, g['v'] - vv)):
    print('%s: du mean %.3e absmax %.3e | dv mean %.3e absmax %.3e' % (nm, a.mean(), np.abs(a).max(), b.mean(), np.abs(b).max()))

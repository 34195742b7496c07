import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gims_b200 import GMatcher, _lib
from gims_b200.synth import make_pair, make_state_dict
m = GMatcher({'sinkhorn_iterations': 30, 'match_threshold': 0.005}); m.load_state_dict(make_state_dict(0, damped=True)); m = m.cuda().eval()
dev = torch.device('cuda')
sizes = [(300, 280), (513, 450), (128, 190), (700, 700)]
items = []
for k, (a, b) in enumerate(sizes):
    d = make_pair(a, b, seed=700 + k, width=400, height=300)
    items.append((d['keypoints0'][0].to(dev), d['descriptors0'][0].to(dev), d['scores0'][0].to(dev),
                  d['keypoints1'][0].to(dev), d['descriptors1'][0].to(dev), d['scores1'][0].to(dev), d['image0'].shape, d['image1'].shape))
for mode in (_lib.GEMM_SIMT, _lib.GEMM_TC, _lib.GEMM_TC_F16):
    singles = [m.run_pair(*it, debug=True, gemm_mode=mode) for it in items]
    batch = m.run_pairs(items, debug=True, gemm_mode=mode)
    torch.cuda.synchronize()
    for k in range(4):
        a, b = singles[k], batch[k]
        n0, n1 = [int(x) for x in a['n_kept_dev'].cpu()[:2]]
        n0i = sizes[k][0]
        def diff(key):
            x, y = a[key], b[key]
            rows = torch.cat([torch.arange(n0), n0i + torch.arange(n1)]).to(dev)
            return float((x[rows] - y[rows]).abs().max()), float((x - y).abs().max())
        print('mode', mode, 'pair', k, 'N\'', n0, n1, 'desc_in', diff('desc_in'), 'desc_gnn', diff('desc_gnn'), 'mdesc', diff('mdesc'),
              'matches equal', bool(torch.equal(a['matches0'][:n0], b['matches0'][:n0])))

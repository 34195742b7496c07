"""Row f-3 timing: gims_gt_matches + gims_match_counts against the reference's formulation (utils/preprocess_utils.py:98-132:
cdist + two argmin + unique / cat per round) run with torch ops on the same GPU.  `python tools/bench_postprocess.py [n]`"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gims_b200 import postprocess as pp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(1)
k0 = torch.rand(n, 2, generator=g) * torch.tensor([800.0, 600.0])
H = torch.tensor([[1.03, 0.04, 7.0], [-0.03, 0.97, -4.0], [2e-5, -1e-5, 1.0]])
src = torch.cat([k0, torch.ones(n, 1)], 1) @ H.T
k1 = (src[:, :2] / src[:, 2:3])[torch.randperm(n, generator=g)] + torch.randn(n, 2, generator=g)
k0, k1, H = k0.to(dev), k1.to(dev), H.to(dev)


def torch_formulation(a, b, h, thr=3, iters=3):
    m1 = torch.empty(0, dtype=torch.int64, device=dev); m2 = torch.empty(0, dtype=torch.int64, device=dev)
    miss1 = torch.arange(len(a), device=dev); miss2 = torch.arange(len(b), device=dev)
    s = torch.cat([a, torch.ones(len(a), 1, device=dev)], -1)
    d = (h @ s.T).T
    proj = (d / d[:, 2:3])[:, :2]
    for _ in range(iters):
        x, y = proj[miss1], b[miss2]
        dist = torch.sqrt(((x[:, None, :] - y[None, :, :]) ** 2).sum(-1))
        mn1, mn2 = torch.argmin(dist, 1), torch.argmin(dist, 0)
        i2 = torch.where(mn1[mn2] == torch.arange(len(mn2), device=dev))[0]
        i1 = mn2[i2]
        ok = dist[i1, i2] < thr
        i1, i2 = i1[ok], i2[ok]
        a1, a2 = miss1[i1], miss2[i2]
        def sd(u, v):
            q, c = torch.cat((u, v)).unique(return_counts=True)
            return q[c == 1]
        miss1, miss2 = sd(miss1, a1), sd(miss2, a2)
        m1, m2 = torch.cat((m1, a1)), torch.cat((m2, a2))
    return m1, m2, miss1, miss2


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


a = pp.torch_find_matches(k0, k1, H, 3, 3)
b = torch_formulation(k0, k1, H)
print('n = %d: %d ground-truth pairs; lists identical: %s' % (n, len(a[0]), all(torch.equal(x, y) for x, y in zip(a, b))))
print('gims_b200.postprocess.torch_find_matches: %.3f ms   torch formulation on the same GPU: %.3f ms' %
      (timeit(lambda: pp.torch_find_matches(k0, k1, H, 3, 3)), timeit(lambda: torch_formulation(k0, k1, H))))
gt0, _, _ = pp.gt_match_vector(k0, k1, H, 3, 3)
m = gt0.long().clone(); m[::5] = -1
pred = {'matches0': m[None]}
print('gt_match_vector + precision_recall: %.3f ms' % timeit(lambda: pp.precision_recall(pred, pp.gt_match_vector(k0, k1, H, 3, 3)[0])))

"""Row f-3: device-side ground-truth matching and precision / recall against the oracle's restatement
of utils/preprocess_utils.py:98-132 and eval_homography.py:209-228, through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(n0, n1, seed, noise):
    from oracle import gims_oracle as orc
    g = torch.Generator().manual_seed(seed)
    k0 = torch.rand(n0, 2, generator=g) * torch.tensor([800.0, 600.0])
    H = torch.tensor([[1.03, 0.04, 7.0], [-0.03, 0.97, -4.0], [2e-5, -1e-5, 1.0]])
    pick = torch.randperm(n0, generator=g)[:min(n0, n1)]
    k1 = orc.warp_keypoints(k0, H)[pick] + torch.randn(len(pick), 2, generator=g) * noise
    if n1 > len(pick):
        k1 = torch.cat([k1, torch.rand(n1 - len(pick), 2, generator=g) * torch.tensor([800.0, 600.0])])
    return k0, k1, H


@pytest.mark.parametrize('n0,n1,seed,noise,iters', [(300, 260, 1, 1.5, 3), (2048, 2048, 2, 1.0, 3), (1, 5, 3, 0.1, 1),
                                                     (700, 1500, 4, 2.5, 1), (4096, 3900, 5, 0.7, 2)])
def test_gt_matches_vs_oracle(n0, n1, seed, noise, iters):
    from gims_b200 import postprocess as pp
    from oracle import gims_oracle as orc
    k0, k1, H = _case(n0, n1, seed, noise)
    ref = orc.find_gt_matches(k0, k1, H, dist_thresh=3, n_iters=iters)
    got = pp.torch_find_matches(k0.cuda(), k1.cuda(), H.cuda(), dist_thresh=3, n_iters=iters)
    assert all(t.dtype == torch.int64 and t.is_cuda for t in got)
    # dense form: the projection's 3-term dot products may round differently from the host BLAS, which can move a pair
    # across the 3-pixel threshold or flip an exact tie; everything else is identical
    gt_ref = torch.full((n0,), -1, dtype=torch.int64)
    gt_ref[ref[0]] = ref[1]
    gt_got = torch.full((n0,), -1, dtype=torch.int64)
    gt_got[got[0].cpu()] = got[1].cpu()
    agree = (gt_ref == gt_got).float().mean().item()
    print('\n[gt matches %dx%d it=%d] %d reference pairs, agreement %.5f, lists identical: %s' %
          (n0, n1, iters, len(ref[0]), agree, all(torch.equal(a, b.cpu()) for a, b in zip(ref, got))))
    assert agree >= 0.998
    if agree == 1.0:                       # then the ordered lists and the missing sets are the reference's too
        for a, b in zip(ref, got):
            assert torch.equal(a, b.cpu())
    assert len(got[0]) + len(got[2]) == n0 and len(got[1]) + len(got[3]) == n1


def test_precision_recall_vs_oracle():
    from gims_b200 import postprocess as pp
    from oracle import gims_oracle as orc
    k0, k1, H = _case(1500, 1400, 9, 1.0)
    ma0, ma1, _, _ = orc.find_gt_matches(k0, k1, H, dist_thresh=3, n_iters=3)
    g = torch.Generator().manual_seed(10)
    matches = torch.full((1500,), -1, dtype=torch.int64)
    matches[ma0] = ma1                                      # start from the truth ...
    wrong = torch.randperm(1500, generator=g)[:300]
    matches[wrong] = torch.randint(-1, 1400, (300,), generator=g)   # ... then corrupt a fifth of it
    p_ref, r_ref = orc.precision_recall(matches.numpy(), ma0.numpy(), ma1.numpy())
    gt0, _, _ = pp.gt_match_vector(k0.cuda(), k1.cuda(), H.cuda(), 3, 3)
    assert torch.equal(gt0.cpu().long()[ma0], ma1)
    p, r, counts = pp.precision_recall({'matches0': matches.cuda()[None]}, gt0)
    assert abs(float(p) - p_ref) < 1e-12 and abs(float(r) - r_ref) < 1e-12
    assert int(counts[1]) == int((matches > -1).sum())


def test_matched_points_and_no_cpu_fallback():
    from gims_b200 import postprocess as pp
    from gims_b200._lib import GimsError
    pred = {'keypoints0': torch.arange(12.).reshape(1, 6, 2).cuda(), 'keypoints1': torch.arange(10.).reshape(1, 5, 2).cuda() * 2,
            'matches0': torch.tensor([[2, -1, 0, -1, 4, 1]]).cuda(), 'matching_scores0': torch.rand(1, 6).cuda()}
    a, b, c = pp.matched_points(pred)
    assert a.is_cuda and torch.equal(a.cpu(), pred['keypoints0'][0].cpu()[[0, 2, 4, 5]])
    assert torch.equal(b.cpu(), pred['keypoints1'][0].cpu()[[2, 0, 4, 1]]) and c.shape == (4,)
    with pytest.raises(GimsError):
        pp.torch_find_matches(torch.rand(4, 2), torch.rand(4, 2), torch.eye(3))

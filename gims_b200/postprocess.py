"""Caller-side result consumption on the device — what eval_homography.py:180-229 does with a
matcher result, without moving it to the host first.

  torch_find_matches(kpts0, kpts1, homography, dist_thresh, n_iters)
      same signature and results as utils/preprocess_utils.py:98-132: (match_list_1, match_list_2, missing_1, missing_2),
      int64 device tensors, the pairs ordered by round and then by image-1 index as the reference produces them.
      One C-ABI call (gims_gt_matches: projection + n_iters rounds of mutual nearest neighbours) instead of
      n_iters x (N0 x N1 distance matrix, two argmin, unique / cat bookkeeping).
  gt_match_vector(...)   the dense form eval_homography.py:210-211 builds: gt[i] = partner index in image 1 or -1.
  precision_recall(pred, gt0)   eval_homography.py:224-228 as three device-side counts.
  matched_points(pred)   eval_homography.py:187-189: the matched keypoint pairs and confidences, still on the device.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GimsError


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def gt_match_vector(kpts0, kpts1, homography, dist_thresh=3, n_iters=1):
    """(gt0 int32 [N0], gt1 int32 [N1], round0 int32 [N0]) on the device of `kpts0`."""
    dev = kpts0.device
    if dev.type != 'cuda':
        raise GimsError('gims_b200.postprocess runs on a CUDA device only (no CPU fallback)')
    L = _lib.lib()
    k0 = kpts0.detach().to(torch.float32).contiguous()
    k1 = kpts1.detach().to(device=dev, dtype=torch.float32).contiguous()
    n0, n1 = k0.shape[0], k1.shape[0]
    if n0 < 1 or n1 < 1:
        raise GimsError('torch_find_matches: empty keypoint set')          # the reference raises inside torch.argmin
    h = torch.as_tensor(homography).detach().to(device=dev, dtype=torch.float32).contiguous().reshape(9)
    ws = torch.empty(L.gims_gt_workspace_bytes(n0, n1), dtype=torch.uint8, device=dev)
    gt0 = torch.empty(n0, dtype=torch.int32, device=dev)
    gt1 = torch.empty(n1, dtype=torch.int32, device=dev)
    rnd = torch.empty(n0, dtype=torch.int32, device=dev)
    _lib.check(L.gims_gt_matches(_lib.ptr(k0), n0, _lib.ptr(k1), n1, _lib.ptr(h), float(dist_thresh), int(n_iters),
                                 _lib.ptr(ws), ws.numel(), _lib.ptr(gt0), _lib.ptr(gt1), _lib.ptr(rnd), _stream(dev)),
               'gims_gt_matches')
    return gt0, gt1, rnd


def torch_find_matches(src_keypoints1, src_keypoints2, homography, dist_thresh=3, n_iters=1):
    """Drop-in for utils/preprocess_utils.py:98-132."""
    gt0, gt1, rnd = gt_match_vector(src_keypoints1, src_keypoints2, homography, dist_thresh, n_iters)
    n1 = gt1.numel()
    idx0 = torch.nonzero(gt0 >= 0).squeeze(1)
    partner = gt0[idx0].long()
    order = torch.argsort(rnd[idx0].long() * n1 + partner)                # by round, then by image-1 index
    match_1, match_2 = idx0[order], partner[order]
    missing_1 = torch.nonzero(gt0 < 0).squeeze(1)
    missing_2 = torch.nonzero(gt1 < 0).squeeze(1)
    return match_1, match_2, missing_1, missing_2


def precision_recall(pred, gt0):
    """eval_homography.py:224-228 for one pair: `pred` is the matcher's output dict (matches0 (1, N0') int64 on the device),
    `gt0` the vector of gt_match_vector.  Returns (precision, recall, counts) — counts = device int32 [tp, predicted, missed];
    the two ratios are device scalars (0-d tensors), nothing is synchronised here."""
    m0 = pred['matches0'] if torch.is_tensor(pred) is False else pred
    m0 = m0.reshape(-1).contiguous()
    dev = m0.device
    if dev.type != 'cuda':
        raise GimsError('gims_b200.postprocess runs on a CUDA device only (no CPU fallback)')
    if m0.dtype != torch.int64 or gt0.dtype != torch.int32 or gt0.numel() != m0.numel():
        raise GimsError('precision_recall: matches0 must be int64 and gt0 int32 of the same length')
    L = _lib.lib()
    counts = torch.empty(3, dtype=torch.int32, device=dev)
    _lib.check(L.gims_match_counts(_lib.ptr(m0), _lib.ptr(gt0.contiguous()), m0.numel(), None, _lib.ptr(counts), _stream(dev)),
               'gims_match_counts')
    c = counts.to(torch.float64)
    return c[0] / c[1], c[0] / (c[0] + c[2]), counts


def matched_points(pred):
    """eval_homography.py:187-189 on the device: (mkpts0, mkpts1, mconf) of the pairs with matches0 > -1."""
    kpts0, kpts1 = pred['keypoints0'][0], pred['keypoints1'][0]
    matches, conf = pred['matches0'][0], pred['matching_scores0'][0]
    valid = matches > -1
    return kpts0[valid], kpts1[matches[valid]], conf[valid]

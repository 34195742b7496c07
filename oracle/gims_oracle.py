"""TEST INFRASTRUCTURE — CPU restatement (numpy + torch-CPU fp32) of the GIMS matcher forward path.

This file is the parity oracle.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it; the product (`gims_b200/`) never does.

It restates, function by function, the reference algorithm of
  /root/reference/models/agc.py:367-391, 413-449, 476-565, 660-709   (adaptive graph construction)
  /root/reference/models/gmatcher.py:11-162, 219-307                  (SAGE, kenc, attention, OT, matches)
plus the two DGL call semantics the reference relies on (`dgl==1.1.2`, un-vendored; see
oracle/ref_shims.py for the published behaviour that is restated).

PARITY PINNING: the reference has no tests / golden vectors for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container:
`oracle/make_golden.py` imports the unmodified reference under import shims and commits its
outputs as fixtures in `tests/golden/`; `tests/test_oracle_golden.py` checks this restatement
against them.  The DGL part of that pin is the shim's restatement, not real DGL.

NUMERIC CONTRACT of the graph builder ("bit-exact under a stated tie-break").
Where the reference's result depends on library-internal floating-point summation order or on
unspecified tie-breaks, this oracle fixes ONE deterministic definition that the CUDA path must
reproduce bit for bit:
  C1  row normalisation: n_i = fl32(sqrt(sum_k x_ik^2 in fp64)); xhat_ik = fl32(x_ik / max(n_i, 1e-12))
      (reference: torch F.normalize fp32, agc.py:389 — may differ in the last ulp of n_i).
  C2  cosine: S_ij = fl32(sum_k xhat_ik * xhat_jk accumulated in fp64)  (reference: MKL sgemm,
      agc.py:390 — differs by fp32 summation order, |d| ~ 1e-7).
  C3  threshold: exact k-th smallest of the strict upper triangle, k = min(int(L*p/100), L-1)
      (agc.py:377-379, 439-440); edge iff S_ij >= thr (agc.py:446).
  C4  radius test: fp64 dx^2+dy^2 <= r^2 on the fp32 coordinates, inclusive (cKDTree.query_pairs).
  C5  nearest neighbour (agc.py:489, 547, 560-562): fp64 squared distance, ties -> lowest index.
  C6  component centroid (agc.py:542-543): mean accumulated in fp64, rounded to fp32 (reference:
      fp32 running sum in Python-set iteration order, which is itself not stable across runs).
  C7  closest pair between components i and j=nn(i) (agc.py:558-564): argmin over (d2, v, u)
      lexicographically, u in comp_i, v in comp_j.
  C8  CSR: nodes relabelled in ascending original id (agc.py:677, dgl.from_networkx 'sorted'),
      neighbour lists ascending, both directions of every edge.
Results equal the reference's except on exact ties / |S_ij - thr| below fp32 summation noise;
`tests/test_oracle_golden.py` shows they are identical on every committed fixture.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

NUM_HEADS = 4            # gmatcher.py:131
BN_EPS = 1e-5


# --------------------------------------------------------------------------------------------
# a-1 .. a-6  Adaptive graph construction
# --------------------------------------------------------------------------------------------
def normalize_rows(descs):
    """C1 — agc.py:388-389."""
    x = np.ascontiguousarray(descs, dtype=np.float32)
    n = np.sqrt((x.astype(np.float64) ** 2).sum(1)).astype(np.float32)
    return (x / np.maximum(n, np.float32(1e-12))[:, None]).astype(np.float32)


def cosine_matrix(descs):
    """C2 — agc.py:382-391 (`fast_cosine_similarity_matrix`)."""
    xh = normalize_rows(descs).astype(np.float64)
    return (xh @ xh.T).astype(np.float32)


def percentile_threshold(sim, percentile):
    """C3 — agc.py:367-380, 439-440 (`fast_percentile_threshold` on the strict upper triangle)."""
    vals = sim[np.triu_indices_from(sim, k=1)]
    k = int(len(vals) * percentile / 100)
    if k >= len(vals):
        k = len(vals) - 1
    return np.partition(vals, k)[k], k


def _sqdist(p, q):
    """fp64 squared distances between fp32 point sets p (a,2), q (b,2)."""
    p = p.astype(np.float64)
    q = q.astype(np.float64)
    dx = p[:, None, 0] - q[None, :, 0]
    dy = p[:, None, 1] - q[None, :, 1]
    return dx * dx + dy * dy


def _components(n, und_edges):
    """Connected components; label = smallest member id (networkx order, agc.py:511, 535)."""
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in und_edges:
        ra, rb = find(a), find(b)
        if ra != rb:
            if ra < rb:
                parent[rb] = ra
            else:
                parent[ra] = rb
    return np.array([find(i) for i in range(n)], dtype=np.int64)


def agc_build(kpts, descs, radius, percentile, min_size):
    """agc.py:682-709 `build_optimize_graph_with_cosine_similarity` for one image.

    kpts (N,2) fp32, descs (N,D) fp32.  Returns dict: kept (N',) int64 ascending original ids,
    indptr (N'+1,), indices (E,) int64 relabelled CSR (C8), thr fp32, n_components (of the cleaned
    graph, the number agc.py:540 prints), edges_base / edges_iso / edges_comp (original ids).
    """
    kpts = np.ascontiguousarray(kpts, dtype=np.float32)
    n = kpts.shape[0]
    sim = cosine_matrix(descs)
    thr, _ = percentile_threshold(sim, percentile)
    d2 = _sqdist(kpts, kpts)
    adj = (d2 <= float(radius) * float(radius)) & (sim >= thr)          # C4, agc.py:435-447
    np.fill_diagonal(adj, False)
    iu, ju = np.nonzero(np.triu(adj, 1))
    edges_base = list(zip(iu.tolist(), ju.tolist()))

    # a-3 connect_isolated_nodes (agc.py:476-495): live degree, ascending node order
    edges_iso = []
    if n > 0 and len(edges_base) > 0:
        deg = adj.sum(1).astype(np.int64)
        for i in np.nonzero(deg == 0)[0].tolist():
            if deg[i] != 0:
                continue
            row = d2[i].copy()
            row[i] = np.inf
            j = int(np.argmin(row))                                       # C5
            edges_iso.append((i, j))
            deg[i] += 1
            deg[j] += 1

    # a-4 remove_small_components (agc.py:497-516)
    label = _components(n, edges_base + edges_iso)
    sizes = np.bincount(label, minlength=n)
    keep_mask = sizes[label] >= min_size
    kept = np.nonzero(keep_mask)[0].astype(np.int64)

    # a-5 fast_connect_components (agc.py:518-565), one round on the cleaned graph
    roots = np.unique(label[kept])                                        # ascending smallest-member id
    edges_comp = []
    if len(roots) > 1:
        members = [kept[label[kept] == r] for r in roots]
        cent = np.stack([kpts[m].astype(np.float64).mean(0) for m in members]).astype(np.float32)  # C6
        cd = _sqdist(cent, cent)
        np.fill_diagonal(cd, np.inf)
        nn = cd.argmin(1)                                                 # C5
        for i, j in enumerate(nn.tolist()):
            if j < i and nn[j] == i:                                      # (j,i) already connected, agc.py:550-551
                continue
            dd = _sqdist(kpts[members[j]], kpts[members[i]])             # rows v in comp_j, cols u in comp_i
            flat = int(np.argmin(dd))                                     # C7: first min in (v,u) order
            v = int(members[j][flat // dd.shape[1]])
            u = int(members[i][flat % dd.shape[1]])
            edges_comp.append((u, v))

    # a-6 relabel + CSR (agc.py:699-708, dgl.from_networkx)
    new_id = -np.ones(n, dtype=np.int64)
    new_id[kept] = np.arange(len(kept))
    nbrs = [[] for _ in range(len(kept))]
    for a, b in edges_base + edges_iso + edges_comp:
        if keep_mask[a] and keep_mask[b]:
            nbrs[new_id[a]].append(int(new_id[b]))
            nbrs[new_id[b]].append(int(new_id[a]))
    indptr = np.zeros(len(kept) + 1, dtype=np.int64)
    for i, l in enumerate(nbrs):
        l.sort()
        indptr[i + 1] = indptr[i] + len(l)
    indices = np.array([x for l in nbrs for x in l], dtype=np.int64)
    return {'kept': kept, 'indptr': indptr, 'indices': indices, 'thr': np.float32(thr),
            'n_components': int(len(roots)), 'edges_base': edges_base, 'edges_iso': edges_iso,
            'edges_comp': edges_comp}


# --------------------------------------------------------------------------------------------
# a-8 .. a-15  network forward (torch CPU fp32, same op sequence as the reference)
# --------------------------------------------------------------------------------------------
def normalize_keypoints(kpts, image_shape):
    """gmatcher.py:26-33 — shape is taken positionally (`_, _, height, width = image_shape`)."""
    _, _, height, width = image_shape
    one = kpts.new_tensor(1)
    size = torch.stack([one * width, one * height])[None]
    center = size / 2
    scaling = size.max(1, keepdim=True).values * 0.7
    return (kpts - center[:, None, :]) / scaling[:, None, :]


def _sage_bias(sd, i):
    k = 'gnn_encoder.layers.%d.bias' % i
    return sd[k] if k in sd else sd['gnn_encoder.layers.%d.fc_self.bias' % i]


def sage_forward(sd, indptr, indices, feat):
    """gmatcher.py:145-162 + DGL SAGEConv('mean') — feat (N',D) -> (N',D)."""
    indptr = torch.as_tensor(indptr, dtype=torch.int64)
    indices = torch.as_tensor(indices, dtype=torch.int64)
    n = feat.shape[0]
    deg = (indptr[1:] - indptr[:-1])
    dst = torch.repeat_interleave(torch.arange(n), deg)
    degf = deg.to(feat.dtype).clamp(min=1).unsqueeze(1)
    h = feat
    for i in range(3):
        w_n = sd['gnn_encoder.layers.%d.fc_neigh.weight' % i]
        w_s = sd['gnn_encoder.layers.%d.fc_self.weight' % i]
        before = w_n.shape[1] > w_n.shape[0]
        msg = F.linear(h, w_n) if before else h
        agg = torch.zeros(n, msg.shape[1], dtype=h.dtype)
        agg.index_add_(0, dst, msg[indices])
        neigh = agg / degf
        if not before:
            neigh = F.linear(neigh, w_n)
        h = F.linear(h, w_s) + neigh + _sage_bias(sd, i)
        if i != 2:
            h = F.relu(h)
    return h


def _mlp(sd, prefix, idx_list, x):
    """gmatcher.py:11-24 — Conv1d(k=1) [+ BatchNorm1d(eval) + ReLU] chain; x (B,C,N)."""
    last = idx_list[-1]
    for ci in idx_list:
        x = F.conv1d(x, sd['%s.%d.weight' % (prefix, ci)], sd['%s.%d.bias' % (prefix, ci)])
        if ci != last:
            b = '%s.%d' % (prefix, ci + 1)
            x = F.batch_norm(x, sd[b + '.running_mean'], sd[b + '.running_var'], sd[b + '.weight'],
                             sd[b + '.bias'], False, 0.0, BN_EPS)
            x = F.relu(x)
    return x


def kenc_forward(sd, kpts_norm):
    """gmatcher.py:87-97 — (B,N,2) -> (B,D,N); scores unused (score=False, gmatcher.py:184)."""
    n_conv = len([k for k in sd if k.startswith('kenc.encoder.') and k.endswith('.weight')
                  and sd[k].dim() == 3])
    return _mlp(sd, 'kenc.encoder', [3 * i for i in range(n_conv)], kpts_norm.transpose(1, 2))


def attention(query, key, value):
    """gmatcher.py:35-39."""
    dim = query.shape[1]
    scores = torch.einsum('bdhn,bdhm->bhnm', query, key) / dim ** .5
    prob = F.softmax(scores, dim=-1)
    return torch.einsum('bhnm,bdhm->bdhn', prob, value)


def attn_propagation(sd, l, x, source):
    """gmatcher.py:99-125 — one AttentionalPropagation call; returns delta (B,D,N)."""
    p = 'gnn.layers.%d' % l
    b, d = x.shape[0], x.shape[1]
    dim = d // NUM_HEADS
    q, k, v = [F.conv1d(t, sd['%s.attn.proj.%d.weight' % (p, j)], sd['%s.attn.proj.%d.bias' % (p, j)])
               .view(b, dim, NUM_HEADS, -1) for j, t in enumerate((x, source, source))]
    msg = attention(q, k, v).contiguous().view(b, d, -1)
    msg = F.conv1d(msg, sd[p + '.attn.merge.weight'], sd[p + '.attn.merge.bias'])
    return _mlp(sd, p + '.mlp', [0, 3], torch.cat([x, msg], dim=1))


def gnn_forward(sd, names, desc0, desc1, trace=None):
    """gmatcher.py:135-143."""
    for l, name in enumerate(names):
        if name == 'cross':
            src0, src1 = desc1, desc0
        else:
            src0, src1 = desc0, desc1
        delta0, delta1 = attn_propagation(sd, l, desc0, src0), attn_propagation(sd, l, desc1, src1)
        desc0, desc1 = desc0 + delta0, desc1 + delta1
        if trace is not None:
            trace.append((desc0, desc1))
    return desc0, desc1


def log_optimal_transport(scores, alpha, iters):
    """gmatcher.py:41-69; additionally returns the potentials u, v and the couplings (b, m+1, n+1)."""
    b, m, n = scores.shape
    one = scores.new_tensor(1)
    ms, ns = (m * one).to(scores), (n * one).to(scores)
    bins0 = alpha.expand(b, m, 1)
    bins1 = alpha.expand(b, 1, n)
    alpha = alpha.expand(b, 1, 1)
    couplings = torch.cat([torch.cat([scores, bins0], -1), torch.cat([bins1, alpha], -1)], 1)
    norm = -(ms + ns).log()
    log_mu = torch.cat([norm.expand(m), ns.log()[None] + norm])[None].expand(b, -1)
    log_nu = torch.cat([norm.expand(n), ms.log()[None] + norm])[None].expand(b, -1)
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(couplings + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(couplings + u.unsqueeze(2), dim=1)
    z = couplings + u.unsqueeze(2) + v.unsqueeze(1) - norm
    return z, u, v, couplings


def extract_matches(z, match_threshold):
    """gmatcher.py:284-294.  Also returns the pre-threshold argmax indices."""
    max0, max1 = z[:, :-1, :-1].max(2), z[:, :-1, :-1].max(1)
    indices0, indices1 = max0.indices, max1.indices
    ar0 = torch.arange(indices0.shape[1])[None]
    ar1 = torch.arange(indices1.shape[1])[None]
    mutual0 = ar0 == indices1.gather(1, indices0)
    mutual1 = ar1 == indices0.gather(1, indices1)
    zero = z.new_tensor(0)
    mscores0 = torch.where(mutual0, max0.values.exp(), zero)
    mscores1 = torch.where(mutual1, mscores0.gather(1, indices1), zero)
    valid0 = mutual0 & (mscores0 > match_threshold)
    valid1 = mutual1 & valid0.gather(1, indices1)
    matches0 = torch.where(valid0, indices0, indices0.new_tensor(-1))
    matches1 = torch.where(valid1, indices1, indices1.new_tensor(-1))
    return {'matches0': matches0, 'matches1': matches1, 'matching_scores0': mscores0,
            'matching_scores1': mscores1, 'indices0': indices0, 'indices1': indices1,
            'mutual0': mutual0, 'mutual1': mutual1}


DEFAULT_CONFIG = {
    'descriptor_dim': 256, 'keypoint_encoder': [32, 64, 128, 256],
    'transformer_layers': ['self', 'cross'] * 9, 'sinkhorn_iterations': 100, 'match_threshold': 0.2,
}


def gmatcher_forward(sd, data, config=None, stages=False, timings=None):
    """gmatcher.py:219-307 `GMatcher.forward` (test mode, B=1, dynamic-threshold graph).

    `data` as for the reference (keypoints* (1,N,2), descriptors* (1,D,N), scores* (1,N), image*
    — only `.shape` is used —, optional radius/percentile/min_size).  Returns the reference's
    output dict; with `stages=True` also every intermediate the parity tests compare.
    """
    import time
    cfg = {**DEFAULT_CONFIG, **(config or {})}
    radius = data.get('radius', 25)
    percentile = data.get('percentile', 7)
    min_size = data.get('min_size', 8)
    tt = time.perf_counter()
    g = []
    for s in ('0', '1'):
        kp = data['keypoints' + s][0].numpy()
        de = data['descriptors' + s][0].t().contiguous().numpy()
        g.append(agc_build(kp, de, radius, percentile, min_size))
    t_agc = time.perf_counter() - tt
    out = {}
    kp, de, sc = [], [], []
    for s, gi in zip(('0', '1'), g):
        kept = torch.from_numpy(gi['kept'])
        kp.append(data['keypoints' + s][0][kept][None])
        de.append(data['descriptors' + s][0].t()[kept].contiguous())       # (N',D) == ndata['feat']
        sc.append(data['scores' + s][0][kept][None])
    out['keypoints0'], out['keypoints1'] = kp
    out['descriptors0'], out['descriptors1'] = de[0][None].permute(0, 2, 1), de[1][None].permute(0, 2, 1)
    if kp[0].shape[1] == 0 or kp[1].shape[1] == 0:                          # gmatcher.py:257-264
        shape0, shape1 = kp[0].shape[:-1], kp[1].shape[:-1]
        out.update({'matches0': kp[0].new_full(shape0, -1, dtype=torch.int),
                    'matches1': kp[1].new_full(shape1, -1, dtype=torch.int),
                    'matching_scores0': kp[0].new_zeros(shape0), 'matching_scores1': kp[1].new_zeros(shape1)})
        return out
    tt = time.perf_counter()
    kn0 = normalize_keypoints(kp[0], data['image0'].shape)
    kn1 = normalize_keypoints(kp[1], data['image1'].shape)
    sage0 = sage_forward(sd, g[0]['indptr'], g[0]['indices'], de[0])
    sage1 = sage_forward(sd, g[1]['indptr'], g[1]['indices'], de[1])
    enc0, enc1 = kenc_forward(sd, kn0), kenc_forward(sd, kn1)
    desc0 = sage0[None].permute(0, 2, 1) + enc0
    desc1 = sage1[None].permute(0, 2, 1) + enc1
    t_enc = time.perf_counter() - tt
    tt = time.perf_counter()
    trace = [] if stages else None
    gd0, gd1 = gnn_forward(sd, cfg['transformer_layers'], desc0, desc1, trace)
    t_gnn = time.perf_counter() - tt
    tt = time.perf_counter()
    mdesc0 = F.conv1d(gd0, sd['final_proj.weight'], sd['final_proj.bias'])
    mdesc1 = F.conv1d(gd1, sd['final_proj.weight'], sd['final_proj.bias'])
    scores = torch.einsum('bdn,bdm->bnm', mdesc0, mdesc1) / cfg['descriptor_dim'] ** .5
    t_score = time.perf_counter() - tt
    tt = time.perf_counter()
    z, u, v, _ = log_optimal_transport(scores, sd['bin_score'], cfg['sinkhorn_iterations'])
    t_ot = time.perf_counter() - tt
    tt = time.perf_counter()
    m = extract_matches(z, cfg['match_threshold'])
    t_match = time.perf_counter() - tt
    if timings is not None:
        timings.update({'agc': t_agc, 'sage_kenc': t_enc, 'attention': t_gnn, 'score': t_score,
                        'sinkhorn': t_ot, 'match': t_match})
    out.update({k: m[k] for k in ('matches0', 'matches1', 'matching_scores0', 'matching_scores1')})
    out['mdesc0'] = mdesc0.permute(0, 2, 1).squeeze()
    out['mdesc1'] = mdesc1.permute(0, 2, 1).squeeze()
    if stages:
        out['_stages'] = {
            'graph0': g[0], 'graph1': g[1], 'kpts_norm0': kn0, 'kpts_norm1': kn1,
            'sage0': sage0, 'sage1': sage1, 'kenc0': enc0, 'kenc1': enc1,
            'desc_in0': desc0, 'desc_in1': desc1, 'gnn_trace': trace, 'desc_out0': gd0, 'desc_out1': gd1,
            'scores': scores, 'Z': z, 'u': u, 'v': v,
            'indices0': m['indices0'], 'indices1': m['indices1'],
            'mutual0': m['mutual0'], 'mutual1': m['mutual1'],
        }
    return out


# --------------------------------------------------------------------------------------------
# f-3  caller-side result consumption (test infrastructure like everything in this file)
# --------------------------------------------------------------------------------------------
def warp_keypoints(keypoints, homography_mat):
    """utils/preprocess_utils.py:82-96."""
    source = torch.cat([keypoints, torch.ones(len(keypoints), 1)], dim=-1)
    dest = (homography_mat @ source.T).T
    dest = dest / dest[:, 2:3]
    return dest[:, :2]


def find_gt_matches(kpts0, kpts1, homography, dist_thresh=3, n_iters=1):
    """utils/preprocess_utils.py:98-132 `torch_find_matches` (with torch_cdist :74-78 and torch_setdiff1d :80-84)."""
    match1 = torch.empty(0, dtype=torch.int64)
    match2 = torch.empty(0, dtype=torch.int64)
    miss1 = torch.arange(len(kpts0), dtype=torch.long)
    miss2 = torch.arange(len(kpts1), dtype=torch.long)
    proj = warp_keypoints(kpts0, homography)
    for _ in range(n_iters):
        a, b = proj[miss1, :], kpts1[miss2, :]
        distance = torch.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))
        min1 = torch.argmin(distance, 1)
        min2 = torch.argmin(distance, 0)
        inter2 = torch.where(min1[min2] == torch.arange(len(min2)))[0]
        inter1 = min2[inter2]
        ok = distance[inter1, inter2] < dist_thresh
        inter1, inter2 = inter1[ok], inter2[ok]
        m1, m2 = miss1[inter1], miss2[inter2]

        def setdiff(x, y):
            unq, count = torch.cat((x, y)).unique(return_counts=True)
            return unq[count == 1]
        miss1, miss2 = setdiff(miss1, m1), setdiff(miss2, m2)
        match1, match2 = torch.cat((match1, m1)), torch.cat((match2, m2))
    return match1, match2, miss1, miss2


def precision_recall(matches, ma_0, ma_1):
    """eval_homography.py:209-211, 224-228 (numpy): matches = matches0 of the pair."""
    matches = np.asarray(matches)
    gt = np.ones((len(matches),), dtype=np.int64) * -1
    gt[np.asarray(ma_0)] = np.asarray(ma_1)
    valid = matches > -1
    match_flag = (matches[np.asarray(ma_0)] == np.asarray(ma_1))
    precision = match_flag.sum() / valid.sum()
    fn_flag = np.logical_and((matches != gt), (matches == -1))
    recall = match_flag.sum() / (match_flag.sum() + fn_flag.sum())
    return precision, recall

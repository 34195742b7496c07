"""CPU: the C-ABI library loads and exports every symbol include/gims_b200.h declares; host-only
queries work without a GPU; host-side logic (weight packing, k-rank, config mirror)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from gims_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'gims_b200.h')).read()
    declared = set(re.findall(r'GIMS_API[^;]*?\b(gims_\w+)\s*\(', hdr))
    assert len(declared) >= 17
    from gims_b200 import _lib
    assert declared == set(_lib.SIGNATURES), 'ctypes table and header disagree'
    for name in declared:
        assert hasattr(lib, name), name


def test_host_queries(lib):
    assert lib.gims_version() >= 100
    assert lib.gims_agc_workspace_bytes(2048, 131072) > 2048 * 2048 * 4
    assert lib.gims_sinkhorn_workspace_bytes(2048, 2048) > 0
    assert lib.gims_attn_scratch_floats(4096) == 4096 * 256 * 10 + 256 * 256
    from gims_b200 import GMatcher
    m = GMatcher({})
    c = m.c_config()
    assert lib.gims_packed_blob_count(C.byref(c)) == 1 + 7 * 5 + 7 * 3 + 21 * 18 + 7
    assert [c.layer_is_cross[i] for i in range(4)] == [0, 1, 0, 1]
    assert [c.kenc_dims[i] for i in range(6)] == [2, 32, 64, 128, 256, 256]


def test_no_cpu_fallback():
    from gims_b200 import GMatcher, _lib
    from gims_b200.synth import make_pair
    m = GMatcher({})
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.GimsError):
        m(make_pair(64, seed=1, width=100, height=100))


def test_state_dict_keys_match_reference_schema():
    from gims_b200 import GMatcher
    from gims_b200.config import state_dict_schema
    m = GMatcher({})
    sch = state_dict_schema()
    sd = m.state_dict()
    assert set(sd) == set(sch)
    for k, shape in sch.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sum(v.numel() for k, v in sd.items() if 'num_batches' not in k and 'running' not in k) == 12187617
    # DGL spelling with the bias on fc_self is accepted too
    alt = {k.replace('gnn_encoder.layers.0.bias', 'gnn_encoder.layers.0.fc_self.bias'): v for k, v in sd.items()}
    m.load_state_dict(alt)


def test_k_rank_matches_reference_arithmetic():
    from gims_b200 import GMatcher
    for n, p in [(2048, 7), (512, 2), (300, 50), (64, 100), (97, 0), (1000, 33.3)]:
        length = n * (n - 1) // 2
        vals = np.zeros(length, dtype=np.float32)
        k = int(len(vals) * p / 100)
        if k >= len(vals):
            k = len(vals) - 1
        assert GMatcher._k_rank(n, p) == k


def test_packing_equals_oracle_layer():
    """BN folding, head de-interleave and Q|K|V stacking: a torch emulation of what the kernels compute
    from the PACKED weights equals the oracle's AttentionalPropagation on the raw state dict."""
    from gims_b200.packing import pack_state_dict
    from gims_b200.synth import make_state_dict
    from oracle import gims_oracle as orc
    sd = make_state_dict(5)
    flat, off, names = pack_state_dict(sd)
    blob = {}
    for i, nme in enumerate(names):
        end = off[i + 1] if i + 1 < len(off) else flat.numel()
        blob[nme] = flat[off[i]:end]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(70, 256, generator=g)
    src = torch.randn(50, 256, generator=g)
    l = 3
    wqkv = blob['l3.wqkv'][:768 * 256].view(768, 256)
    bqkv = blob['l3.bqkv'][:768]
    q = x @ wqkv[:256].t() + bqkv[:256]
    k = src @ wqkv[256:512].t() + bqkv[256:512]
    v = src @ wqkv[512:].t() + bqkv[512:]
    att = torch.empty(70, 256)
    for h in range(4):
        sl = slice(h * 64, (h + 1) * 64)
        p = torch.softmax(q[:, sl] @ k[:, sl].t() / 8.0, dim=-1)
        att[:, sl] = p @ v[:, sl]
    # the merge conv is composed into w1 at pack time: w1 acts on [x | attention output]
    hid = F.relu(torch.cat([x, att], 1) @ blob['l3.w1'][:512 * 512].view(512, 512).t() + blob['l3.b1'][:512])
    delta = hid @ blob['l3.w2'][:256 * 512].view(256, 512).t() + blob['l3.b2'][:256]
    want = orc.attn_propagation(sd, l, x.t()[None], src.t()[None])[0].t()
    assert torch.allclose(delta, want, rtol=1e-4, atol=1e-5)


def test_fp16_weight_planes():
    """split_f16: hi + lo reproduce W * 2^e to max(2^-22 relative, 2^-25 absolute), raw bits survive packing."""
    from gims_b200.packing import pack_state_dict, split_f16
    from gims_b200.synth import make_state_dict
    w = torch.randn(64, 128, generator=torch.Generator().manual_seed(5)) * 0.05
    w[0, :4] = torch.tensor([1e-6, -3e-6, 2e-5, 0.0])          # tiny entries: lo is an fp16 subnormal there
    h, l, sinv = split_f16(w)
    hi = h.view(torch.float16).float().view(64, 128)
    lo = l.view(torch.float16).float().view(64, 128)
    rec = (hi + lo) * sinv
    # error class of the split: 2^-22 relative, or the fp16 subnormal spacing (2^-24 in the scaled domain) for tiny entries
    assert ((rec - w).abs() <= w.abs() * 2.0 ** -21 + 2.0 ** -24 * float(sinv)).all()
    assert float(hi.abs().max()) < 2048 and float(1.0 / sinv) == 2.0 ** round(float(torch.log2(1.0 / sinv)))
    flat, off, names = pack_state_dict(make_state_dict(0))
    i = names.index('l3.w2.h16')
    want = split_f16(flat[off[names.index('l3.w2')]:off[names.index('l3.w2')] + 256 * 512].view(256, 512))[0]
    assert torch.equal(flat[off[i]:off[i] + want.numel()].view(torch.int32), want.view(torch.int32))


def test_c_packer_equals_python_packer(lib):
    """gims_pack_weights (csrc/pack.cu, for integrators without Python) against gims_b200/packing.py on a random state_dict
    with non-trivial BatchNorm statistics: same blob offsets; every blob bit-identical except W1 / b1, whose merge composition
    is a 256-term fp64 product summed in a different order (<= 1 fp32 ulp), and the planes derived from W1."""
    from gims_b200 import GMatcher, _lib
    from gims_b200.packing import pack_state_dict
    from gims_b200.synth import make_state_dict
    sd = make_state_dict(3)
    flat, offsets, names = pack_state_dict(sd)
    gm = GMatcher({})
    cfg = gm.c_config()
    items = [(k, v.detach().float().contiguous()) for k, v in sd.items() if v.dtype.is_floating_point]
    arr = (_lib.NamedTensor * len(items))()
    for a, (k, v) in zip(arr, items):
        a.name, a.data, a.numel = k.encode(), v.data_ptr(), v.numel()
    n_floats = lib.gims_pack_weights_floats(C.byref(cfg))
    assert n_floats == flat.numel()
    out = torch.full((n_floats,), float('nan'))
    offs = (C.c_int64 * len(offsets))()
    rc = lib.gims_pack_weights(C.byref(cfg), arr, len(items), C.c_void_p(out.data_ptr()), n_floats, offs, len(offsets))
    assert rc == 0, lib.gims_last_error()
    assert list(offs) == offsets
    bounds = offsets + [flat.numel()]
    for i, name in enumerate(names):
        a, b = flat[bounds[i]:bounds[i + 1]], out[bounds[i]:bounds[i + 1]]
        if '.w1' in name or '.b1' in name:
            if name.endswith('.w1') or name.endswith('.b1') or name.endswith('.hi'):
                assert torch.allclose(a, b, rtol=3e-7, atol=1e-9), name
            elif name.endswith('.lo'):
                assert (a - b).abs().max() <= 1e-6 * flat[bounds[i - 2]:bounds[i - 1]].abs().max(), name
            elif name.endswith('.sinv'):
                assert torch.equal(a, b), name
            else:                                   # fp16 planes: raw bit patterns, compare the values they encode
                ha, hb = a.view(torch.float16).float(), b.view(torch.float16).float()
                if name.endswith('.h16'):
                    assert (ha - hb).abs().max() <= 2.0 ** -10 * ha.abs().max(), name
        else:
            assert torch.equal(a.view(torch.int32), b.view(torch.int32)), name
    # a missing entry is an error that names the key
    rc = lib.gims_pack_weights(C.byref(cfg), arr, len(items) - 1, C.c_void_p(out.data_ptr()), n_floats, offs, len(offsets))
    assert rc != 0 and items[-1][0].encode() in lib.gims_last_error()


def test_c_packer_other_config(lib):
    """A smaller network (4 attention layers, a 3-conv keypoint encoder): blob count, offsets and the exactly reproducible
    blobs of the two packers agree — the C packer derives every shape from the config, not from the default network."""
    from gims_b200 import GMatcher, _lib
    from gims_b200.packing import pack_state_dict
    from gims_b200.synth import make_state_dict
    cfg = {'transformer_layers': ['self', 'cross'] * 2, 'keypoint_encoder': [32, 64]}
    sd = make_state_dict(8, config=cfg)
    flat, offsets, names = pack_state_dict(sd, cfg)
    c = GMatcher(cfg).c_config()
    assert lib.gims_packed_blob_count(C.byref(c)) == len(offsets)
    items = [(k, v.detach().float().contiguous()) for k, v in sd.items() if v.dtype.is_floating_point]
    arr = (_lib.NamedTensor * len(items))()
    for a, (k, v) in zip(arr, items):
        a.name, a.data, a.numel = k.encode(), v.data_ptr(), v.numel()
    n_floats = lib.gims_pack_weights_floats(C.byref(c))
    assert n_floats == flat.numel()
    out = torch.zeros(n_floats)
    offs = (C.c_int64 * len(offsets))()
    assert lib.gims_pack_weights(C.byref(c), arr, len(items), C.c_void_p(out.data_ptr()), n_floats, offs, len(offsets)) == 0
    assert list(offs) == offsets
    bounds = offsets + [flat.numel()]
    for i, name in enumerate(names):
        if '.w1' in name or '.b1' in name:
            continue                                   # merge composition: summation order (see the test above)
        assert torch.equal(flat[bounds[i]:bounds[i + 1]].view(torch.int32), out[bounds[i]:bounds[i + 1]].view(torch.int32)), name


def test_packing_equals_oracle_sage_kenc():
    from gims_b200.packing import pack_state_dict
    from gims_b200.synth import make_state_dict
    from oracle import gims_oracle as orc
    sd = make_state_dict(6)
    flat, off, names = pack_state_dict(sd)
    blob = {nme: flat[off[i]:(off[i + 1] if i + 1 < len(off) else flat.numel())] for i, nme in enumerate(names)}
    g = torch.Generator().manual_seed(1)
    n = 40
    feat = torch.randn(n, 256, generator=g)
    # ring graph + a chord, symmetric CSR
    nbrs = [sorted({(i - 1) % n, (i + 1) % n} | ({20} if i == 0 else set()) | ({0} if i == 20 else set())) for i in range(n)]
    indptr = np.cumsum([0] + [len(a) for a in nbrs])
    indices = np.array([j for a in nbrs for j in a])
    want = orc.sage_forward(sd, indptr, indices, feat)

    def mean(hh):
        return torch.stack([hh[a].sum(0) / len(a) for a in nbrs])
    w0 = blob['sage.w0'][:256 * 256].view(256, 256)
    y0 = feat @ w0.t()
    h1 = F.relu(mean(y0[:, :128]) + y0[:, 128:] + blob['sage.b0'][:128])
    w1 = blob['sage.w1'][:128 * 256].view(128, 256)
    h2 = F.relu(torch.cat([h1, mean(h1)], 1) @ w1.t() + blob['sage.b1'][:128])
    w2 = blob['sage.w2'][:256 * 256].view(256, 256)
    out = torch.cat([h2, mean(h2)], 1) @ w2.t() + blob['sage.b2'][:256]
    assert torch.allclose(out, want, rtol=1e-4, atol=1e-5)
    # kenc with folded BN
    kp = torch.rand(1, n, 2, generator=g) * 2 - 1
    want_k = orc.kenc_forward(sd, kp)[0].t()
    hcur = kp[0]
    dims = [2, 32, 64, 128, 256, 256]
    for i in range(5):
        w = blob['kenc.w%d' % i][:dims[i + 1] * dims[i]].view(dims[i + 1], dims[i])
        hcur = hcur @ w.t() + blob['kenc.b%d' % i][:dims[i + 1]]
        if i < 4:
            hcur = F.relu(hcur)
    assert torch.allclose(hcur, want_k, rtol=1e-4, atol=1e-5)


def test_weight_change_detection():
    """The packed model is rebuilt when weights change: in-place edits move the version sum, load_state_dict /
    .to() drop the cached tensor list (gims_b200/gmatcher.py::_param_version)."""
    from gims_b200 import GMatcher
    from gims_b200.synth import make_state_dict
    gm = GMatcher({})
    gm.load_state_dict(make_state_dict(3))
    v0 = gm._param_version()
    assert gm._param_version() == v0                       # stable without changes
    with torch.no_grad():
        gm.final_proj.weight.mul_(2.0)                     # in-place edit
    v1 = gm._param_version()
    assert v1 != v0
    gm.load_state_dict(make_state_dict(4))                 # copies in place and invalidates
    assert gm.__dict__['_ptensors'] is None
    v2 = gm._param_version()
    assert v2 != v1
    gm = gm.double().float()                               # _apply replaces the tensors: the cached list is dropped
    assert gm.__dict__['_ptensors'] is None
    ts = list(gm.parameters()) + list(gm.buffers())
    assert gm._param_version()[1] == len(ts)


def test_bench_flop_model():
    """bench.py's algorithmic FLOP model (SURVEY.md 8a totals): 256 GFLOP per pair at 2048 surviving keypoints."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('bench', os.path.join(os.path.dirname(__file__), '..', 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    total = bench.pair_flops(2048, 2048)
    assert abs(total / 1e9 - 256.0) < 3.0
    assert bench.projection_flops(2048, 2048) < total

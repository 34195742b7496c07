"""Host-side mirror of the reference's `GMatcher` (models/gmatcher.py:165-307) on top of the C ABI.

Same constructor config, same `state_dict` keys (a reference checkpoint loads unchanged), same
dict-in / dict-out `forward`; every arithmetic step runs in libgims_b200.so (hand-written sm_100a
kernels).  There is no PyTorch / CPU fallback: without the library or a CUDA device this raises.

Only the inference forward with the dynamic-threshold graph is implemented:
`mode='train'` (forward_train, gmatcher.py:309-386) and `delaunay=True` (broken in the reference
snapshot) raise NotImplementedError.
"""
import ctypes as C
import os
from copy import deepcopy

import threading
import time

import torch
import torch.nn as nn

from . import _lib
from .config import DEFAULT_CONFIG, NUM_HEADS, kenc_channels, sage_dims
from .packing import pack_state_dict


# ---- parameter containers with the reference's module tree (never executed) -----------------------
def _mlp_container(channels):
    """Same layer indices as the reference MLP (gmatcher.py:11-24): conv, bn, relu, conv, ..."""
    layers = []
    n = len(channels)
    for i in range(1, n):
        layers.append(nn.Conv1d(channels[i - 1], channels[i], kernel_size=1, bias=True))
        if i < n - 1:
            layers.append(nn.BatchNorm1d(channels[i]))
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


# one lock for the lazily built native state (packed model, per-stream workspaces) of every GMatcher in the process
_HANDLE_LOCK = threading.RLock()

class _KeypointEncoderParams(nn.Module):
    def __init__(self, feature_dim, layers):
        super().__init__()
        self.encoder = _mlp_container([2] + list(layers) + [feature_dim])
        nn.init.constant_(self.encoder[-1].bias, 0.0)


class _MHAParams(nn.Module):
    def __init__(self, d_model):
        super().__init__()
        self.merge = nn.Conv1d(d_model, d_model, kernel_size=1)
        self.proj = nn.ModuleList([deepcopy(self.merge) for _ in range(3)])


class _PropagationParams(nn.Module):
    def __init__(self, feature_dim):
        super().__init__()
        self.attn = _MHAParams(feature_dim)
        self.mlp = _mlp_container([feature_dim * 2, feature_dim * 2, feature_dim])
        nn.init.constant_(self.mlp[-1].bias, 0.0)


class _GNNParams(nn.Module):
    def __init__(self, feature_dim, layer_names):
        super().__init__()
        self.layers = nn.ModuleList([_PropagationParams(feature_dim) for _ in layer_names])
        self.names = list(layer_names)


class _SAGEConvParams(nn.Module):
    """Parameter layout of dgl.nn.SAGEConv(in, out, 'mean') (dgl 1.x): fc_neigh, fc_self, bias."""

    def __init__(self, cin, cout):
        super().__init__()
        self.fc_neigh = nn.Linear(cin, cout, bias=False)
        self.fc_self = nn.Linear(cin, cout, bias=False)
        self.bias = nn.Parameter(torch.zeros(cout))
        gain = nn.init.calculate_gain('relu')
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)


class _GraphSAGEParams(nn.Module):
    def __init__(self, dims):
        super().__init__()
        self.layers = nn.ModuleList([_SAGEConvParams(a, b) for a, b in dims])


class GMatcher(nn.Module):
    """Drop-in for the reference `GMatcher` (inference).  `forward(data)` takes and returns the same dicts."""

    default_config = dict(DEFAULT_CONFIG)

    def __init__(self, config):
        super().__init__()
        self.config = {**self.default_config, **config}
        cfg = self.config
        if cfg['use_layernorm']:
            raise NotImplementedError('use_layernorm=True is outside the hot path (reference default is False)')
        d = cfg['descriptor_dim']
        self.kenc = _KeypointEncoderParams(d, cfg['keypoint_encoder'])
        self.gnn = _GNNParams(d, cfg['transformer_layers'])
        self.gnn_encoder = _GraphSAGEParams(sage_dims(cfg))
        if cfg['input_dim'] != d:        # constructed but unused by forward, as in gmatcher.py:198-201
            self.input_proj = nn.Linear(cfg['input_dim'], d, bias=True)
        else:
            self.input_proj = nn.Identity()
        self.final_proj = nn.Conv1d(d, d, kernel_size=1, bias=True)
        self.register_parameter('bin_score', torch.nn.Parameter(torch.tensor(1.0)))
        self._model = None            # C handle
        self._packed = None           # device buffer the handle points into
        self._packed_key = None
        self._inflight = {}           # handle value -> number of host calls currently inside the library with it
        self._retired = []            # (handle, packed buffer) replaced by a re-pack, destroyed once idle
        self._ws = {}                 # (device, stream slot) -> [workspace tensor, n0 cap, n1 cap, edge cap]
        self._meta_pinned = {}        # (device, stream, thread) -> pinned staging buffer of the per-call metadata
        self._in_pinned = {}          # (device, stream, thread, role) -> [pinned staging buffer of a pageable input, event]
        self.edge_cap_factor = 64     # initial capacity of the CSR edge list = factor * max(n0, n1); grows on overflow
        if cfg['weights_path']:
            weights = torch.load(cfg['weights_path'], map_location='cpu', weights_only=False)
            if ('ema' in weights) and (weights['ema'] is not None):
                load_dict = weights['ema']
            elif 'model' in weights:
                load_dict = weights['model']
            else:
                load_dict = weights
            self.load_state_dict(load_dict)
            print('Loaded GMatcher model ("{}" weights)'.format(cfg['weights_path']))

    # -- state dict compatibility -------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = dict(state_dict)
        for i in range(len(self.gnn_encoder.layers)):       # other DGL releases keep the bias on fc_self
            alt = 'gnn_encoder.layers.%d.fc_self.bias' % i
            if alt in sd and ('gnn_encoder.layers.%d.bias' % i) not in sd:
                sd['gnn_encoder.layers.%d.bias' % i] = sd.pop(alt)
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._invalidate()
        return out

    def _invalidate(self):
        """Drop the packed weights (parameters changed / moved).  Another thread may still be inside the library with
        the current handle, and enqueued kernels may still read the packed buffer: both are only RETIRED here and
        destroyed by `_reap` once no host call holds the handle (after a device synchronisation)."""
        with _HANDLE_LOCK:
            if getattr(self, '_model', None) is not None:
                self._retired.append((self._model, self._packed))
            self._model = None
            self._packed = None
            self._packed_key = None
            self.__dict__['_ptensors'] = None
            self._reap()

    def _reap(self, force=False):
        keep = []
        for h, packed in getattr(self, '_retired', []):
            if not force and self._inflight.get(h.value, 0) > 0:
                keep.append((h, packed))
                continue
            if packed is not None and packed.is_cuda:
                torch.cuda.synchronize(packed.device)       # kernels that read the old weights have finished
            _lib.lib().gims_model_destroy(h)
        self._retired = keep

    def repack(self):
        """Force a re-pack of the weights on the next call.  Needed only after edits that bypass autograd's version
        counter (`param.data[...] = ...`, raw pointer writes); in-place ops, `.to()` and `load_state_dict` are
        detected automatically."""
        self._invalidate()

    def __del__(self):
        try:
            with _HANDLE_LOCK:
                self._invalidate()
                self._reap(force=True)
        except Exception:
            pass

    # -- C model handle -----------------------------------------------------------------------------
    def c_config(self):
        cfg = self.config
        c = _lib.Config()
        c.descriptor_dim = cfg['descriptor_dim']
        names = list(cfg['transformer_layers'])
        if len(names) > _lib.MAX_LAYERS:
            raise ValueError('too many transformer layers')
        c.num_layers = len(names)
        for i, nm in enumerate(names):
            c.layer_is_cross[i] = 1 if nm == 'cross' else 0       # gmatcher.py:137-140: anything else is 'self'
        ch = kenc_channels(cfg)
        c.kenc_num = len(ch) - 1
        for i, v in enumerate(ch):
            c.kenc_dims[i] = v
        c.sinkhorn_iterations = int(cfg['sinkhorn_iterations'])
        c.match_threshold = float(cfg['match_threshold'])
        return c

    def _param_version(self):
        """Cheap change detector for the packed weights: in-place edits bump a tensor's `_version`; `.to()` /
        `load_state_dict` go through `_apply` / `_invalidate`, which drop the cached tensor list.  (Walking the module
        tree and asking every tensor for its data_ptr on each call cost ~0.25 ms of GIL time per forward.)"""
        ts = self.__dict__.get('_ptensors')
        if ts is None:
            ts = list(self.parameters()) + list(self.buffers())
            self.__dict__['_ptensors'] = ts
        v = 0
        for t in ts:
            v += t._version
        return (v, len(ts), ts[0].data_ptr(), ts[-1].data_ptr())

    def handle(self):
        """(Re)pack the weights if needed and return the `gims_model*`.  Safe to call from several threads."""
        with _HANDLE_LOCK:
            return self._handle_locked()

    def _acquire(self):
        """handle() + mark it in use by this host call (see `_invalidate`); pair with `_release`."""
        with _HANDLE_LOCK:
            h = self._handle_locked()
            self._inflight[h.value] = self._inflight.get(h.value, 0) + 1
            return h

    def _release(self, h):
        with _HANDLE_LOCK:
            left = self._inflight.get(h.value, 1) - 1
            if left:
                self._inflight[h.value] = left
            else:
                self._inflight.pop(h.value, None)
                if self._retired:
                    self._reap()

    def _handle_locked(self):
        dev = self.bin_score.device
        if dev.type != 'cuda':
            raise _lib.GimsError('GMatcher must live on a CUDA device (no CPU path): call .to("cuda")')
        key = (self._param_version(), self.config['sinkhorn_iterations'], self.config['match_threshold'])
        if self._model is not None and key == self._packed_key:
            return self._model
        self._invalidate()
        key = (self._param_version(), self.config['sinkhorn_iterations'], self.config['match_threshold'])
        flat, offsets, _ = pack_state_dict(self.state_dict(), self.config)
        self._packed = flat.to(dev)
        cfg = self.c_config()
        off = (C.c_int64 * len(offsets))(*offsets)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gims_model_create(C.byref(cfg), _lib.ptr(self._packed), off, len(offsets),
                                                    C.byref(h)), 'gims_model_create')
        self._model = h
        self._packed_key = key
        return h

    # -- forward ------------------------------------------------------------------------------------
    @staticmethod
    def _k_rank(n, percentile):
        """agc.py:377-378 with the reference's own Python arithmetic."""
        length = n * (n - 1) // 2
        k = int(length * percentile / 100)
        if k >= length:
            k = length - 1
        return k

    def _workspace(self, sizes, edge_cap, dev, slot=0):
        """One GROW-ONLY workspace per (device, stream slot), sized for the largest batch seen so far on that slot
        (rounded up, so that real data with a different keypoint count per pair does not reallocate on every call).
        `sizes`: [(n0, n1), ...] of the batch.  The library carves it for the actual sizes of each call."""
        key = (str(dev), slot)
        n = len(sizes)
        a0 = (C.c_int * n)(*[-(-s[0] // 256) * 256 for s in sizes])
        a1 = (C.c_int * n)(*[-(-s[1] // 256) * 256 for s in sizes])
        need = _lib.lib().gims_batch_workspace_bytes(n, a0, a1, int(edge_cap))
        with _HANDLE_LOCK:
            ws = self._ws.get(key)
            if ws is not None and ws.numel() >= need:
                return ws
            if ws is not None:
                self._ws[key] = None          # release the old block before taking the larger one
                del ws
            ws = torch.zeros(need, dtype=torch.uint8, device=dev)
            self._ws[key] = ws
            return ws

    def _upload(self, t, dev, role):
        """fp32 contiguous copy of an input on `dev`.  Device tensors and pinned host tensors go as they are (an asynchronous
        copy); a PAGEABLE host tensor is first copied into a pinned staging buffer of the calling thread — a pageable
        host->device copy is staged by the driver under the lock the other caller threads need for their launches (see
        `_read_meta`).  One staging buffer per (thread, stream, input role), reused once its last upload has left the host."""
        if t.device.type != 'cpu' or t.is_pinned():
            return t.to(dev, torch.float32, non_blocking=True).contiguous()
        t = t.to(torch.float32).contiguous()
        st = torch.cuda.current_stream(dev)
        key = (str(dev), st.cuda_stream, threading.get_ident(), role)
        ent = self._in_pinned.get(key)
        if ent is not None and ent[1] is not None:
            ent[1].synchronize()                    # the previous upload from this buffer is done
        if ent is None or ent[0].numel() < t.numel():
            ent = [torch.empty(max(t.numel(), 1024), dtype=torch.float32).pin_memory(), None]
            with _HANDLE_LOCK:
                self._in_pinned[key] = ent
        stage = ent[0][:t.numel()].view(t.shape)
        stage.copy_(t)
        out = stage.to(dev, non_blocking=True)
        ent[1] = torch.cuda.Event()
        ent[1].record(st)
        return out

    def _read_meta(self, meta_dev, dev):
        """Device -> host copy of the per-call metadata through a PINNED staging buffer + a stream synchronize.
        `meta_dev.cpu()` is a pageable copy: the driver stages it while holding a lock that the other caller threads need for
        their launches, for as long as this stream still has work — with 8 caller threads that cost 10 % of the throughput
        (tools/prof_e2e.py: 558 -> 618 pairs/s).  One buffer per (device, stream, thread), grow-only."""
        st = torch.cuda.current_stream(dev)
        key = (str(dev), st.cuda_stream, threading.get_ident())
        n = meta_dev.numel()
        buf = self._meta_pinned.get(key)
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 8 + 2 * 4096), dtype=meta_dev.dtype).pin_memory()
            with _HANDLE_LOCK:
                self._meta_pinned[key] = buf
        view = buf[:n]
        view.copy_(meta_dev, non_blocking=True)
        if os.environ.get('GIMS_SPIN_SYNC') == '1':
            st.synchronize()
        else:
            # The forward takes milliseconds: poll an event with short sleeps instead of spinning in the driver.  A spinning
            # caller occupies a core for the whole forward; with more caller threads than cores (8 ranks x 8 callers on the
            # 16-core host of an 8-GPU box) the spinners starve the threads that have launches to issue.
            ev = torch.cuda.Event()
            ev.record(st)
            while not ev.query():
                time.sleep(1e-4)
        return view.clone()

    def _prepare(self, kpts0, desc0, scores0, kpts1, desc1, scores1, shape0, shape1, radius, percentile, min_size,
                 edge_cap, debug, gemm_mode, dev):
        """Output tensors + the C descriptors of one pair."""
        L = _lib.lib()
        n0, n1 = int(kpts0.shape[0]), int(kpts1.shape[0])
        if n0 < 2 or n1 < 2:
            raise ValueError('each image needs at least 2 keypoints (the reference fails earlier, agc.py:439)')
        if max(n0, n1) > _lib.MAX_KPTS:
            raise ValueError('more than %d keypoints per image' % _lib.MAX_KPTS)
        d = self.config['descriptor_dim']
        if edge_cap is None:
            edge_cap = max(1024, int(self.edge_cap_factor) * max(n0, n1))
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        # counts and kept indices share one buffer so that the host needs ONE device->host copy per call
        meta = torch.zeros(8 + n0 + n1, **i32)
        out = {
            'meta': meta,
            'n_kept_dev': meta[:8],                       # [0:2] N', [2:4] E, [4:6] #components, [6] status
            'kept_idx0': meta[8:8 + n0], 'kept_idx1': meta[8 + n0:],
            'thr_dev': torch.zeros(2, **f32),
            'mdesc': torch.empty(n0 + n1, d, **f32),
            'u': torch.empty(n0 + 1, **f32), 'v': torch.empty(n1 + 1, **f32),
        }
        ns = (n0, n1)
        for s in (0, 1):
            out['csr_indptr%d' % s] = torch.empty(ns[s] + 1, **i32)
            out['csr_indices%d' % s] = torch.empty(edge_cap, **i32)
            out['kpts%d' % s] = torch.empty(ns[s], 2, **f32)
            out['feat%d' % s] = torch.empty(ns[s], d, **f32)
            out['scores%d' % s] = torch.empty(ns[s], **f32)
            out['matches%d' % s] = torch.empty(ns[s], dtype=torch.int64, device=dev)
            out['mscores%d' % s] = torch.empty(ns[s], **f32)
            out['indices%d' % s] = torch.empty(ns[s], **i32)
        if debug:
            out['couplings_buf'] = torch.empty(n0 + 1, L.gims_couplings_ld(n1), **f32)      # pitch of the C ABI
            out['couplings'] = out['couplings_buf'][:, :n1 + 1]
            out['desc_gnn'] = torch.empty(n0 + n1, d, **f32)
            out['desc_in'] = torch.empty(n0 + n1, d, **f32)
        pin = _lib.PairInputs()
        ins = ((kpts0, desc0, scores0), (kpts1, desc1, scores1))
        keep = []
        for s in (0, 1):
            k, de, sc = [self._upload(t, dev, (s, i)) for i, t in enumerate(ins[s])]
            keep += [k, de, sc]
            pin.kpts[s], pin.desc[s], pin.scores[s] = k.data_ptr(), de.data_ptr(), sc.data_ptr()
            pin.n[s] = ns[s]
            _, _, height, width = shape0 if s == 0 else shape1        # positional, as gmatcher.py:28
            pin.img_w[s], pin.img_h[s] = float(width), float(height)
            pin.k_rank[s] = self._k_rank(ns[s], percentile)
        pin.desc_channel_major = 1
        pin.radius = float(radius)
        pin.min_size = int(min_size)
        pin.edge_cap = int(edge_cap)
        pin.gemm_mode = 0 if gemm_mode is None else int(gemm_mode) + 1        # 0 = the library default
        po = _lib.PairOutputs()
        cnt = out['n_kept_dev']
        po.n_kept_dev = cnt.data_ptr()
        po.n_edges_dev = cnt.data_ptr() + 8
        po.n_comp_dev = cnt.data_ptr() + 16
        po.status_dev = cnt.data_ptr() + 24
        po.thr_dev = out['thr_dev'].data_ptr()
        for s in (0, 1):
            po.kept_idx[s] = out['kept_idx%d' % s].data_ptr()
            po.csr_indptr[s] = out['csr_indptr%d' % s].data_ptr()
            po.csr_indices[s] = out['csr_indices%d' % s].data_ptr()
            po.kpts[s] = out['kpts%d' % s].data_ptr()
            po.feat[s] = out['feat%d' % s].data_ptr()
            po.scores[s] = out['scores%d' % s].data_ptr()
            po.matches[s] = out['matches%d' % s].data_ptr()
            po.mscores[s] = out['mscores%d' % s].data_ptr()
            po.indices[s] = out['indices%d' % s].data_ptr()
        po.mdesc = out['mdesc'].data_ptr()
        po.u, po.v = out['u'].data_ptr(), out['v'].data_ptr()
        po.couplings = out['couplings_buf'].data_ptr() if debug else None
        po.desc_gnn = out['desc_gnn'].data_ptr() if debug else None
        po.desc_in = out['desc_in'].data_ptr() if debug else None
        out['_inputs'] = keep          # keep the staged inputs alive until the stream has consumed them
        out['edge_cap'] = edge_cap
        return pin, po, out

    def run_pairs(self, items, radius=25, percentile=7, min_size=8, edge_cap=None, debug=False, stream=None, slot=0,
                  gemm_mode=None):
        """A batch of up to `_lib.MAX_BATCH` pairs through ONE `gims_forward_pairs` call (every projection GEMM and
        attention layer is one launch for the whole batch).  `items`: tuples (kpts0 (N,2), desc0 (D,N), scores0 (N,),
        kpts1, desc1, scores1, shape0, shape1) of tensors (host tensors are copied to the device here).  Enqueues
        everything on `stream` (default: current) and returns one dict of device tensors per pair, sized for the INPUT
        counts, plus `n_kept_dev`; nothing synchronises."""
        L = _lib.lib()
        dev = self.bin_score.device          # host inputs are copied here (the H2D of the e2e path)
        if dev.type != 'cuda':
            raise _lib.GimsError('GMatcher must live on a CUDA device (no CPU path): call .to("cuda")')
        n = len(items)
        if not 1 <= n <= _lib.MAX_BATCH:
            raise ValueError('a batch holds 1..%d pairs' % _lib.MAX_BATCH)
        pins = (_lib.PairInputs * n)()
        pos = (_lib.PairOutputs * n)()
        outs = []
        for i, it in enumerate(items):
            pin, po, out = self._prepare(*it, radius, percentile, min_size, edge_cap, debug, gemm_mode, dev)
            pins[i], pos[i] = pin, po
            outs.append(out)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        # one workspace per stream: calls on one stream are ordered, calls on different streams (several batches in
        # flight, or several host threads calling forward() concurrently) must not share scratch memory
        cap = max(o['edge_cap'] for o in outs)
        ws = self._workspace([(int(it[0].shape[0]), int(it[3].shape[0])) for it in items], cap, dev, (slot, st.cuda_stream))
        model = self._acquire()
        try:
            with torch.cuda.device(dev):
                _lib.check(L.gims_forward_pairs(model, n, pins, pos, _lib.ptr(ws), ws.numel(),
                                                C.c_void_p(st.cuda_stream)), 'gims_forward_pairs')
        finally:
            self._release(model)
        return outs

    def run_pair(self, kpts0, desc0, scores0, kpts1, desc1, scores1, shape0, shape1, radius=25, percentile=7,
                 min_size=8, edge_cap=None, debug=False, stream=None, slot=0, gemm_mode=None):
        """One pair on the device (see `run_pairs`).  kpts (N,2), desc (D,N) channel-major, scores (N,) fp32 tensors."""
        return self.run_pairs([(kpts0, desc0, scores0, kpts1, desc1, scores1, shape0, shape1)], radius, percentile,
                              min_size, edge_cap, debug, stream, slot, gemm_mode)[0]

    def forward(self, data, **kwargs):
        if kwargs.get('mode', 'test') == 'train':
            raise NotImplementedError('forward_train (gmatcher.py:309-386) is outside the B200 hot path')
        if data.get('delaunay', False):
            raise NotImplementedError('the Delaunay graph branch is broken in the reference snapshot '
                                      '(gmatcher.py:223-231 never sets kept_kpts*_indices)')
        radius = data.get('radius', 25)
        percentile = data.get('percentile', 7)
        min_size = data.get('min_size', 8)
        dev = self.bin_score.device
        batch = data['keypoints0'].shape[0]
        per_item = []
        for b in range(batch):
            cap, mode = None, None
            while True:
                r = self.run_pair(data['keypoints0'][b], data['descriptors0'][b], data['scores0'][b],
                                  data['keypoints1'][b], data['descriptors1'][b], data['scores1'][b],
                                  data['image0'].shape, data['image1'].shape, radius, percentile, min_size,
                                  edge_cap=cap, gemm_mode=mode)
                meta = self._read_meta(r['meta'], dev)  # the one device->host sync of the call: counts + kept indices
                counts = meta[:8]
                status = int(counts[6])
                if status & _lib.STATUS_EDGE_OVERFLOW:
                    # the graph of this attempt was reported empty (nothing downstream ran on it): retry with more room
                    n_max = max(r['kpts0'].shape[0], r['kpts1'].shape[0])
                    if r['edge_cap'] >= n_max * n_max:
                        raise _lib.GimsError('edge capacity overflow')
                    cap = min(r['edge_cap'] * 4, n_max * n_max)
                    continue
                if status & _lib.STATUS_FP16_RANGE:
                    # an attention operand left the fp16 range (|x| >= 32768): redo this pair on the 3xTF32 kernels
                    if mode == _lib.GEMM_TC:
                        raise _lib.GimsError('fp16 range flag raised by the tf32 path')
                    mode = _lib.GEMM_TC
                    continue
                if status & _lib.STATUS_SINKHORN_TIMEOUT:
                    raise _lib.GimsError('Sinkhorn kernel: a grid-wide wait timed out (GPU shared / preempted?); '
                                         'the results of this call are invalid')
                break
            r['counts'] = counts
            r['meta_host'] = meta
            per_item.append(r)
        def stack(lst):                         # batch 1 (the usual call): a view instead of a copy kernel
            return lst[0].unsqueeze(0) if len(lst) == 1 else torch.stack(lst)

        res = {s: [] for s in ('k0', 'k1', 'd0', 'd1', 's0', 's1', 'm0', 'm1', 'ms0', 'ms1', 'md0', 'md1')}
        kept0, kept1, g0, g1 = [], [], [], []
        for r in per_item:
            a, c = int(r['counts'][0]), int(r['counts'][1])
            e0, e1 = int(r['counts'][2]), int(r['counts'][3])
            n0_in = r['kpts0'].shape[0]
            res['k0'].append(r['kpts0'][:a]); res['k1'].append(r['kpts1'][:c])
            res['d0'].append(r['feat0'][:a]); res['d1'].append(r['feat1'][:c])
            res['s0'].append(r['scores0'][:a]); res['s1'].append(r['scores1'][:c])
            res['m0'].append(r['matches0'][:a]); res['m1'].append(r['matches1'][:c])
            res['ms0'].append(r['mscores0'][:a]); res['ms1'].append(r['mscores1'][:c])
            res['md0'].append(r['mdesc'][:a]); res['md1'].append(r['mdesc'][n0_in:n0_in + c])
            kept0.append(r['meta_host'][8:8 + a].tolist())
            kept1.append(r['meta_host'][8 + n0_in:8 + n0_in + c].tolist())
            g0.append((r['csr_indptr0'][:a + 1], r['csr_indices0'][:e0]))
            g1.append((r['csr_indptr1'][:c + 1], r['csr_indices1'][:e1]))
        # same side effects on the caller's dict as gmatcher.py:244-252 (graphs are CSR pairs, not DGL)
        data['keypoints0'] = stack(res['k0'])
        data['descriptors0'] = stack(res['d0']).permute(0, 2, 1)
        data['keypoints1'] = stack(res['k1'])
        data['descriptors1'] = stack(res['d1']).permute(0, 2, 1)
        data['scores0'] = stack(res['s0'])
        data['scores1'] = stack(res['s1'])
        data['kept_kpts0_indices'] = kept0
        data['kept_kpts1_indices'] = kept1
        data['graph0'], data['graph1'] = g0, g1
        kpts0, kpts1 = data['keypoints0'], data['keypoints1']
        if kpts0.shape[1] == 0 or kpts1.shape[1] == 0:      # gmatcher.py:257-264
            shape0, shape1 = kpts0.shape[:-1], kpts1.shape[:-1]
            return {
                'matches0': kpts0.new_full(shape0, -1, dtype=torch.int),
                'matches1': kpts1.new_full(shape1, -1, dtype=torch.int),
                'matching_scores0': kpts0.new_zeros(shape0),
                'matching_scores1': kpts1.new_zeros(shape1),
            }
        mdesc0, mdesc1 = stack(res['md0']), stack(res['md1'])
        return {
            'keypoints0': data['keypoints0'],
            'keypoints1': data['keypoints1'],
            'descriptors0': data['descriptors0'],
            'descriptors1': data['descriptors1'],
            'matches0': stack(res['m0']),
            'matches1': stack(res['m1']),
            'matching_scores0': stack(res['ms0']),
            'matching_scores1': stack(res['ms1']),
            'mdesc0': mdesc0.squeeze(),
            'mdesc1': mdesc1.squeeze(),
        }

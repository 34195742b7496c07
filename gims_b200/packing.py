"""Pack a reference `GMatcher.state_dict()` into the flat fp32 buffer the CUDA library consumes.

Done once at load time, in float64 then rounded to fp32:
  * eval-mode BatchNorm1d folded into the preceding Conv1d(k=1)  (models/gmatcher.py:11-24):
        s = gamma / sqrt(running_var + eps);  W' = s[:,None] * W;  b' = (b - running_mean) * s + beta
  * attention heads de-interleaved: the reference views channels as (dim=64, heads=4), i.e. channel
    c = d*4 + h (gmatcher.py:108-111); packed channel c' = h*64 + d, applied to the output rows of
    proj[0..2] and to the input columns of `merge` (which is then composed into the first MLP conv)
  * proj[0] | proj[1] | proj[2] stacked into one [768][256] matrix (Q | K | V)
  * SAGEConv layer 0 (in > out: fc_neigh before aggregation) stacked as rows [fc_neigh; fc_self];
    layers 1, 2 stacked along K as [fc_self | fc_neigh] acting on cat[h, mean_neigh(h)]
  * Conv1d weights (out, in, 1) -> (out, in)

  * every weight matrix W is followed by its tensor-core planes W_hi = tf32(W) (round-to-nearest, ties away,
    like cvt.rna.tf32.f32) and W_lo = W - W_hi (exact in fp32), consumed by the 3xTF32 tcgen05 GEMM

Blob order (== offsets passed to gims_model_create); each matrix W stands for six blobs W, W_hi, W_lo (tf32 planes),
W_h16, W_l16 (fp16 planes of W * 2^e, raw bits), 2^-e:
  bin_score, (kenc W_i, b_i) for every kenc conv, (sage W_l, b_l) l=0..2,
  per attention layer: Wqkv, bqkv, W1 (merge composed in, BN folded), b1, W2, b2; final_proj W, b.
"""
import torch

from .config import BN_EPS, DEFAULT_CONFIG, NUM_HEADS, kenc_channels


def _fold_bn(w, b, sd, bn_prefix):
    gamma = sd[bn_prefix + '.weight'].double()
    beta = sd[bn_prefix + '.bias'].double()
    mean = sd[bn_prefix + '.running_mean'].double()
    var = sd[bn_prefix + '.running_var'].double()
    s = gamma / torch.sqrt(var + BN_EPS)
    return w * s[:, None], (b - mean) * s + beta


def head_permutation(d):
    """perm[c'] = original channel of packed channel c' = h*hd + dd."""
    hd = d // NUM_HEADS
    cp = torch.arange(d)
    return (cp % hd) * NUM_HEADS + cp // hd


def split_tf32(w32):
    """fp32 tensor -> (hi, lo): hi = w rounded to a 10-bit mantissa (ties away from zero), lo = w - hi."""
    bits = w32.contiguous().view(torch.int32)
    hi_bits = (bits + 0x1000) & ~0x1FFF          # sign-magnitude encoding: adding half an ulp rounds |w| away on ties
    hi = hi_bits.view(torch.float32)
    return hi, w32 - hi


def split_f16(w32):
    """fp32 matrix -> (hi bits, lo bits, 2^-e): hi = fp16(W * 2^e), lo = fp16(W * 2^e - hi) (round to nearest even, what
    cvt.rn.f16.f32 computes), the two planes returned as float32 tensors that hold the raw 16-bit patterns (two per
    word, row-major); e puts max|W| * 2^e into [512, 1024], so that both planes are normal fp16 numbers for every weight
    within 2^-13 of the largest one."""
    w32 = w32.contiguous()
    amax = float(w32.abs().max()) if w32.numel() else 0.0
    e = int(torch.floor(torch.log2(torch.tensor(1024.0 / amax)))) if amax > 0 else 0
    e = max(-12, min(24, e))
    scaled = w32 * float(2.0 ** e)
    hi = scaled.to(torch.float16)
    lo = (scaled - hi.float()).to(torch.float16)
    assert w32.numel() % 2 == 0
    return hi.view(torch.float32).reshape(-1), lo.view(torch.float32).reshape(-1), torch.tensor([2.0 ** -e], dtype=torch.float32)


def pack_state_dict(sd, config=None):
    """Returns (flat fp32 CPU tensor, list of float offsets, list of blob names)."""
    cfg = {**DEFAULT_CONFIG, **(config or {})}
    d = cfg['descriptor_dim']
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    blobs = []

    def add(name, t):
        t = t.double().contiguous()
        blobs.append((name, t))
        if t.dim() == 2:                          # weight matrix: append the tensor-core planes
            hi, lo = split_tf32(t.float())
            blobs.append((name + '.hi', hi.double()))
            blobs.append((name + '.lo', lo.double()))
            h16, l16, sinv = split_f16(t.float())
            blobs.append((name + '.h16', h16))    # raw 16-bit patterns: float32 words, never converted
            blobs.append((name + '.l16', l16))
            blobs.append((name + '.sinv', sinv))

    add('bin_score', sd['bin_score'].reshape(1))
    ch = kenc_channels(cfg)
    n_conv = len(ch) - 1
    for i in range(n_conv):
        w = sd['kenc.encoder.%d.weight' % (3 * i)].double().squeeze(-1)
        b = sd['kenc.encoder.%d.bias' % (3 * i)].double()
        if i < n_conv - 1:
            w, b = _fold_bn(w, b, sd, 'kenc.encoder.%d' % (3 * i + 1))
        add('kenc.w%d' % i, w)
        add('kenc.b%d' % i, b)
    for l in range(3):
        p = 'gnn_encoder.layers.%d' % l
        wn, ws = sd[p + '.fc_neigh.weight'].double(), sd[p + '.fc_self.weight'].double()
        bias = sd[p + '.bias'] if (p + '.bias') in sd else sd[p + '.fc_self.bias']
        if wn.shape[1] > wn.shape[0]:
            w = torch.cat([wn, ws], 0)          # rows: [fc_neigh; fc_self]
        else:
            w = torch.cat([ws, wn], 1)          # cols: [fc_self | fc_neigh]
        add('sage.w%d' % l, w)
        add('sage.b%d' % l, bias.double())
    perm = head_permutation(d)
    for l in range(len(cfg['transformer_layers'])):
        p = 'gnn.layers.%d' % l
        wq = [sd['%s.attn.proj.%d.weight' % (p, j)].double().squeeze(-1)[perm] for j in range(3)]
        bq = [sd['%s.attn.proj.%d.bias' % (p, j)].double()[perm] for j in range(3)]
        add('l%d.wqkv' % l, torch.cat(wq, 0))
        add('l%d.bqkv' % l, torch.cat(bq, 0))
        # `merge` (Conv1d on the attention output, gmatcher.py:114) is linear and feeds only the first MLP conv
        # (gmatcher.py:125): it is composed into that conv here, in fp64 —
        #   W1 [x | merge(att)] + b1 = [W1x | W1m Wm] [x | att] + (b1 + W1m bm)
        wm = sd[p + '.attn.merge.weight'].double().squeeze(-1)[:, perm]
        bm = sd[p + '.attn.merge.bias'].double()
        w1_raw, b1_raw = sd[p + '.mlp.0.weight'].double().squeeze(-1), sd[p + '.mlp.0.bias'].double()
        w1_raw = torch.cat([w1_raw[:, :d], w1_raw[:, d:] @ wm], 1)
        b1_raw = b1_raw + sd[p + '.mlp.0.weight'].double().squeeze(-1)[:, d:] @ bm
        w1, b1 = _fold_bn(w1_raw, b1_raw, sd, p + '.mlp.1')
        add('l%d.w1' % l, w1)
        add('l%d.b1' % l, b1)
        add('l%d.w2' % l, sd[p + '.mlp.3.weight'].double().squeeze(-1))
        add('l%d.b2' % l, sd[p + '.mlp.3.bias'].double())
    add('final.w', sd['final_proj.weight'].double().squeeze(-1))
    add('final.b', sd['final_proj.bias'].double())

    offsets, names, parts, off = [], [], [], 0
    for name, t in blobs:
        flat = t.reshape(-1) if t.dtype == torch.float32 else t.reshape(-1).float()
        pad = (-flat.numel()) % 64              # keep every blob 256-byte aligned
        offsets.append(off)
        names.append(name)
        parts.append(flat)
        if pad:
            parts.append(torch.zeros(pad))
        off += flat.numel() + pad
    return torch.cat(parts).contiguous(), offsets, names

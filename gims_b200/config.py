"""Configuration and state_dict schema of the GMatcher forward path.

Mirrors the reference's `GMatcher.default_config` (models/gmatcher.py:166-176) and the parameter
names/shapes its modules register (models/gmatcher.py:11-24, 87-162, 177-207), so that a reference
checkpoint loads unchanged.
"""
from collections import OrderedDict

DEFAULT_CONFIG = {
    'descriptor_dim': 256,
    'weights_path': None,
    'keypoint_encoder': [32, 64, 128, 256],
    'transformer_layers': ['self', 'cross'] * 9,
    'sinkhorn_iterations': 100,
    'match_threshold': 0.2,
    'use_layernorm': False,
    'input_dim': 256,
    'num_heads': 4,          # ignored by the reference too: 4 is hard-coded (gmatcher.py:131)
}

NUM_HEADS = 4                # models/gmatcher.py:131
BN_EPS = 1e-5                # torch.nn.BatchNorm1d default
SAGE_LAYERS = 3              # models/gmatcher.py:192-197


def kenc_channels(config):
    """Channel chain of the keypoint encoder MLP (gmatcher.py:91): [2] + layers + [D]."""
    return [2] + list(config['keypoint_encoder']) + [config['descriptor_dim']]


def sage_dims(config):
    """(in, out) of the three SAGEConv layers (gmatcher.py:192-197, 148-151)."""
    d = config['descriptor_dim']
    h = int(d / 2)
    return [(d, h), (h, h), (h, d)]


def state_dict_schema(config=None):
    """Ordered {key: shape} of every tensor in the reference `GMatcher.state_dict()`.

    SAGE bias is listed under the DGL-1.x spelling `gnn_encoder.layers.i.bias`; loaders also accept
    `gnn_encoder.layers.i.fc_self.bias`.
    """
    cfg = {**DEFAULT_CONFIG, **(config or {})}
    if cfg['use_layernorm']:
        raise NotImplementedError('use_layernorm=True is not on the hot path (reference default False)')
    d = cfg['descriptor_dim']
    sch = OrderedDict()

    def conv(prefix, cin, cout):
        sch[prefix + '.weight'] = (cout, cin, 1)
        sch[prefix + '.bias'] = (cout,)

    def bn(prefix, c):
        sch[prefix + '.weight'] = (c,)
        sch[prefix + '.bias'] = (c,)
        sch[prefix + '.running_mean'] = (c,)
        sch[prefix + '.running_var'] = (c,)
        sch[prefix + '.num_batches_tracked'] = ()

    sch['bin_score'] = ()
    ch = kenc_channels(cfg)
    for i in range(1, len(ch)):
        conv('kenc.encoder.%d' % (3 * (i - 1)), ch[i - 1], ch[i])
        if i < len(ch) - 1:
            bn('kenc.encoder.%d' % (3 * (i - 1) + 1), ch[i])
    for l in range(len(cfg['transformer_layers'])):
        p = 'gnn.layers.%d' % l
        conv(p + '.attn.merge', d, d)
        for j in range(3):
            conv(p + '.attn.proj.%d' % j, d, d)
        conv(p + '.mlp.0', 2 * d, 2 * d)
        bn(p + '.mlp.1', 2 * d)
        conv(p + '.mlp.3', 2 * d, d)
    for i, (cin, cout) in enumerate(sage_dims(cfg)):
        p = 'gnn_encoder.layers.%d' % i
        sch[p + '.bias'] = (cout,)
        sch[p + '.fc_neigh.weight'] = (cout, cin)
        sch[p + '.fc_self.weight'] = (cout, cin)
    if cfg['input_dim'] != d:   # constructed but never used in forward (gmatcher.py:198-201)
        sch['input_proj.weight'] = (d, cfg['input_dim'])
        sch['input_proj.bias'] = (d,)
    conv('final_proj', d, d)
    return sch

"""Helpers shared by the parity tests: golden fixtures (tests/golden/*.npz, written by
oracle/make_golden.py from the unmodified reference) and their synthetic-input recipes."""
import ast
import glob
import os

import numpy as np
import torch

from gims_b200.synth import make_pair, make_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + '*.npz')))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    rec = ast.literal_eval(str(g['recipe']))
    return rec, g


def inputs_for(rec):
    data = make_pair(rec['n0'], rec['n1'], seed=rec['seed'], width=rec['width'], height=rec['height'],
                     image_style=rec.get('image_style', 'tensor'))
    data.update({'radius': rec['radius'], 'percentile': rec['percentile'], 'min_size': rec['min_size']})
    return data


def weights_for(rec):
    return make_state_dict(rec['wseed'], peaked=rec['peaked'], damped=rec['damped'])


def index_agreement(ours, ref_idx, ref_best, ref_at_ours, tol=1e-4):
    """Tie-aware agreement: a disagreement counts only if the reference's Z at our
    index is more than `tol` below its maximum."""
    ours = np.asarray(ours)
    bad = (ours != ref_idx) & ((ref_best - ref_at_ours) > tol)
    return 1.0 - bad.mean() if len(ours) else 1.0

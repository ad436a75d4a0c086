"""Shared test helpers: golden loading and seeded case reconstruction."""
import json
import os

import numpy as np

from n2nmn_b200 import synth, weights as wts

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(family):
    z = np.load(os.path.join(GOLDEN, 'golden_%s.npz' % family))
    meta = json.loads(str(z['meta']))
    return z, meta


def case_inputs(meta):
    """Regenerate the exact inputs/weights the golden generator used."""
    feat, word_vecs = synth.make_inputs(meta['N'], meta['H'], meta['W'], meta['D'], meta['T'],
                                        seed=meta['seed_in'])
    W = wts.init_weights(meta['family'], meta['H'], meta['W'], meta['D'], meta['C'],
                         seed=meta['seed_w'], bias_std=0.1)
    return feat, word_vecs, W


def node_inputs(meta, n, seed):
    rng = np.random.RandomState(seed)
    t = rng.randint(0, meta['T'], size=n).astype(np.int32)
    b = rng.randint(0, meta['N'], size=n).astype(np.int32)
    a0 = (2.0 * rng.standard_normal((n, meta['H'], meta['W'], 1))).astype(np.float32)
    a1 = (2.0 * rng.standard_normal((n, meta['H'], meta['W'], 1))).astype(np.float32)
    return t, b, a0, a1

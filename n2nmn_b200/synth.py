"""Seeded synthetic inputs and layouts for tests and bench.

No dataset is reachable offline, so every workload is generated: post-ReLU-like feature grids,
small-variance text vectors, and Reverse-Polish layouts drawn either from a fixed expert-like
mix, from the real SHAPES / VQA layout histograms, or by running the
reference's decoder validity automaton (Assembler.P/W/b) with uniformly random admissible tokens.
All generators use ``numpy.random.RandomState`` so streams are stable across numpy versions.
"""
from __future__ import annotations

import os

import numpy as np

from .assembler import Assembler, MODULE_SIGNATURE

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def vocab_file(family):
    return os.path.join(_DATA, 'vocabulary_layout_%s.txt' % family)


def make_inputs(N, H, W, D, T, seed=1234, text_dim=300):
    """image_feat_grid [N,H,W,D] = relu(N(0,1)) (pool5 / res5c are post-ReLU);
    word_vecs [T,N,text_dim] = N(0, 0.3^2)."""
    rng = np.random.RandomState(seed)
    feat = np.maximum(rng.standard_normal((N, H, W, D)), 0).astype(np.float32)
    word_vecs = (0.3 * rng.standard_normal((T, N, text_dim))).astype(np.float32)
    return feat, word_vecs


# Expert-like CLEVR layouts (every one assembles valid).
CLEVR_EXPERT_MIX = [
    ['_Find', '_Count'],
    ['_Scene', '_Count'],
    ['_Find', '_Transform', '_Filter', '_Count'],
    ['_Find', '_Transform', '_Filter', '_Describe'],
    ['_Find', '_Find', '_And', '_Exist'],
    ['_Find', '_Filter', '_Find', '_Filter', '_EqualNum'],
    ['_Find', '_FindSameProperty', '_Filter', '_Count'],
    ['_Find', '_Transform', '_Find', '_Transform', '_Or', '_Count'],
    ['_Find', '_Find', '_SameProperty'],
    ['_Find', '_Transform', '_Filter', '_Transform', '_Filter', '_Describe'],
]

# The only three layouts in the real SHAPES data, with their test-split counts (App. C).
SHAPES_LAYOUTS = [
    (['_Find', '_Find', '_Transform', '_And', '_Answer'], 640),
    (['_Find', '_Find', '_Transform', '_Transform', '_And', '_Answer'], 256),
    (['_Find', '_Find', '_And', '_Answer'], 128),
]

# Head of the VQA gt-layout histogram (App. C: 46.5 / 42.4 / 8.8 %; the rest are longer chains).
VQA_LAYOUTS = [
    (['_Find', '_Find', '_And', '_Describe'], 465),
    (['_Find', '_Describe'], 424),
    (['_Find', '_Transform', '_Describe'], 88),
    (['_Find', '_Transform', '_Find', '_And', '_Describe'], 14),
    (['_Find', '_Find', '_And', '_Find', '_And', '_Describe'], 9),
]


def tokens_from_layouts(assembler, layouts, T):
    """List of module-name lists -> int32 [T, N] matrix, <eos>-padded."""
    cols = [assembler.module_list2tokens(l, T) for l in layouts]
    return np.ascontiguousarray(np.array(cols, np.int32).T)


def expert_mix_tokens(assembler, N, T):
    layouts = [CLEVR_EXPERT_MIX[i % len(CLEVR_EXPERT_MIX)] for i in range(N)]
    return tokens_from_layouts(assembler, layouts, T)


def histogram_tokens(assembler, weighted_layouts, N, T, seed=7):
    rng = np.random.RandomState(seed)
    w = np.array([c for _, c in weighted_layouts], np.float64)
    pick = rng.choice(len(weighted_layouts), size=N, p=w / w.sum())
    return tokens_from_layouts(assembler, [weighted_layouts[i][0] for i in pick], T)


def layout_depth(assembler, tokens_col):
    """Tree depth of one RPN column (Find/Scene = 1); 0 if it does not parse."""
    stack = []
    for tok in tokens_col:
        if tok == assembler.EOS_idx:
            break
        arity = MODULE_SIGNATURE[assembler.module_names[tok]][0]
        if len(stack) < arity:
            return 0
        d = 1 + max([stack.pop() for _ in range(arity)], default=0)
        stack.append(d)
    return stack[0] if len(stack) == 1 else 0


def random_valid_tokens(assembler, N, T, seed=7, ans_weight=1.0, min_depth=1, max_depth=None,
                        max_tries=2000):
    """Sample layouts by walking the decoder validity automaton with random admissible tokens
    (the masking rule of models_clevr/nmn3_netgen_att.py:8-15 applied to Assembler.P/W/b).
    ``ans_weight`` < 1 down-weights answer tokens to get deeper trees; layouts whose depth falls
    outside [min_depth, max_depth] are rejected and redrawn."""
    rng = np.random.RandomState(seed)
    P, Wm, b = (assembler.P.astype(np.int64), assembler.W.astype(np.int64),
                assembler.b.astype(np.int64))
    V = assembler.num_vocab_nmn
    is_ans = np.array([n != '<eos>' and MODULE_SIGNATURE[n][1] == 'ans'
                       for n in assembler.module_names])
    out = np.full((T, N), assembler.EOS_idx, np.int32)
    for n in range(N):
        for _ in range(max_tries):
            x = np.array([0, 0, T], np.int64)
            col = []
            for _t in range(T):
                ok = np.all(np.einsum('k,kvc->vc', x, Wm) - b >= 0, axis=1)
                p = ok.astype(np.float64) * np.where(is_ans, ans_weight, 1.0)
                tok = int(rng.choice(V, p=p / p.sum()))
                col.append(tok)
                x = x + P[tok]
            d = layout_depth(assembler, col)
            if d >= min_depth and (max_depth is None or d <= max_depth):
                out[:, n] = col
                break
        else:
            raise RuntimeError('could not sample a layout within the depth bounds')
    return out


def default_assembler(family):
    return Assembler(vocab_file(family))

#!/usr/bin/env python
"""Generate tests/golden/*.npz|json by EXECUTING the reference's own files from /root/reference.

Run here (authoring container) only:  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read the committed fixtures, never this script's imports.

What runs unmodified from the reference:
  * models_{clevr,shapes,vqa}/nmn3_assembler.py  (pure numpy; TF imports are unused)
  * models_{clevr,shapes,vqa}/nmn3_modules.py + util/cnn.py + util/empty_safe_conv.py, executed
    eagerly on oracle/tf1_shim.py (numpy stand-in for the TF ops they call).
The tree recursion (TF Fold is unavailable) is a plain post-order walk calling the reference's
module methods with n=1 — the evaluation order of exp_shapes/visualize_shapes.ipynb cell 9.

Fixtures store seeds + outputs; inputs are regenerated from the seeds through
n2nmn_b200.synth / n2nmn_b200.weights (numpy RandomState streams, version-stable).
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'
if not hasattr(np, 'bool'):     # the reference uses the removed alias np.bool
    np.bool = bool

from n2nmn_b200 import synth, weights as wts  # noqa: E402
from oracle import tf1_shim  # noqa: E402

CASES = {
    # family: N images, grid, feature channels handed to Modules, T, C, seeds
    'clevr': dict(N=6, H=10, W=15, D=512, T=12, C=28, seed_in=101, seed_w=11),
    'shapes': dict(N=5, H=3, W=3, D=64, T=8, C=2, seed_in=102, seed_w=12),
    'vqa': dict(N=3, H=14, W=14, D=38, T=8, C=37, seed_in=103, seed_w=13),
}


def ref_import(name):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return importlib.import_module(name)


def store_for(family, W):
    prefix = '' if family == 'shapes' else 'module_variables/'
    return {prefix + k: v for k, v in W.items()}


def node_inputs(c, n, seed):
    rng = np.random.RandomState(seed)
    t = rng.randint(0, c['T'], size=n).astype(np.int32)
    b = rng.randint(0, c['N'], size=n).astype(np.int32)
    a0 = (2.0 * rng.standard_normal((n, c['H'], c['W'], 1))).astype(np.float32)
    a1 = (2.0 * rng.standard_normal((n, c['H'], c['W'], 1))).astype(np.float32)
    return t, b, a0, a1


MODULE_CALLS = {
    'clevr': [('SceneModule', 0), ('FindModule', 0), ('FilterModule', 1),
              ('FindSamePropertyModule', 1), ('TransformModule', 1), ('AndModule', 2),
              ('OrModule', 2), ('ExistModule', 1), ('CountModule', 1), ('EqualNumModule', 2),
              ('MoreNumModule', 2), ('LessNumModule', 2), ('SamePropertyModule', 2),
              ('DescribeModule', 1)],
    'shapes': [('FindModule', 0), ('TransformModule', 1), ('AndModule', 2), ('AnswerModule', 1)],
    'vqa': [('FindModule', 0), ('TransformModule', 1), ('AndModule', 2), ('DescribeModule', 1)],
}

TOKEN_METHOD = {
    '_Scene': 'SceneModule', '_Find': 'FindModule', '_Filter': 'FilterModule',
    '_FindSameProperty': 'FindSamePropertyModule', '_Transform': 'TransformModule',
    '_And': 'AndModule', '_Or': 'OrModule', '_Exist': 'ExistModule', '_Count': 'CountModule',
    '_EqualNum': 'EqualNumModule', '_MoreNum': 'MoreNumModule', '_LessNum': 'LessNumModule',
    '_SameProperty': 'SamePropertyModule', '_Describe': 'DescribeModule',
    '_Answer': 'AnswerModule'}


def build_ref_modules(family, c):
    feat, word_vecs = synth.make_inputs(c['N'], c['H'], c['W'], c['D'], c['T'], seed=c['seed_in'])
    W = wts.init_weights(family, c['H'], c['W'], c['D'], c['C'], seed=c['seed_w'], bias_std=0.1)
    tf1_shim.uninstall()
    tf1_shim.install(store_for(family, W))
    mod = ref_import('models_%s.nmn3_modules' % family)
    feat_t = tf1_shim._t(feat, static=[None, c['H'], c['W'], c['D']])
    wv_t = tf1_shim._t(word_vecs)
    if family == 'vqa':
        m = mod.Modules(feat_t, wv_t, None, c['C'])
    else:
        m = mod.Modules(feat_t, wv_t, c['C'])
    return m


def call_module(family, m, name, ins, t, b):
    kw = {}
    if family == 'shapes':
        kw = dict(reuse=None)
    return np.asarray(getattr(m, name)(*ins, t, b, **kw))


def gen_modules(family):
    c = CASES[family]
    m = build_ref_modules(family, c)
    out = {}
    for k, (name, arity) in enumerate(MODULE_CALLS[family]):
        n = 5
        t, b, a0, a1 = node_inputs(c, n, seed=1000 + k)
        ins = [tf1_shim._t(a) for a in (a0, a1)[:arity]]
        out[name] = call_module(family, m, name, ins, t, b).astype(np.float32)
    # every variable the reference created must be one we supplied (names line up with App. B)
    created = sorted(set(tf1_shim.STATE.created))
    return out, created


def layouts_for(family, asm):
    T = CASES[family]['T']
    N = CASES[family]['N']
    if family == 'clevr':
        cols = [synth.CLEVR_EXPERT_MIX[i] for i in (2, 5, 6, 7, 8, 9)]
    elif family == 'shapes':
        cols = [l for l, _ in synth.SHAPES_LAYOUTS] + [['_Find', '_Answer'],
                                                       ['_Find', '_Transform', '_Answer']]
    else:
        cols = [l for l, _ in synth.VQA_LAYOUTS][:N]
    toks = np.array([asm.module_list2tokens(l, T) for l in cols[:N]], np.int32).T
    return np.ascontiguousarray(toks)


def gen_executor(family):
    c = CASES[family]
    asm_mod = ref_import('models_%s.nmn3_assembler' % family)
    asm = asm_mod.Assembler(os.path.join(REF, 'exp_%s/data/vocabulary_layout.txt' % family))
    tokens = layouts_for(family, asm)
    if family == 'clevr':   # make question 3 invalid (att-typed root): zero scores row
        tokens[:, 3] = asm.module_list2tokens(['_Find', '_Transform'], c['T'])
    expr_list, validity = asm.assemble(tokens)
    m = build_ref_modules(family, c)
    att_maps = {}

    def ev(e):
        ins = [ev(e[k]) for k in ('input_0', 'input_1') if k in e]
        t = np.array([e['time_idx']], np.int32)
        b = np.array([e['batch_idx']], np.int32)
        out = call_module(family, m, TOKEN_METHOD[e['module']], ins, t, b)
        if e['output_type'] == 'att':
            att_maps['att_b%d_t%d' % (e['batch_idx'], e['time_idx'])] = \
                np.asarray(out)[0, :, :, 0].astype(np.float32)
        return tf1_shim._t(out)

    scores = np.zeros((len(expr_list), c['C']), np.float32)
    for i, e in enumerate(expr_list):
        if e['module'] != asm_mod.INVALID_EXPR:
            scores[i] = np.asarray(ev(e))[0]
    return tokens, validity, scores, att_maps


def gen_assembler():
    """Known-answer cases for Assembler.assemble / module_list2tokens / P,W,b from the reference's
    own assembler files (no TF needed: the imports are unused)."""
    tf1_shim.uninstall()
    tf1_shim.install({})
    out = {}
    rng = np.random.RandomState(5)
    for family in ('clevr', 'shapes', 'vqa'):
        asm_mod = ref_import('models_%s.nmn3_assembler' % family)
        asm = asm_mod.Assembler(os.path.join(REF, 'exp_%s/data/vocabulary_layout.txt' % family))
        T = 9
        V = len(asm.module_names)
        # random token matrices (mostly invalid) + structured valid ones
        tokens = rng.randint(0, V, size=(T, 40)).astype(np.int32)
        tokens[rng.randint(2, T, size=40), np.arange(40)] = asm.EOS_idx
        good = {'clevr': synth.CLEVR_EXPERT_MIX,
                'shapes': [l for l, _ in synth.SHAPES_LAYOUTS],
                'vqa': [l for l, _ in synth.VQA_LAYOUTS]}[family]
        good_cols = np.array([asm.module_list2tokens(l, T) for l in good], np.int32).T
        no_eos = np.full((T, 1), asm.name2idx_dict['_Find'], np.int32)
        tokens = np.concatenate([tokens, good_cols, no_eos], axis=1)
        expr_list, validity = asm.assemble(tokens)
        entry = {'module_names': asm.module_names, 'EOS_idx': int(asm.EOS_idx),
                 'tokens': tokens.tolist(), 'validity': [bool(v) for v in validity],
                 'expr_list': expr_list}
        if hasattr(asm, 'P'):
            entry.update(P=asm.P.tolist(), W=asm.W.tolist(), b=asm.b.tolist())
        out[family] = entry
    return out


def main():
    for family in ('clevr', 'shapes', 'vqa'):
        mods, created = gen_modules(family)
        tokens, validity, scores, att_maps = gen_executor(family)
        meta = dict(CASES[family], family=family, created_variables=created,
                    node_seed_base=1000, module_calls=MODULE_CALLS[family])
        arrays = {'mod_' + k: v for k, v in mods.items()}
        arrays.update(att_maps)
        np.savez_compressed(os.path.join(HERE, 'golden_%s.npz' % family),
                            meta=json.dumps(meta), exec_tokens=tokens,
                            exec_validity=validity.astype(np.uint8), exec_scores=scores, **arrays)
        print(family, 'modules:', {k: v.shape for k, v in mods.items()})
        print(family, 'exec scores', scores.shape, 'valid', validity.tolist(),
              'att maps', len(att_maps))
    with open(os.path.join(HERE, 'golden_assembler.json'), 'w') as f:
        json.dump(gen_assembler(), f)
    print('assembler goldens written')


if __name__ == '__main__':
    main()

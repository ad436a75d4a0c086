"""CPU: Assembler drop-in vs goldens made by the reference's own nmn3_assembler.py files."""
import json
import os

import numpy as np
import pytest

from n2nmn_b200 import synth
from n2nmn_b200.assembler import Assembler, INVALID_EXPR
from tests.helpers import GOLDEN

with open(os.path.join(GOLDEN, 'golden_assembler.json')) as f:
    G = json.load(f)


@pytest.mark.parametrize('family', ['clevr', 'shapes', 'vqa'])
def test_assemble_matches_reference(family):
    g = G[family]
    asm = Assembler(synth.vocab_file(family))
    assert asm.module_names == g['module_names']
    assert asm.EOS_idx == g['EOS_idx']
    assert asm.num_vocab_nmn == len(g['module_names'])
    tokens = np.array(g['tokens'], np.int32)
    expr_list, validity = asm.assemble(tokens)
    assert validity.dtype == bool and validity.tolist() == g['validity']
    assert expr_list == g['expr_list']          # same nested dicts, same error strings
    if 'P' in g:
        np.testing.assert_array_equal(asm.P, np.array(g['P']))
        np.testing.assert_array_equal(asm.W, np.array(g['W']))
        np.testing.assert_array_equal(asm.b, np.array(g['b']))
        assert asm.P.dtype == asm.W.dtype == asm.b.dtype == np.int32


def test_module_list2tokens_and_errors():
    asm = Assembler(synth.vocab_file('clevr'))
    toks = asm.module_list2tokens(['_Find', '_Count'], T=5)
    assert toks == [1, 8, 14, 14, 14]
    assert asm.module_list2tokens(['_Find', '_Count']) == [1, 8]
    with pytest.raises(ValueError, match='Not enough time steps'):
        asm.module_list2tokens(['_Find', '_Count'], T=2)


def test_hand_derived_cases():
    asm = Assembler(synth.vocab_file('clevr'))
    T = 8
    cases = {
        'chain': (['_Find', '_Transform', '_Filter', '_Count'], True, None),
        'underflow': (['_And', '_Count'], False, 'not enough input for _And'),
        'att_root': (['_Find', '_Transform'], False, 'result type must be ans, not att'),
        'leftover': (['_Find', '_Find', '_Count'], False,
                     'final stack size not equal to 1 (2 remains)'),
        'ans_input': (['_Find', '_Count', '_Exist'], False, 'input incompatible for _Exist'),
    }
    cols = np.array([asm.module_list2tokens(v[0], T) for v in cases.values()], np.int32).T
    exprs, valid = asm.assemble(cols)
    for (name, (_, ok, err)), e, v in zip(cases.items(), exprs, valid):
        assert bool(v) == ok, name
        if not ok:
            assert e['module'] == INVALID_EXPR and e['error'] == err, name
    chain = exprs[0]
    assert chain['module'] == '_Count' and chain['time_idx'] == 3 and chain['batch_idx'] == 0
    assert chain['input_0']['module'] == '_Filter'
    assert chain['input_0']['input_0']['module'] == '_Transform'
    # operand order: the last popped is input_0
    two = asm.assemble(np.array([asm.module_list2tokens(
        ['_Find', '_Scene', '_And', '_Exist'], T)], np.int32).T)[0][0]
    assert two['input_0']['input_0']['module'] == '_Find'
    assert two['input_0']['input_1']['module'] == '_Scene'
    # no <eos> at all
    e, v = asm.assemble(np.full((T, 1), 1, np.int32))
    assert not v[0] and e[0]['error'] == 'cannot find <eos>'


def test_sampler_only_emits_valid_layouts():
    asm = Assembler(synth.vocab_file('clevr'))
    toks = synth.random_valid_tokens(asm, 64, 20, seed=7)
    _, valid = asm.assemble(toks)
    assert valid.all()
    deep = synth.random_valid_tokens(asm, 16, 20, seed=8, ans_weight=0.15, min_depth=4,
                                     max_depth=12)
    _, valid = asm.assemble(deep)
    assert valid.all()
    assert all(4 <= synth.layout_depth(asm, deep[:, i]) <= 12 for i in range(16))

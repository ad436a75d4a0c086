"""CPU: the numpy oracle reproduces the goldens made by executing the reference's module code
(tests/golden/make_golden.py), and its two executors agree with each other."""
import numpy as np
import pytest

from n2nmn_b200 import synth
from n2nmn_b200.assembler import Assembler
from oracle.nmn_oracle import OracleModules, run_depth_batched, run_sequential
from tests.helpers import case_inputs, load_golden, node_inputs

FAMILIES = ['clevr', 'shapes', 'vqa']


@pytest.mark.parametrize('family', FAMILIES)
def test_oracle_modules_match_reference_goldens(family):
    z, meta = load_golden(family)
    feat, word_vecs, W = case_inputs(meta)
    m = OracleModules(feat, word_vecs, meta['C'], W, family=family)
    for k, (name, arity) in enumerate(meta['module_calls']):
        t, b, a0, a1 = node_inputs(meta, 5, meta['node_seed_base'] + k)
        out = getattr(m, name)(*(a0, a1)[:arity], t, b)
        ref = z['mod_' + name]
        assert out.shape == ref.shape, name
        np.testing.assert_allclose(out, ref, rtol=0, atol=2e-5, err_msg=name)


@pytest.mark.parametrize('family', FAMILIES)
def test_oracle_executors_match_reference_goldens(family):
    z, meta = load_golden(family)
    feat, word_vecs, W = case_inputs(meta)
    m = OracleModules(feat, word_vecs, meta['C'], W, family=family)
    asm = Assembler(synth.vocab_file(family))
    expr_list, validity = asm.assemble(z['exec_tokens'])
    assert validity.tolist() == [bool(v) for v in z['exec_validity']]
    s_seq, att_seq = run_sequential(m, expr_list, return_att=True)
    s_bat, att_bat = run_depth_batched(m, expr_list, return_att=True)
    np.testing.assert_allclose(s_seq, z['exec_scores'], rtol=0, atol=5e-5)
    np.testing.assert_allclose(s_bat, z['exec_scores'], rtol=0, atol=5e-5)
    n_att = 0
    for (b, t), a in att_bat.items():
        np.testing.assert_allclose(a, z['att_b%d_t%d' % (b, t)], rtol=0, atol=5e-5)
        np.testing.assert_allclose(a, att_seq[(b, t)], rtol=0, atol=1e-5)
        n_att += 1
    assert n_att == sum(1 for k in z.files if k.startswith('att_b'))
    # invalid layouts -> zero rows (models_clevr/nmn3_model.py:144-155)
    for i, v in enumerate(validity):
        if not v:
            assert not s_seq[i].any() and not s_bat[i].any()


def test_fp64_floor_is_small():
    """fp32 oracle vs fp64 oracle: the error floor used to budget the tensor-core path."""
    z, meta = load_golden('clevr')
    feat, word_vecs, W = case_inputs(meta)
    m32 = OracleModules(feat, word_vecs, meta['C'], W, family='clevr')
    m64 = OracleModules(feat, word_vecs, meta['C'], W, family='clevr', dtype=np.float64)
    t, b, a0, a1 = node_inputs(meta, 5, 1234)
    for name, arity in meta['module_calls']:
        o32 = getattr(m32, name)(*(a0, a1)[:arity], t, b)
        o64 = getattr(m64, name)(*(a0, a1)[:arity], t, b)
        assert np.max(np.abs(o32 - o64)) < 5e-5, name


def test_known_answers():
    """Hand-derivable cases."""
    z, meta = load_golden('clevr')
    feat, word_vecs, W = case_inputs(meta)
    H, Wd, C = meta['H'], meta['W'], meta['C']
    t = np.zeros(2, np.int32)
    b = np.arange(2, dtype=np.int32)
    # Scene is the constant 3
    m = OracleModules(feat, word_vecs, C, W)
    assert np.all(m.SceneModule(t, b) == 3.0)
    # text projection zeroed -> all-zero product -> eps branch -> output = conv_eltwise bias
    W0 = dict(W)
    W0['FindModule/fc_text/weights'] = np.zeros_like(W['FindModule/fc_text/weights'])
    W0['FindModule/fc_text/biases'] = np.zeros_like(W['FindModule/fc_text/biases'])
    out = OracleModules(feat, word_vecs, C, W0).FindModule(t, b)
    np.testing.assert_allclose(out, W['FindModule/conv_eltwise/biases'][0], atol=0)
    # Exist on a constant map: [c, c, c] @ W + b
    a = np.full((1, H, Wd, 1), 0.5, np.float32)
    ex = m.ExistModule(a, t[:1], b[:1])
    want = 0.5 * W['ExistModule/fc_scores/weights'].sum(0) + W['ExistModule/fc_scores/biases']
    np.testing.assert_allclose(ex[0], want, atol=1e-6)
    # Count weight-row order: one-hot at (y,x) picks row y*W+x, plus min(=0) and max(=1) rows
    a = np.zeros((1, H, Wd, 1), np.float32)
    a[0, 3, 7, 0] = 1.0
    cw = W['CountModule/fc_scores/weights']
    want = cw[3 * Wd + 7] + cw[H * Wd + 1] + W['CountModule/fc_scores/biases']
    np.testing.assert_allclose(m.CountModule(a, t[:1], b[:1])[0], want, atol=1e-6)

"""CPU: the torch-autograd oracle used for the backward pass agrees with the numpy oracle (and so
with the reference goldens) on the forward pass, and its hand-set TF tie rules behave."""
import numpy as np
import pytest
import torch

from n2nmn_b200 import synth
from n2nmn_b200.assembler import Assembler
from oracle import nmn_oracle_torch as ot
from tests.helpers import case_inputs, load_golden


@pytest.mark.parametrize('family', ['clevr', 'shapes', 'vqa'])
def test_torch_oracle_forward_matches_goldens(family):
    z, meta = load_golden(family)
    feat, word_vecs, W = case_inputs(meta)
    m = ot.TorchOracleModules(feat, word_vecs, meta['C'], W, family=family)
    asm = Assembler(synth.vocab_file(family))
    exprs, valid = asm.assemble(z['exec_tokens'])
    scores = ot.forward_scores(m, exprs).detach().numpy()
    np.testing.assert_allclose(scores, z['exec_scores'], rtol=0, atol=5e-5)
    # the depth-batched torch executor (bench.py's second CPU port) gives the same rows
    np.testing.assert_allclose(ot.run_depth_batched(m, exprs), z['exec_scores'], rtol=0, atol=5e-5)


def test_tf_tie_rules_and_loss():
    x = torch.tensor([1.0, 2.0, 3.0], requires_grad=True)
    y = torch.tensor([1.0, 5.0, 0.0], requires_grad=True)
    ot._TFMin.apply(x, y).sum().backward()
    assert x.grad.tolist() == [1.0, 1.0, 0.0] and y.grad.tolist() == [0.0, 0.0, 1.0]
    x.grad = None; y.grad = None
    ot._TFMax.apply(x, y).sum().backward()
    assert x.grad.tolist() == [1.0, 0.0, 1.0] and y.grad.tolist() == [0.0, 1.0, 0.0]
    # loss: invalid rows contribute the constant 0.5 and no gradient
    z, meta = load_golden('clevr')
    feat, word_vecs, W = case_inputs(meta)
    m = ot.TorchOracleModules(feat, word_vecs, meta['C'], W)
    asm = Assembler(synth.vocab_file('clevr'))
    exprs, valid = asm.assemble(z['exec_tokens'])
    labels = np.arange(len(exprs)) % meta['C']
    scores, per, avg, g, g_wv = ot.loss_and_grads(m, exprs, valid, labels)
    assert per[3] == 0.5 and not valid[3]
    assert abs(avg - per.mean()) < 1e-6
    assert all(np.isfinite(v).all() for v in g.values())
    assert np.abs(g['FindModule/conv_image/weights']).max() > 0
    assert not g_wv[:, 3].any()          # question 3 is invalid: nothing flows to its words

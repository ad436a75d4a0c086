"""CPU: the numpy restatement of AttentionSeq2Seq (oracle/seq2seq_oracle.py) against the goldens
produced by executing the reference's nmn3_netgen_att.py on the TF shim."""
import os

import numpy as np

from n2nmn_b200 import synth
from n2nmn_b200.assembler import Assembler
from oracle import seq2seq_oracle as so

Z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'golden_seq2seq.npz'))


def golden_weights():
    return {k[2 + len('encoder_decoder/'):]: Z[k] for k in Z.files if k.startswith('w:')}


def test_encoder_and_both_decoding_modes_match_reference():
    N, T_enc, T_dec, V_txt, E_txt, E_nmn, L, layers, seed = [int(v) for v in Z['cfg']]
    asm = Assembler(synth.vocab_file('clevr'))
    w = golden_weights()
    enc, dec = so.run(w, Z['input_seq'], Z['seq_length'], T_dec, layers, asm.P, asm.W, asm.b)
    emb, outs, state, ht, nf = enc
    np.testing.assert_allclose(outs, Z['encoder_outputs'], atol=2e-6)
    np.testing.assert_allclose(ht, Z['encoder_h_transformed'], atol=2e-6)
    for l in range(layers):
        np.testing.assert_allclose(state[l][0], Z['encoder_c%d' % l], atol=2e-6)
        np.testing.assert_allclose(state[l][1], Z['encoder_h%d' % l], atol=2e-6)
    tokens, probs, nent, wv, atts = dec
    assert np.array_equal(tokens, Z['greedy_predicted_tokens'])
    np.testing.assert_allclose(probs, Z['greedy_token_probs'], atol=2e-6)
    np.testing.assert_allclose(nent, Z['greedy_neg_entropy'], atol=1e-5)
    np.testing.assert_allclose(wv, Z['greedy_word_vecs'], atol=2e-6)
    np.testing.assert_allclose(atts, Z['greedy_atts'], atol=2e-6)
    assert asm.assemble(tokens)[1].all()          # the validity masks make every layout parse
    _, dec = so.run(w, Z['input_seq'], Z['seq_length'], T_dec, layers, asm.P, asm.W, asm.b,
                    use_gt_layout=True, gt_layout=Z['gt_layout'])
    assert np.array_equal(dec[0], Z['gt_layout'])
    np.testing.assert_allclose(dec[1], Z['gt_token_probs'], atol=2e-6)
    np.testing.assert_allclose(dec[2], Z['gt_neg_entropy'], atol=1e-5)
    np.testing.assert_allclose(dec[3], Z['gt_word_vecs'], atol=2e-6)


def test_sampled_decoding_matches_reference():
    """decoder_sampling=True (nmn3_netgen_att.py:234-256) with the golden's uniform numbers: the
    reference file itself ran on the shim's inverse-CDF `tf.multinomial`."""
    N, T_enc, T_dec, V_txt, E_txt, E_nmn, L, layers, seed = [int(v) for v in Z['cfg']]
    asm = Assembler(synth.vocab_file('clevr'))
    margins = []
    _, dec = so.run(golden_weights(), Z['input_seq'], Z['seq_length'], T_dec, layers, asm.P, asm.W,
                    asm.b, sample_uniforms=Z['sample_uniforms'], margins=margins)
    assert np.array_equal(dec[0], Z['sample_predicted_tokens'])
    assert (dec[0] != Z['greedy_predicted_tokens']).any()
    np.testing.assert_allclose(dec[1], Z['sample_token_probs'], atol=2e-6)
    np.testing.assert_allclose(dec[2], Z['sample_neg_entropy'], atol=1e-5)
    np.testing.assert_allclose(dec[3], Z['sample_word_vecs'], atol=2e-6)
    np.testing.assert_allclose(dec[4], Z['sample_atts'], atol=2e-6)
    assert asm.assemble(dec[0])[1].all()
    # no draw sits close enough to a CDF boundary for an fp32 scan (the GPU) to flip it
    assert np.min(margins) > 2e-5, np.min(margins)

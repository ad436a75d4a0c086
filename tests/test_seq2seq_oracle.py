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


def test_encoder_matches_torch_lstm_independently():
    """Independent pin of the restated TF semantics (BasicLSTMCell gate order i,j,f,o with
    forget_bias 1; dynamic_rnn: zero output and carried state past the sequence end): the oracle's
    encoder against torch.nn.LSTM over packed sequences, weights mapped TF -> torch
    (rows [x; h] -> weight_ih / weight_hh, gates i,j,f,o -> i,f,g,o, +1 on the forget bias)."""
    import torch
    from n2nmn_b200.weights import init_seq2seq_weights
    rng = np.random.RandomState(5)
    V_txt, E, L, layers, T, N = 30, 20, 24, 2, 9, 7
    w = init_seq2seq_weights(V_txt, E, 15, 12, L, layers, seed=3)
    for k in list(w):       # biases are zero-initialised: make them count
        if k.endswith('biases'):
            w[k] = (0.2 * rng.standard_normal(w[k].shape)).astype(np.float32)
    seq = rng.randint(0, V_txt, size=(T, N)).astype(np.int32)
    lens = rng.randint(1, T + 1, size=N).astype(np.int32)
    lens[0], lens[1] = T, 1
    emb, outs, state, ht, nf = so.encode(w, seq, lens, layers)

    lstm = torch.nn.LSTM(E, L, num_layers=layers)
    perm = np.concatenate([np.arange(0, L), np.arange(2 * L, 3 * L), np.arange(L, 2 * L),
                           np.arange(3 * L, 4 * L)])           # TF i,j,f,o -> torch i,f,g,o
    with torch.no_grad():
        for l in range(layers):
            W, b = so._cell_vars(w, 'encoder', l)
            nin = E if l == 0 else L
            b = b.copy()
            b[2 * L:3 * L] += 1.0                                # forget_bias
            getattr(lstm, 'weight_ih_l%d' % l).copy_(torch.from_numpy(W[:nin, perm].T.copy()))
            getattr(lstm, 'weight_hh_l%d' % l).copy_(torch.from_numpy(W[nin:, perm].T.copy()))
            getattr(lstm, 'bias_ih_l%d' % l).copy_(torch.from_numpy(b[perm]))
            getattr(lstm, 'bias_hh_l%d' % l).zero_()
        packed = torch.nn.utils.rnn.pack_padded_sequence(
            torch.from_numpy(emb), torch.from_numpy(lens.astype(np.int64)), enforce_sorted=False)
        y, (hn, cn) = lstm(packed)
        y, _ = torch.nn.utils.rnn.pad_packed_sequence(y, total_length=T)
    np.testing.assert_allclose(outs, y.numpy(), atol=2e-6)
    for l in range(layers):
        np.testing.assert_allclose(state[l][0], cn[l].numpy(), atol=2e-6)
        np.testing.assert_allclose(state[l][1], hn[l].numpy(), atol=2e-6)

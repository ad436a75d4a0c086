"""CPU ORACLE (test infrastructure, not product code) — numpy restatement of the attentional
seq2seq layout generator ``AttentionSeq2Seq`` (models_clevr/nmn3_netgen_att.py:46-322; the VQA
copy is identical) without dropout: greedy decoding under the validity masks, teacher forcing
with ground-truth layouts, or sampling (decoder_sampling=True, :234-256) with caller-supplied
uniform numbers (inverse-CDF; TF's own multinomial generator is not reproducible outside TF).

Pinned against golden vectors produced by executing the reference file itself on the numpy TF shim
(tests/golden/make_golden_seq2seq.py -> golden_seq2seq.npz); the TF-op semantics it relies on
(BasicLSTMCell gate order i,j,f,o with forget_bias 1, dynamic_rnn's zero output / state
carry-through past the sequence end, raw_rnn's loop_fn protocol) are restated from the TF 1.0
sources — the residual "parity unpinned" part, as for the module network.

Weights: dict keyed by the TF variable names relative to ``encoder_decoder/``.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell(x, c, h, w, b):
    """BasicLSTMCell(forget_bias=1): gates = [x, h]·W + b split as i, j, f, o
    (nmn3_netgen_att.py:17-44 builds MultiRNNCell([BasicLSTMCell]*num_layers))."""
    g = np.concatenate([x, h], axis=1) @ w + b
    i, j, f, o = np.split(g, 4, axis=1)
    c2 = c * _sigmoid(f + np.float32(1.0)) + _sigmoid(i) * np.tanh(j)
    return c2.astype(np.float32), (np.tanh(c2) * _sigmoid(o)).astype(np.float32)


def _cell_vars(w, side, l):
    p = '%s/lstm/multi_rnn_cell/cell_%d/basic_lstm_cell/' % (side, l)
    return w[p + 'weights'], w[p + 'biases']


def encode(w, input_seq, seq_length, num_layers):
    """_build_encoder (nmn3_netgen_att.py:73-120). Returns embedded_seq [T,N,E], encoder_outputs
    [T,N,L] (zero past the sequence end), states [(c,h)] per layer, encoder_h_transformed,
    seq_not_finished [T,N,1]."""
    T, N = input_seq.shape
    emb = w['encoder/embedding_mat'][input_seq]                       # :88
    L = w['encoder/encoder_h_transform/weights'].shape[0]
    state = [(np.zeros((N, L), np.float32), np.zeros((N, L), np.float32)) for _ in range(num_layers)]
    outs = np.zeros((T, N, L), np.float32)
    for t in range(T):                                                # dynamic_rnn, :95-99
        live = (t < seq_length)[:, None]
        x = emb[t]
        new = []
        for l in range(num_layers):
            wl, bl = _cell_vars(w, 'encoder', l)
            c2, h2 = lstm_cell(x, state[l][0], state[l][1], wl, bl)
            new.append((np.where(live, c2, state[l][0]), np.where(live, h2, state[l][1])))
            x = h2
        outs[t] = np.where(live, x, 0.0)
        state = new
    ht = (outs.reshape(-1, L) @ w['encoder/encoder_h_transform/weights'] +
          w['encoder/encoder_h_transform/biases']).reshape(T, N, L).astype(np.float32)   # :104-108
    not_finished = (np.arange(T)[:, None, None] < seq_length[:, None]).astype(np.float32)  # :112-116
    return emb, outs, state, ht, not_finished


def decode(w, enc, T_dec, num_layers, P, W, b, use_gt_layout=False, gt_layout=None,
           sample_uniforms=None, margins=None):
    """_build_decoder (nmn3_netgen_att.py:122-322), greedy (decoder_sampling=False), sampled
    (`sample_uniforms` [T_dec,N] in [0,1): token = first class whose cumulative probability under
    softmax(scores - 50·invalid) exceeds u, greedy fallback if that token is invalid, :234-256) or
    teacher forced. Returns predicted_tokens [T_dec,N] int32, token_probs [T_dec,N], neg_entropy
    [N], word_vecs [T_dec,N,E], atts [T_dec,T_enc,N,1]. `margins` (a list) receives, per step, the
    distance of u from the nearest CDF boundary — how far a draw is from flipping under another
    summation order."""
    emb, outs, state, ht, not_finished = enc
    N = emb.shape[1]
    V = w['decoder/embedding_mat'].shape[0]
    Wa, ba, v = (w['decoder/att_prediction/weights'], w['decoder/att_prediction/biases'],
                 w['decoder/att_prediction/v'])
    Wy, by = w['decoder/token_prediction/weights'], w['decoder/token_prediction/biases']
    x = np.tile(w['decoder/go_embedding'], (N, 1))                    # :202
    X = np.tile(np.array([[0, 0, T_dec]], np.int64), (N, 1))         # :293
    tokens = np.zeros((T_dec, N), np.int32)
    probs_out = np.zeros((T_dec, N), np.float32)
    atts = np.zeros((T_dec,) + ht.shape[:2] + (1,), np.float32)
    neg_entropy = np.zeros(N, np.float32)
    for t in range(T_dec):
        new = []
        for l in range(num_layers):
            wl, bl = _cell_vars(w, 'decoder', l)
            c2, h2 = lstm_cell(x, state[l][0], state[l][1], wl, bl)
            new.append((c2, h2))
            x = h2
        state = new
        out = x
        att_raw = np.sum(np.tanh((out @ Wa + ba) + ht) * v, axis=2, keepdims=True)      # :208-212
        e = np.exp(att_raw - att_raw.max(axis=0, keepdims=True))
        att = e / e.sum(axis=0, keepdims=True) * not_finished                            # :215
        att = att / att.sum(axis=0, keepdims=True)                                       # :216
        d2 = np.sum(att * outs, axis=0)                                                   # :218
        scores = (np.concatenate([out, d2], axis=1) @ Wy + by).astype(np.float32)        # :221-223
        valid = np.all(np.tensordot(X, W.astype(np.int64), axes=1) - b >= 0, axis=2)     # :8-11
        if use_gt_layout:
            valid = np.ones_like(valid)                                                   # :230-233
        vm = valid.astype(np.float32)
        masked = np.where(valid, scores, scores.min() - 1)                                # :259-261
        pred = np.argmax(masked, axis=1).astype(np.int32)
        if sample_uniforms is not None:
            z = scores.astype(np.float64) - (1.0 - vm) * 50.0                             # :235
            q = np.exp(z - z.max(axis=1, keepdims=True))
            cdf = np.cumsum(q, axis=1) / q.sum(axis=1, keepdims=True)
            u = np.asarray(sample_uniforms[t], np.float64)
            samp = np.minimum((cdf <= u[:, None]).sum(axis=1), V - 1)                     # :238-239
            pred = np.where(valid[np.arange(N), samp], samp, pred).astype(np.int32)       # :244-256
            if margins is not None:
                margins.append(np.abs(cdf[:, :-1] - u[:, None]).min(axis=1))
        if use_gt_layout:
            pred = gt_layout[t].astype(np.int32)                                          # :264-266
        es = np.exp(scores - scores.max(axis=1, keepdims=True))
        all_p = es / es.sum(axis=1, keepdims=True) * vm                                   # :270
        all_p = all_p / all_p.sum(axis=1, keepdims=True)                                  # :272
        probs_out[t] = all_p[np.arange(N), pred]                                          # :281
        neg_entropy += np.sum(all_p * np.log(np.maximum(1e-5, all_p + (1 - vm))), axis=1)  # :283-285
        X = X + P[pred]                                                                   # :288-289
        tokens[t] = pred
        atts[t] = att
        x = w['decoder/embedding_mat'][pred]                                              # :293
    word_vecs = np.sum(atts * emb[None], axis=1)                                          # :312
    return tokens, probs_out, neg_entropy.astype(np.float32), word_vecs.astype(np.float32), atts


def run(w, input_seq, seq_length, T_dec, num_layers, P, W, b, use_gt_layout=False,
        gt_layout=None, sample_uniforms=None, margins=None):
    enc = encode(w, np.asarray(input_seq), np.asarray(seq_length), num_layers)
    return enc, decode(w, enc, T_dec, num_layers, np.asarray(P), np.asarray(W), np.asarray(b),
                       use_gt_layout, gt_layout, sample_uniforms, margins)

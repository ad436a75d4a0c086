"""Host-side mirror of the reference's layout generator ``AttentionSeq2Seq``
(models_clevr/nmn3_netgen_att.py:46-322; models_vqa/ and models_shapes/ carry the same file) over
the C ABI (`n2nmn_seq2seq_*`, include/n2nmn_b200.h). Forward pass without dropout: greedy decoding
under the Assembler's validity masks, `decoder_sampling=True` (one draw per step from the masked
token distribution, nmn3_netgen_att.py:234-256; the uniform numbers come from `torch.rand` on the
device or from the caller, `forward(..., sample_uniforms=)`) or teacher forcing (`use_gt_layout`).

The reference builds a TF graph whose placeholders are fed per batch; here the constructor takes
the shapes (and optionally the first batch) and ``forward`` / ``__call__`` runs a batch. The
attribute names after a forward are the reference's: ``predicted_tokens`` [T_decoder, N] int32,
``token_probs`` [T_decoder, N], ``neg_entropy`` [N], ``word_vecs`` [T_decoder, N, embed_dim_txt],
``atts`` [T_decoder, T_encoder, N, 1]; ``log_seq_prob`` as nmn3_model.py:45 computes it.
PyTorch only owns the device buffers; there is no CPU path.

`precision='fp32'` (default): every matrix product with fp32 parity (error-compensated TF32 on the
tensor cores); `'tf32'`: one TF32 pass (13 % faster at batch 64, where the step is latency bound),
probabilities within ~1e-3, a token may
differ when two scores are that close.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check


class AttentionSeq2Seq:
    def __init__(self, input_seq_batch, seq_length_batch, T_decoder, num_vocab_txt, embed_dim_txt,
                 num_vocab_nmn, embed_dim_nmn, lstm_dim, num_layers, assembler,
                 encoder_dropout=False, decoder_dropout=False, decoder_sampling=False,
                 use_gt_layout=None, gt_layout_batch=None, scope='encoder_decoder', reuse=None,
                 T_encoder=None, max_batch=None, device=None, weights=None, precision='fp32'):
        if encoder_dropout or decoder_dropout:
            raise NotImplementedError('dropout is a training-time option; the B200 seq2seq is the '
                                      'inference configuration')
        self.decoder_sampling = bool(decoder_sampling)
        self.T_decoder = int(T_decoder)
        self.encoder_num_vocab, self.encoder_embed_dim = int(num_vocab_txt), int(embed_dim_txt)
        self.decoder_num_vocab, self.decoder_embed_dim = int(num_vocab_nmn), int(embed_dim_nmn)
        self.lstm_dim, self.num_layers = int(lstm_dim), int(num_layers)
        self.EOS_token = assembler.EOS_idx
        self.scope = scope
        if input_seq_batch is not None:
            T_encoder = T_encoder or int(input_seq_batch.shape[0])
            max_batch = max_batch or int(input_seq_batch.shape[1])
            if device is None and isinstance(input_seq_batch, torch.Tensor) and input_seq_batch.is_cuda:
                device = input_seq_batch.device
        if T_encoder is None or max_batch is None:
            raise ValueError('give input_seq_batch or T_encoder and max_batch')
        self.device = torch.device(device if device is not None else 'cuda:0')
        if self.device.type != 'cuda':
            raise _lib.N2NMNError('n2nmn_b200 runs on a CUDA (sm_100) device only')
        self.T_encoder, self.max_batch = int(T_encoder), int(max_batch)
        self._L = _lib.lib()
        cfg = _lib.Seq2SeqConfig(_lib.ABI_VERSION, self.encoder_num_vocab, self.encoder_embed_dim,
                                 self.decoder_num_vocab, self.decoder_embed_dim, self.lstm_dim,
                                 self.num_layers, self.T_encoder, self.T_decoder, self.max_batch,
                                 self.device.index or 0, {'fp32': 0, 'tf32': 1}[precision])
        h = C.c_void_p()
        check(self._L.n2nmn_seq2seq_create(C.byref(cfg), C.byref(h)))
        self._h = h
        P = np.ascontiguousarray(assembler.P, np.int32)
        W = np.ascontiguousarray(assembler.W, np.int32)
        b = np.ascontiguousarray(assembler.b, np.int32)
        V = self.decoder_num_vocab
        if P.shape != (V, 3) or W.shape != (3, V, 4) or b.shape != (V, 4):
            raise ValueError('assembler tables do not match num_vocab_nmn')
        i32 = C.POINTER(C.c_int32)
        with torch.cuda.device(self.device):
            check(self._L.n2nmn_seq2seq_set_assembler(self._h, P.ctypes.data_as(i32),
                                                      W.ctypes.data_as(i32), b.ctypes.data_as(i32),
                                                      self._stream()))
        self._keep = None
        if weights is not None:
            self.set_weights(weights)
        self.use_gt_layout, self.gt_layout_batch = use_gt_layout, gt_layout_batch
        if input_seq_batch is not None and weights is not None:
            self.forward(input_seq_batch, seq_length_batch, use_gt_layout, gt_layout_batch)

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._L.n2nmn_seq2seq_destroy(h)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def variables(self):
        """[(name relative to `<scope>/`, shape)] in creation order."""
        out = []
        for i in range(self._L.n2nmn_seq2seq_num_variables(self._h)):
            name, shape, nd = C.c_char_p(), (C.c_int64 * 4)(), C.c_int()
            check(self._L.n2nmn_seq2seq_variable_info(self._h, i, C.byref(name), shape, C.byref(nd)))
            out.append((name.value.decode(), tuple(shape[:nd.value])))
        return out

    def set_weights(self, weights):
        """weights: {TF variable name: array}; names may carry any prefix ending in
        ``<scope>/`` (e.g. ``neural_module_network/layout_generation/encoder_decoder/…``)."""
        tag = self.scope + '/'
        rel = {}
        for k, v in weights.items():
            k = k[2:] if k.startswith('w:') else k
            k = k.split(':')[0]
            if tag in k:
                rel[k[k.index(tag) + len(tag):]] = v
            else:
                rel[k] = v
        with torch.cuda.device(self.device):
            for name, shape in self.variables():
                if name not in rel:
                    raise KeyError('missing seq2seq weight %s' % name)
                t = torch.as_tensor(np.asarray(rel[name], np.float32) if not isinstance(
                    rel[name], torch.Tensor) else rel[name]).to(self.device, torch.float32).contiguous()
                if tuple(t.shape) != shape:
                    raise ValueError('shape of %s is %s, expected %s' % (name, tuple(t.shape), shape))
                shp = (C.c_int64 * len(shape))(*shape)
                check(self._L.n2nmn_seq2seq_set_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()),
                                                       shp, len(shape), self._stream()))
                torch.cuda.current_stream(self.device).synchronize()   # t may be a temporary

    def forward(self, input_seq_batch, seq_length_batch, use_gt_layout=None, gt_layout_batch=None,
                sample_uniforms=None):
        """input_seq_batch [T_enc, N] int, seq_length_batch [N] int (device or host);
        gt_layout_batch [T_decoder, N] with use_gt_layout truthy = teacher forcing;
        sample_uniforms [T_decoder, N] in [0, 1): the numbers the sampled decoding consumes
        (decoder_sampling=True; default `torch.rand` on the device, i.e. torch's generator)."""
        dev = self.device

        def i32(x):
            t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
            return t.to(dev, torch.int32, non_blocking=True).contiguous()
        seq, lens = i32(input_seq_batch), i32(seq_length_batch)
        T, N = seq.shape
        if lens.shape != (N,):
            raise ValueError('seq_length_batch must have one entry per question')
        if T > self.T_encoder or N > self.max_batch:
            raise _lib.N2NMNError('batch exceeds the capacity the seq2seq was created for')
        gt = None
        if use_gt_layout:
            gt = i32(gt_layout_batch)
            if gt.shape != (self.T_decoder, N):
                raise ValueError('gt_layout_batch must be [T_decoder, N]')
        u = None
        if self.decoder_sampling or sample_uniforms is not None:
            if sample_uniforms is None:
                u = torch.rand((self.T_decoder, N), dtype=torch.float32, device=dev)
            else:
                u = (sample_uniforms if isinstance(sample_uniforms, torch.Tensor) else
                     torch.as_tensor(np.asarray(sample_uniforms, np.float32)))
                u = u.to(dev, torch.float32).contiguous()
                if u.shape != (self.T_decoder, N):
                    raise ValueError('sample_uniforms must be [T_decoder, N]')
        tokens = torch.empty((self.T_decoder, N), dtype=torch.int32, device=dev)
        probs = torch.empty((self.T_decoder, N), dtype=torch.float32, device=dev)
        ent = torch.empty((N,), dtype=torch.float32, device=dev)
        wv = torch.empty((self.T_decoder, N, self.encoder_embed_dim), dtype=torch.float32, device=dev)
        atts = torch.empty((self.T_decoder, T, N, 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(self._L.n2nmn_seq2seq_set_sampling(
                self._h, C.c_void_p(u.data_ptr()) if u is not None else None))
            check(self._L.n2nmn_seq2seq_forward(
                self._h, C.c_void_p(seq.data_ptr()), C.c_void_p(lens.data_ptr()), T, N,
                C.c_void_p(gt.data_ptr()) if gt is not None else None,
                C.c_void_p(tokens.data_ptr()), C.c_void_p(probs.data_ptr()),
                C.c_void_p(ent.data_ptr()), C.c_void_p(wv.data_ptr()), C.c_void_p(atts.data_ptr()),
                self._stream()))
        self._keep = (seq, lens, gt, u)   # alive until the stream has consumed them
        self.predicted_tokens, self.token_probs, self.neg_entropy = tokens, probs, ent
        self.word_vecs, self.atts = wv, atts
        self.log_seq_prob = torch.log(probs).sum(0)          # nmn3_model.py:45
        return tokens, probs, ent, wv, atts

    __call__ = forward

    def launch_count(self):
        return int(self._L.n2nmn_seq2seq_launch_count(self._h))

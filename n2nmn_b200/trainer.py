"""Policy-search train step of the module network (exp_clevr/train_clevr_rl_gt_layout.py:101-139).

    loss_i   = softmax_cross_entropy(scores_i, label_i)  if layout i is valid else 0.5      (:108-114)
    avg      = mean_i loss_i                                                               (:119)
    baseline += (1 - 0.99) * (avg - baseline)             (EMA, not trained)               (:120-122)
    pg       = mean_i stop_gradient(loss_i - baseline) * log_seq_prob_i                    (:123-124)
    total    = pg + avg + 0.005 * entropy_reg + 5e-6 * l2_reg                              (:126-129)
    Adam(1e-4) on per-tensor clip_by_norm(grad, 10)                                        (:132-139)

The module network receives gradient only through `avg` and the l2 term; `pg` and the entropy term
reach the seq2seq layout generator, which is off the hot path: this class returns what that
generator needs (d total/d word_vecs and the per-sample REINFORCE coefficients) instead of
differentiating it. Data parallel: every rank holds a shard of the questions; ONE all-reduce over
the flat gradient buffer (+ the loss scalar riding in its last element) per step (NCCL via
torch.distributed), then clip + Adam replicated on every rank.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class ModuleNetTrainer:
    def __init__(self, executor, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, max_grad_l2_norm=10.0,
                 weight_decay=5e-6, baseline_decay=0.99, invalid_expr_loss=0.5,
                 lambda_entropy=0.005, process_group=None):
        self.ex = executor
        self.m = executor.modules
        self._lib = executor._lib
        self.hyper = dict(lr=lr, beta1=beta1, beta2=beta2, eps=eps, max_norm=max_grad_l2_norm,
                          weight_decay=weight_decay)
        self.baseline_decay = baseline_decay
        self.invalid_expr_loss = invalid_expr_loss
        self.lambda_entropy = lambda_entropy
        self.baseline = float(invalid_expr_loss)        # tf.Variable(invalid_expr_loss) (:120)
        self.step = 0
        self.pg = process_group
        h = self.m._h
        self.flat_size = int(self._lib.n2nmn_flat_size(h))
        dev = self.m.device
        # flat buffers; the last element of `g` carries the loss sum through the all-reduce
        self.w = torch.zeros(self.flat_size, dtype=torch.float32, device=dev)
        self.g = torch.zeros(self.flat_size + 1, dtype=torch.float32, device=dev)
        self.m1 = torch.zeros(self.flat_size, dtype=torch.float32, device=dev)
        self.m2 = torch.zeros(self.flat_size, dtype=torch.float32, device=dev)
        self.layout = {}
        nvar = self._lib.n2nmn_num_variables(h)
        for i in range(nvar):
            name = C.c_char_p()
            shape = (C.c_int64 * 4)()
            nd = C.c_int()
            _lib.check(self._lib.n2nmn_variable_info(h, i, C.byref(name), shape, C.byref(nd)))
            off, cnt = C.c_int64(), C.c_int64()
            _lib.check(self._lib.n2nmn_flat_offset(h, i, C.byref(off), C.byref(cnt)))
            self.layout[name.value.decode()] = (off.value, cnt.value, tuple(shape[:nd.value]))
        for name, t in self.m.get_weights().items():
            off, cnt, shp = self.layout[name]
            self.w[off:off + cnt] = t.reshape(-1)
        # the context is re-packed from self.w after every step: hand out those values, not the
        # tensors cached by set_weights() at construction time (replicas, checkpoints)
        self.m._weights_source = self.weights
        self._loss = torch.zeros(1 + self.m.max_batch, dtype=torch.float32, device=dev)
        self._decay_mask = torch.zeros(self.flat_size, dtype=torch.float32, device=dev)
        for name, (off, cnt, _) in self.layout.items():
            if name.endswith('/weights'):
                self._decay_mask[off:off + cnt] = 1.0

    # -- views ------------------------------------------------------------------------------------
    def weights(self):
        return {n: self.w[o:o + c].reshape(s) for n, (o, c, s) in self.layout.items()}

    def l2_reg(self):
        """Σ tf.nn.l2_loss over the '.../weights' variables (nmn3_model.py:163-166)."""
        return float(0.5 * (self.w * self.w * self._decay_mask).sum())

    def grads(self):
        return {n: self.g[o:o + c].reshape(s) for n, (o, c, s) in self.layout.items()}

    # -- one step ---------------------------------------------------------------------------------
    def forward_backward(self, image_feat_grid, word_vecs, layout_tokens, labels,
                         want_dword=True):
        """Forward + backward of this rank's shard. Fills self.g (gradient of the LOCAL mean loss)
        and returns (scores, validity, per_sample_loss tensor, d_word_vecs or None)."""
        m = self.m
        tok = np.ascontiguousarray(layout_tokens, dtype=np.int32)
        T, N = tok.shape
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        assert lab.shape == (N,)
        scores = torch.empty((N, self.ex.num_choices), dtype=torch.float32, device=m.device)
        dword = torch.empty((T, N, m.text_dim), dtype=torch.float32, device=m.device) \
            if want_dword else None
        validity = np.empty(N, np.uint8)
        m.image_feat_grid, m.word_vecs, m.N, m.T = image_feat_grid, word_vecs, N, T
        _lib.check(self._lib.n2nmn_train_backward(
            m._h, image_feat_grid.data_ptr(), word_vecs.data_ptr(), tok.ctypes.data, T, N,
            self.ex._vocab_ptr, len(self.ex.vocab_ops), lab.ctypes.data,
            C.c_float(self.invalid_expr_loss), scores.data_ptr(), self.g.data_ptr(),
            dword.data_ptr() if dword is not None else None, self._loss.data_ptr(),
            validity.ctypes.data, torch.cuda.current_stream(m.device).cuda_stream))
        return scores, validity.view(bool), self._loss[1:1 + N], dword

    def train_step(self, image_feat_grid, word_vecs, layout_tokens, labels, log_seq_prob=None,
                   entropy_reg=0.0):
        """One optimiser step. Returns a dict with the reference's logged quantities."""
        import torch.distributed as dist
        scores, validity, per_sample, dword = self.forward_backward(
            image_feat_grid, word_vecs, layout_tokens, labels)
        N = scores.shape[0]
        world = 1
        self.g[self.flat_size] = self._loss[0] / N          # local mean loss rides along
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            world = dist.get_world_size(self.pg)
            dist.all_reduce(self.g, op=dist.ReduceOp.SUM, group=self.pg)   # the ONE collective
            self.g /= world                                  # equal shards: mean of shard means
            if dword is not None:
                dword /= world
        avg_sample_loss = float(self.g[self.flat_size])      # global mean (syncs the stream)
        # REINFORCE pieces for the (off-path) layout generator; baseline read before its update
        coeff = (per_sample - self.baseline) / (N * world)
        pg_loss = float((coeff * log_seq_prob).sum()) if log_seq_prob is not None else 0.0
        self.baseline += (1.0 - self.baseline_decay) * (avg_sample_loss - self.baseline)
        l2 = self.l2_reg()                                   # of the weights the loss was taken at
        self.step += 1
        hp = self.hyper
        _lib.check(self._lib.n2nmn_adam_step(
            self.m._h, self.w.data_ptr(), self.g.data_ptr(), self.m1.data_ptr(),
            self.m2.data_ptr(), self.step, hp['lr'], hp['beta1'], hp['beta2'], hp['eps'],
            hp['max_norm'], hp['weight_decay'],
            torch.cuda.current_stream(self.m.device).cuda_stream))
        return {'scores': scores, 'validity': validity, 'avg_sample_loss': avg_sample_loss,
                'policy_gradient_loss': pg_loss, 'baseline': self.baseline,
                'l2_reg': l2,
                'total_loss': pg_loss + avg_sample_loss + self.lambda_entropy * entropy_reg +
                hp['weight_decay'] * l2,
                'd_word_vecs': dword, 'reinforce_coeff': coeff}

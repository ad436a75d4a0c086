"""CPU ORACLE (test infrastructure) — torch-autograd restatement of the module network, used to
check the CUDA backward pass / train step.

Forward math is the numpy oracle's (oracle/nmn_oracle.py, pinned to the reference goldens) op for
op; tests assert the two forwards agree. Backward follows TF 1.0's registered gradients where they
differ from torch's defaults:
  * tf.minimum(x, y): gradient to x where x <= y, to y elsewhere (ties -> input_0);
    tf.maximum(x, y): gradient to x where x >= y (torch splits ties 0.5/0.5);
  * reduce_min / reduce_max split the gradient equally among tied extrema (torch amin/amax do too);
  * l2_normalize = x * rsqrt(max(sum x^2, eps)): the max picks the constant branch below eps.
Loss pieces follow exp_clevr/train_clevr_rl_gt_layout.py:108-139.

Only tests/ and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-12
INVALID_EXPR = 'INVALID_EXPR'


class _TFMin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x <= y)
        return torch.minimum(x, y)

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        return g * m, g * (~m)


class _TFMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x >= y)
        return torch.maximum(x, y)

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        return g * m, g * (~m)


def l2_normalize(x, dim):
    ss = torch.sum(x * x, dim=dim, keepdim=True)
    return x * torch.rsqrt(torch.clamp_min(ss, EPS))


def add_coords(feat):
    n, H, W, _ = feat.shape
    xs = torch.linspace(-1.0, 1.0, W, dtype=feat.dtype).reshape(1, 1, W, 1).expand(n, H, W, 1)
    ys = torch.linspace(-1.0, 1.0, H, dtype=feat.dtype).reshape(1, H, 1, 1).expand(n, H, W, 1)
    return torch.cat([feat, xs, ys], dim=3)


class TorchOracleModules:
    def __init__(self, image_feat_grid, word_vecs, num_choices, weights, family='clevr',
                 dtype=torch.float32):
        self.family = family
        feat = torch.as_tensor(np.asarray(image_feat_grid), dtype=dtype)
        if family == 'vqa':
            feat = add_coords(feat)
        self.feat = feat
        self.word_vecs = torch.as_tensor(np.asarray(word_vecs), dtype=dtype).requires_grad_(True)
        self.C = num_choices
        self.w = {k: torch.as_tensor(np.asarray(v), dtype=dtype).clone().requires_grad_(True)
                  for k, v in weights.items()}
        self.N = self.word_vecs.shape[1]

    def _text(self, t, b):
        return self.word_vecs[torch.as_tensor(t, dtype=torch.long), torch.as_tensor(b, dtype=torch.long)]

    def _fc(self, scope, x):
        return x @ self.w[scope + '/weights'] + self.w[scope + '/biases']

    def _conv1x1(self, scope, x):
        n, H, W, D = x.shape
        return self._fc(scope, x.reshape(-1, D)).reshape(n, H, W, -1)

    def SceneModule(self, t, b):
        n = len(t)
        H, W = self.feat.shape[1:3]
        return torch.full((n, H, W, 1), 3.0, dtype=self.feat.dtype)

    def FindModule(self, t, b, scope='FindModule'):
        feat = self.feat[torch.as_tensor(b, dtype=torch.long)]
        n = len(t)
        mapped = self._conv1x1(scope + '/conv_image', feat)
        tmap = self._fc(scope + '/fc_text', self._text(t, b)).reshape(n, 1, 1, -1)
        return self._conv1x1(scope + '/conv_eltwise', l2_normalize(mapped * tmap, 3))

    def FilterModule(self, a, t, b):
        return _TFMin.apply(a, self.FindModule(t, b))

    def _pooled(self, feat, att):
        n, H, W, _ = feat.shape
        s = F.softmax(att.reshape(n, H * W), dim=1).reshape(n, H, W, 1)
        return torch.sum(feat * s, dim=(1, 2))

    def FindSamePropertyModule(self, a, t, b, scope='FindSamePropertyModule'):
        feat = self.feat[torch.as_tensor(b, dtype=torch.long)]
        n = len(t)
        mapped = self._conv1x1(scope + '/conv_image', feat)
        tmap = self._fc(scope + '/fc_text', self._text(t, b)).reshape(n, 1, 1, -1)
        amap = self._fc(scope + '/fc_att', self._pooled(feat, a)).reshape(n, 1, 1, -1)
        return self._conv1x1(scope + '/conv_eltwise', l2_normalize(mapped * tmap * amap, 3))

    def TransformModule(self, a, t, b, scope='TransformModule'):
        if self.family == 'vqa':
            return self.FindSamePropertyModule(a, t, b, scope=scope)
        n = a.shape[0]
        K = self.w[scope + '/conv_maps/weights']          # [k,k,1,M]
        k = K.shape[0]
        x = a.permute(0, 3, 1, 2)                          # NCHW
        wt = K.permute(3, 2, 0, 1)                         # [M,1,k,k]; conv2d = cross-correlation
        maps = F.conv2d(x, wt, padding=(k - 1) // 2).permute(0, 2, 3, 1) + \
            self.w[scope + '/conv_maps/biases']
        tmap = self._fc(scope + '/text_fc', self._text(t, b)).reshape(n, 1, 1, -1)
        return self._conv1x1(scope + '/conv_eltwise', l2_normalize(maps * tmap, 3))

    def AndModule(self, a0, a1, t=None, b=None):
        return _TFMin.apply(a0, a1)

    def OrModule(self, a0, a1, t=None, b=None):
        return _TFMax.apply(a0, a1)

    def ExistModule(self, a, t=None, b=None, scope='ExistModule'):
        red = torch.cat([a.amin(dim=(1, 2)), a.mean(dim=(1, 2)), a.amax(dim=(1, 2))], dim=1)
        return self._fc(scope + '/fc_scores', red)

    def AnswerModule(self, a, t=None, b=None):
        return self.ExistModule(a, scope='AnswerModule')

    def CountModule(self, a, t=None, b=None):
        n = a.shape[0]
        cat = torch.cat([a.reshape(n, -1), a.amin(dim=(1, 2)), a.amax(dim=(1, 2))], dim=1)
        return self._fc('CountModule/fc_scores', cat)

    def _compare(self, scope, a0, a1):
        parts = []
        for a in (a0, a1):
            parts += [a.reshape(a.shape[0], -1), a.amin(dim=(1, 2)), a.amax(dim=(1, 2))]
        return self._fc(scope + '/fc_scores', torch.cat(parts, dim=1))

    def EqualNumModule(self, a0, a1, t=None, b=None):
        return self._compare('EqualNumModule', a0, a1)

    def MoreNumModule(self, a0, a1, t=None, b=None):
        return self._compare('MoreNumModule', a0, a1)

    def LessNumModule(self, a0, a1, t=None, b=None):
        return self._compare('LessNumModule', a0, a1)

    def SamePropertyModule(self, a0, a1, t, b, scope='SamePropertyModule'):
        feat = self.feat[torch.as_tensor(b, dtype=torch.long)]
        tmap = self._fc(scope + '/fc_text', self._text(t, b))
        p0 = self._fc(scope + '/fc_att_0', self._pooled(feat, a0))
        p1 = self._fc(scope + '/fc_att_1', self._pooled(feat, a1))
        return self._fc(scope + '/fc_eltwise', l2_normalize(p0 * tmap * p1, 1))

    def DescribeModule(self, a, t, b, scope='DescribeModule'):
        feat = self.feat[torch.as_tensor(b, dtype=torch.long)]
        tmap = self._fc(scope + '/fc_text', self._text(t, b))
        amap = self._fc(scope + '/fc_att', self._pooled(feat, a))
        return self._fc(scope + '/fc_eltwise', l2_normalize(tmap * amap, 1))


_TOKEN_METHOD = {
    '_Scene': 'SceneModule', '_Find': 'FindModule', '_Filter': 'FilterModule',
    '_FindSameProperty': 'FindSamePropertyModule', '_Transform': 'TransformModule',
    '_And': 'AndModule', '_Or': 'OrModule', '_Exist': 'ExistModule', '_Count': 'CountModule',
    '_EqualNum': 'EqualNumModule', '_MoreNum': 'MoreNumModule', '_LessNum': 'LessNumModule',
    '_SameProperty': 'SamePropertyModule', '_Describe': 'DescribeModule',
    '_Answer': 'AnswerModule'}


def forward_scores(m, expr_list):
    """Sequential executor (n=1 module calls); returns scores [N,C] with autograd history."""
    rows = []

    def ev(e):
        ins = [ev(e[k]) for k in ('input_0', 'input_1') if k in e]
        return getattr(m, _TOKEN_METHOD[e['module']])(*ins, [e['time_idx']], [e['batch_idx']])

    for e in expr_list:
        if e['module'] == INVALID_EXPR:
            rows.append(torch.zeros(1, m.C, dtype=m.feat.dtype))
        else:
            rows.append(ev(e))
    return torch.cat(rows, dim=0)


def loss_and_grads(m, expr_list, validity, labels, invalid_expr_loss=0.5, weight_decay=0.0):
    """avg_sample_loss of exp_clevr/train_clevr_rl_gt_layout.py:108-119 (+ weight_decay * l2_reg,
    nmn3_model.py:163-166) and its gradients w.r.t. every module variable and word_vecs.
    The REINFORCE and entropy terms have no gradient into the module network (stop_gradient /
    seq2seq-only) and are handled by the caller."""
    scores = forward_scores(m, expr_list)
    ce = F.cross_entropy(scores, torch.as_tensor(labels, dtype=torch.long), reduction='none')
    valid = torch.as_tensor(np.asarray(validity), dtype=torch.bool)
    per_sample = torch.where(valid, ce, torch.full_like(ce, invalid_expr_loss))
    avg = per_sample.mean()
    l2 = sum(0.5 * (w * w).sum() for n, w in m.w.items() if n.endswith('/weights'))
    total = avg + weight_decay * l2
    names = list(m.w)
    grads = torch.autograd.grad(total, [m.w[n] for n in names] + [m.word_vecs], allow_unused=True)
    g = {n: (gi if gi is not None else torch.zeros_like(m.w[n])).detach().numpy()
         for n, gi in zip(names, grads[:-1])}
    g_wv = grads[-1].detach().numpy() if grads[-1] is not None else \
        np.zeros(tuple(m.word_vecs.shape), np.float32)
    return (scores.detach().numpy(), per_sample.detach().numpy(), float(avg), g, g_wv)


def adam_clip_step(weights, grads, state, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8,
                   max_norm=10.0):
    """tf.clip_by_norm per tensor (train_clevr_rl_gt_layout.py:137-138) then tf.train.AdamOptimizer
    (:132): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v updates; w -= lr_t*m/(sqrt(v)+eps)."""
    state['t'] = state.get('t', 0) + 1
    t = state['t']
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    out = {}
    for n, w in weights.items():
        g = np.asarray(grads[n], np.float64)
        nrm = np.sqrt(np.sum(g * g))
        if nrm > max_norm:
            g = g * (max_norm / nrm)
        m = state.setdefault('m', {}).get(n, np.zeros_like(g))
        v = state.setdefault('v', {}).get(n, np.zeros_like(g))
        m = beta1 * m + (1 - beta1) * g
        v = beta2 * v + (1 - beta2) * g * g
        state['m'][n], state['v'][n] = m, v
        out[n] = (np.asarray(w, np.float64) - lr_t * m / (np.sqrt(v) + eps)).astype(np.float32)
    return out


def run_depth_batched(m, expr_list):
    """TF-Fold-like dynamic batching with torch-CPU kernels (MKL / oneDNN, all cores), no autograd:
    the same schedule as oracle/nmn_oracle.py::run_depth_batched (one module call per (depth,
    module type) with leading dim n, gathers materialised). Timed by bench.py as the second CPU
    port of the reference path; its forward is checked against the numpy oracle in
    tests/test_oracle_torch.py."""
    nodes = []   # (depth, module, t, b, child ids, question or -1)

    def walk(e, q_root):
        kids = [walk(e[k], -1) for k in ('input_0', 'input_1') if k in e]
        depth = 1 + max([nodes[k][0] for k in kids], default=0)
        nodes.append((depth, e['module'], e['time_idx'], e['batch_idx'], kids, q_root))
        return len(nodes) - 1

    for q, e in enumerate(expr_list):
        if e['module'] != INVALID_EXPR:
            walk(e, q)
    with torch.no_grad():
        scores = torch.zeros((len(expr_list), m.C), dtype=m.feat.dtype)
        values = [None] * len(nodes)
        max_depth = max([n[0] for n in nodes], default=0)
        for d in range(1, max_depth + 1):
            by_type = {}
            for i, nd in enumerate(nodes):
                if nd[0] == d:
                    by_type.setdefault(nd[1], []).append(i)
            for mod, ids in by_type.items():
                arity = len(nodes[ids[0]][4])
                ins = [torch.stack([values[nodes[i][4][k]] for i in ids]) for k in range(arity)]
                out = getattr(m, _TOKEN_METHOD[mod])(*ins, [nodes[i][2] for i in ids],
                                                     [nodes[i][3] for i in ids])
                for j, i in enumerate(ids):
                    values[i] = out[j]
                    if nodes[i][5] >= 0:
                        scores[nodes[i][5]] = out[j]
    return scores.numpy()

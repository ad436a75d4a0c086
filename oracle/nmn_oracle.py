"""CPU ORACLE (test infrastructure, not product code) — numpy restatement of the N2NMN
module-network forward pass.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package. The product (n2nmn_b200/) never does.

Parity status: the reference (TensorFlow 1.0.0 + TF Fold 0.0.1) cannot run in this image and
ships no tests or golden vectors, so nothing of the *reference's own* pins these numbers. The
restatement is pinned instead against golden vectors produced by executing the reference's
``nmn3_modules.py`` files themselves, unmodified, on a numpy stand-in for the handful of
TF ops they call (oracle/tf1_shim.py; generator tests/golden/make_golden.py; fixtures
tests/golden/*.npz). TF-op semantics in that stand-in are restated from the TF 1.0 API docs
 — that residual assumption is why DESIGN.md says "pinned against the reference's
module code, TF kernels restated".

Every function cites the reference lines it follows (paths under /root/reference).
All tensors are NHWC row-major; ``dtype`` float32 mirrors the reference, float64 gives the
error floor used to budget the TF32 tensor-core path.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12  # tf.nn.l2_normalize default epsilon


# ----------------------------------------------------------------------------- TF op semantics
def xw_plus_b(x, w, b):
    """tf.nn.xw_plus_b (util/cnn.py:116, util/empty_safe_conv.py:29)."""
    return x @ w + b


def l2_normalize(x, axis):
    """tf.nn.l2_normalize(x, dim): x * rsqrt(max(sum(x^2, dim), 1e-12))
    (models_clevr/nmn3_modules.py:107)."""
    ss = np.sum(x * x, axis=axis, keepdims=True)
    return x * (1.0 / np.sqrt(np.maximum(ss, np.asarray(EPS, x.dtype))))


def softmax_lastdim(x):
    """tf.nn.softmax over the last axis, max-subtracted (models_clevr/nmn3_modules.py:170-172)."""
    z = x - np.max(x, axis=-1, keepdims=True)
    e = np.exp(z)
    return e / np.sum(e, axis=-1, keepdims=True)


def conv2d_same(x, filt, bias):
    """tf.nn.conv2d(strides 1, padding 'SAME') + bias_add (util/cnn.py:29-32): cross-correlation,
    zero padding floor((k-1)/2) before / ceil after, filter [kh,kw,cin,cout]."""
    n, H, W, cin = x.shape
    kh, kw, _, cout = filt.shape
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    xp = np.zeros((n, H + kh - 1, W + kw - 1, cin), x.dtype)
    xp[:, pt:pt + H, pl:pl + W, :] = x
    # im2col + one GEMM (what a BLAS-backed conv does for a 1-channel input)
    cols = np.stack([xp[:, dy:dy + H, dx:dx + W, :] for dy in range(kh) for dx in range(kw)],
                    axis=3).reshape(n * H * W, kh * kw * cin)
    out = cols @ filt.reshape(kh * kw * cin, cout)
    return out.reshape(n, H, W, cout) + bias


def add_spatial_coordinate_map(feat):
    """models_vqa/nmn3_modules.py:11-31: append x=linspace(-1,1,W) (varies along W) and
    y=linspace(-1,1,H) (varies along H) channels."""
    n, H, W, _ = feat.shape
    xs = np.linspace(-1.0, 1.0, W).astype(feat.dtype)
    ys = np.linspace(-1.0, 1.0, H).astype(feat.dtype)
    x_map = np.broadcast_to(xs.reshape(1, 1, W, 1), (n, H, W, 1))
    y_map = np.broadcast_to(ys.reshape(1, H, 1, 1), (n, H, W, 1))
    return np.concatenate([feat, x_map, y_map], axis=3)


# ----------------------------------------------------------------------------- the modules
class OracleModules:
    """Restates class Modules of models_{clevr,shapes,vqa}/nmn3_modules.py.

    ``weights`` maps TF variable names relative to ``module_variables/`` (e.g.
    ``FindModule/conv_image/weights``) to arrays (n2nmn_b200.weights.variable_shapes).
    Inputs/outputs use the reference's shapes: att maps [n,H,W,1], answers [n,C],
    time_idx/batch_idx int [n].
    """

    def __init__(self, image_feat_grid, word_vecs, num_choices, weights, family='clevr',
                 dtype=np.float32):
        self.family = family
        self.dtype = dtype
        feat = np.asarray(image_feat_grid, dtype)
        if family == 'vqa':
            feat = add_spatial_coordinate_map(feat)  # models_vqa/nmn3_modules.py:35-36
        self.image_feat_grid = feat
        self.word_vecs = np.asarray(word_vecs, dtype)  # [T, N, Dt]
        self.num_choices = num_choices
        self.w = {k: np.asarray(v, dtype) for k, v in weights.items()}
        T, N, Dt = self.word_vecs.shape
        self.N_full = N
        # models_clevr/nmn3_modules.py:21-26
        self.word_vecs_flat = self.word_vecs.reshape(T * N, Dt)
        self.att_shape = [None] + list(feat.shape[1:3]) + [1]

    # -- gathers (models_clevr/nmn3_modules.py:49-57); materialised on purpose: tf.gather copies
    def _slice_image_feat_grid(self, batch_idx):
        return self.image_feat_grid[np.asarray(batch_idx, np.int64)]

    def _slice_word_vecs(self, time_idx, batch_idx):
        joint = np.asarray(time_idx, np.int64) * self.N_full + np.asarray(batch_idx, np.int64)
        return self.word_vecs_flat[joint]

    def _fc(self, scope, x):
        return xw_plus_b(x, self.w[scope + '/weights'], self.w[scope + '/biases'])

    def _conv1x1(self, scope, x):
        """empty_safe_1x1_conv (util/empty_safe_conv.py:8-32): reshape -> xw_plus_b -> reshape."""
        n, H, W, D = x.shape
        y = xw_plus_b(x.reshape(-1, D), self.w[scope + '/weights'], self.w[scope + '/biases'])
        return y.reshape(n, H, W, -1)

    # models_clevr/nmn3_modules.py:60-72
    def SceneModule(self, time_idx, batch_idx, pos_val=3):
        n = len(time_idx)
        return np.full([n] + self.att_shape[1:], pos_val, self.dtype)

    # models_clevr/nmn3_modules.py:74-111 (SHAPES :28-69, VQA :84-121)
    def FindModule(self, time_idx, batch_idx, scope='FindModule'):
        feat = self._slice_image_feat_grid(batch_idx)
        text = self._slice_word_vecs(time_idx, batch_idx)
        n = len(time_idx)
        mapped = self._conv1x1(scope + '/conv_image', feat)
        tmap = self._fc(scope + '/fc_text', text).reshape(n, 1, 1, -1)
        elt = l2_normalize(mapped * tmap, 3)
        return self._conv1x1(scope + '/conv_eltwise', elt)

    # models_clevr/nmn3_modules.py:113-132 — Find (FindModule weights) then And
    def FilterModule(self, input_0, time_idx, batch_idx):
        return self.AndModule(input_0, self.FindModule(time_idx, batch_idx), None, None)

    # models_clevr/nmn3_modules.py:134-183; VQA TransformModule models_vqa/nmn3_modules.py:123-171
    def FindSamePropertyModule(self, input_0, time_idx, batch_idx,
                               scope='FindSamePropertyModule'):
        feat = self._slice_image_feat_grid(batch_idx)
        text = self._slice_word_vecs(time_idx, batch_idx)
        n, H, W, _ = feat.shape
        mapped = self._conv1x1(scope + '/conv_image', feat)
        tmap = self._fc(scope + '/fc_text', text).reshape(n, 1, 1, -1)
        att_softmax = softmax_lastdim(np.asarray(input_0, self.dtype).reshape(n, H * W)) \
            .reshape(n, H, W, 1)
        att_feat = np.sum(feat * att_softmax, axis=(1, 2))
        amap = self._fc(scope + '/fc_att', att_feat).reshape(n, 1, 1, -1)
        elt = l2_normalize(mapped * tmap * amap, 3)
        return self._conv1x1(scope + '/conv_eltwise', elt)

    # models_clevr/nmn3_modules.py:185-216 (SHAPES :71-101, kernel 3)
    def TransformModule(self, input_0, time_idx, batch_idx, scope='TransformModule'):
        if self.family == 'vqa':
            return self.FindSamePropertyModule(input_0, time_idx, batch_idx, scope=scope)
        text = self._slice_word_vecs(time_idx, batch_idx)
        x = np.asarray(input_0, self.dtype)
        n = x.shape[0]
        maps = conv2d_same(x, self.w[scope + '/conv_maps/weights'],
                           self.w[scope + '/conv_maps/biases'])
        tmap = self._fc(scope + '/text_fc', text).reshape(n, 1, 1, -1)
        elt = l2_normalize(maps * tmap, 3)
        return self._conv1x1(scope + '/conv_eltwise', elt)

    # models_clevr/nmn3_modules.py:218-236 / :238-256
    def AndModule(self, input_0, input_1, time_idx=None, batch_idx=None):
        return np.minimum(input_0, input_1)

    def OrModule(self, input_0, input_1, time_idx=None, batch_idx=None):
        return np.maximum(input_0, input_1)

    # models_clevr/nmn3_modules.py:258-280; SHAPES AnswerModule models_shapes/nmn3_modules.py:123-150
    def ExistModule(self, input_0, time_idx=None, batch_idx=None, scope='ExistModule'):
        x = np.asarray(input_0, self.dtype)
        red = np.concatenate([x.min(axis=(1, 2)), x.mean(axis=(1, 2), dtype=self.dtype),
                              x.max(axis=(1, 2))], axis=1)
        return self._fc(scope + '/fc_scores', red)

    def AnswerModule(self, input_0, time_idx=None, batch_idx=None):
        return self.ExistModule(input_0, scope='AnswerModule')

    # models_clevr/nmn3_modules.py:282-304
    def CountModule(self, input_0, time_idx=None, batch_idx=None, scope='CountModule'):
        x = np.asarray(input_0, self.dtype)
        n = x.shape[0]
        cat = np.concatenate([x.reshape(n, -1), x.min(axis=(1, 2)), x.max(axis=(1, 2))], axis=1)
        return self._fc(scope + '/fc_scores', cat)

    # models_clevr/nmn3_modules.py:306-400 (EqualNum / MoreNum / LessNum: same math, own weights)
    def _compare(self, scope, input_0, input_1):
        parts = []
        for a in (input_0, input_1):
            x = np.asarray(a, self.dtype)
            parts += [x.reshape(x.shape[0], -1), x.min(axis=(1, 2)), x.max(axis=(1, 2))]
        return self._fc(scope + '/fc_scores', np.concatenate(parts, axis=1))

    def EqualNumModule(self, input_0, input_1, time_idx=None, batch_idx=None):
        return self._compare('EqualNumModule', input_0, input_1)

    def MoreNumModule(self, input_0, input_1, time_idx=None, batch_idx=None):
        return self._compare('MoreNumModule', input_0, input_1)

    def LessNumModule(self, input_0, input_1, time_idx=None, batch_idx=None):
        return self._compare('LessNumModule', input_0, input_1)

    def _pooled(self, feat, att):
        n, H, W, _ = feat.shape
        s = softmax_lastdim(np.asarray(att, self.dtype).reshape(n, H * W)).reshape(n, H, W, 1)
        return np.sum(feat * s, axis=(1, 2))

    # models_clevr/nmn3_modules.py:402-452
    def SamePropertyModule(self, input_0, input_1, time_idx, batch_idx,
                           scope='SamePropertyModule'):
        feat = self._slice_image_feat_grid(batch_idx)
        text = self._slice_word_vecs(time_idx, batch_idx)
        tmap = self._fc(scope + '/fc_text', text)
        a0 = self._fc(scope + '/fc_att_0', self._pooled(feat, input_0))
        a1 = self._fc(scope + '/fc_att_1', self._pooled(feat, input_1))
        elt = l2_normalize(a0 * tmap * a1, 1)
        return self._fc(scope + '/fc_eltwise', elt)

    # models_clevr/nmn3_modules.py:454-495; VQA models_vqa/nmn3_modules.py:193-240 (encoder_states
    # is None in the reference model, models_vqa/nmn3_model.py:61, so that branch is dead)
    def DescribeModule(self, input_0, time_idx, batch_idx, scope='DescribeModule'):
        feat = self._slice_image_feat_grid(batch_idx)
        text = self._slice_word_vecs(time_idx, batch_idx)
        tmap = self._fc(scope + '/fc_text', text)
        amap = self._fc(scope + '/fc_att', self._pooled(feat, input_0))
        elt = l2_normalize(tmap * amap, 1)
        return self._fc(scope + '/fc_eltwise', elt)


# ----------------------------------------------------------------------------- executors
_TOKEN_METHOD = {
    '_Scene': 'SceneModule', '_Find': 'FindModule', '_Filter': 'FilterModule',
    '_FindSameProperty': 'FindSamePropertyModule', '_Transform': 'TransformModule',
    '_And': 'AndModule', '_Or': 'OrModule', '_Exist': 'ExistModule', '_Count': 'CountModule',
    '_EqualNum': 'EqualNumModule', '_MoreNum': 'MoreNumModule', '_LessNum': 'LessNumModule',
    '_SameProperty': 'SamePropertyModule', '_Describe': 'DescribeModule',
    '_Answer': 'AnswerModule',
}
INVALID_EXPR = 'INVALID_EXPR'


def _call(modules, expr_module, inputs, t, b):
    fn = getattr(modules, _TOKEN_METHOD[expr_module])
    return fn(*inputs, np.asarray(t, np.int32), np.asarray(b, np.int32))


def run_sequential(modules, expr_list, return_att=False):
    """One question at a time, one node at a time (n=1 calls) — the stack executor of
    exp_shapes/visualize_shapes.ipynb cell 9 generalised to the recursion of
    models_clevr/nmn3_model.py:134-155. Invalid layouts give zeros(C) (:144-155)."""
    C = modules.num_choices
    scores = np.zeros((len(expr_list), C), modules.dtype)
    att_log = []

    def ev(e):
        ins = [ev(e[k]) for k in ('input_0', 'input_1') if k in e]
        out = _call(modules, e['module'], ins, [e['time_idx']], [e['batch_idx']])
        if return_att and e['output_type'] == 'att':
            att_log.append(((e['batch_idx'], e['time_idx']), out[0, :, :, 0].copy()))
        return out

    for i, e in enumerate(expr_list):
        if e['module'] != INVALID_EXPR:
            scores[i] = ev(e)[0]
    return (scores, dict(att_log)) if return_att else scores


def run_depth_batched(modules, expr_list, return_att=False):
    """TF-Fold-like dynamic batching: depth(node) = 1 + max depth of its
    attention inputs; all nodes of one module type at one depth across the batch go through ONE
    module call with leading dim n. This is the timed CPU baseline."""
    C = modules.num_choices
    nodes = []  # (depth, module, t, b, child ids, question or -1)

    def walk(e, q_root):
        kids = [walk(e[k], -1) for k in ('input_0', 'input_1') if k in e]
        depth = 1 + max([nodes[k][0] for k in kids], default=0)
        nodes.append((depth, e['module'], e['time_idx'], e['batch_idx'], kids, q_root))
        return len(nodes) - 1

    for q, e in enumerate(expr_list):
        if e['module'] != INVALID_EXPR:
            walk(e, q)
    scores = np.zeros((len(expr_list), C), modules.dtype)
    values = [None] * len(nodes)
    max_depth = max([n[0] for n in nodes], default=0)
    for d in range(1, max_depth + 1):
        by_type = {}
        for i, nd in enumerate(nodes):
            if nd[0] == d:
                by_type.setdefault(nd[1], []).append(i)
        for mod, ids in by_type.items():
            arity = len(nodes[ids[0]][4])
            ins = [np.stack([values[nodes[i][4][k]] for i in ids]) for k in range(arity)]
            out = _call(modules, mod, ins, [nodes[i][2] for i in ids], [nodes[i][3] for i in ids])
            for j, i in enumerate(ids):
                values[i] = out[j]
                if nodes[i][5] >= 0:
                    scores[nodes[i][5]] = out[j]
    if return_att:
        att = {(nd[3], nd[2]): values[i][:, :, 0] for i, nd in enumerate(nodes) if nd[5] < 0}
        return scores, att
    return scores

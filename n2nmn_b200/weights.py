"""Module-network variables: names, shapes and the reference's initialisers.

Variable names follow the TF scopes the reference creates, relative to
``neural_module_network/layout_execution/module_variables/``:
  models_clevr/nmn3_modules.py:17-18,91-92 (scope capture), :101-108 (Find), :164-181
  (FindSameProperty), :206-213 (Transform; note ``text_fc``), :278,302,334,366,398 (fc_scores),
  :429-450 (SameProperty), :478-493 (Describe).
Initialisers: xavier-uniform weights, zero biases (util/cnn.py:13-16,100-103,
util/empty_safe_conv.py:21-26).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from .config import FamilyConfig, FAMILIES

TF_SCOPE_PREFIX = 'neural_module_network/layout_execution/module_variables/'


def variable_shapes(family, H, W, D_feat, num_choices, text_dim=300):
    """Ordered {name: shape} of every ``weights``/``biases`` variable on the hot path.

    ``D_feat`` is the channel count of the feature grid handed to ``Modules`` (VQA adds its two
    coordinate channels on top, models_vqa/nmn3_modules.py:11-31).
    """
    cfg = family if isinstance(family, FamilyConfig) else FAMILIES[family]
    M, k, C, HW = cfg.map_dim, cfg.kernel_size, num_choices, H * W
    D = D_feat + (2 if cfg.add_coords else 0)
    v = OrderedDict()

    def layer(name, shape):
        v[name + '/weights'] = tuple(shape)
        v[name + '/biases'] = (shape[-1],)

    layer('FindModule/conv_image', (D, M))
    layer('FindModule/fc_text', (text_dim, M))
    layer('FindModule/conv_eltwise', (M, 1))
    if cfg.name == 'vqa':
        # VQA Transform is the attention-pooled variant (models_vqa/nmn3_modules.py:152-168)
        layer('TransformModule/conv_image', (D, M))
        layer('TransformModule/fc_text', (text_dim, M))
        layer('TransformModule/fc_att', (D, M))
        layer('TransformModule/conv_eltwise', (M, 1))
    else:
        layer('TransformModule/conv_maps', (k, k, 1, M))
        layer('TransformModule/text_fc', (text_dim, M))
        layer('TransformModule/conv_eltwise', (M, 1))
    if cfg.name == 'shapes':
        layer('AnswerModule/fc_scores', (3, C))
        return v
    if cfg.name == 'clevr':
        layer('FindSamePropertyModule/conv_image', (D, M))
        layer('FindSamePropertyModule/fc_text', (text_dim, M))
        layer('FindSamePropertyModule/fc_att', (D, M))
        layer('FindSamePropertyModule/conv_eltwise', (M, 1))
        layer('ExistModule/fc_scores', (3, C))
        layer('CountModule/fc_scores', (HW + 2, C))
        for nm in ('EqualNumModule', 'MoreNumModule', 'LessNumModule'):
            layer(nm + '/fc_scores', (2 * (HW + 2), C))
        layer('SamePropertyModule/fc_text', (text_dim, M))
        layer('SamePropertyModule/fc_att_0', (D, M))
        layer('SamePropertyModule/fc_att_1', (D, M))
        layer('SamePropertyModule/fc_eltwise', (M, C))
    layer('DescribeModule/fc_text', (text_dim, M))
    layer('DescribeModule/fc_att', (D, M))
    layer('DescribeModule/fc_eltwise', (M, C))
    return v


def _xavier_limit(shape):
    """tf.contrib.layers.xavier_initializer(): uniform, limit sqrt(6/(fan_in+fan_out));
    conv filters [kh,kw,cin,cout] use fan_in=kh*kw*cin, fan_out=kh*kw*cout."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = rf * shape[-2], rf * shape[-1]
    return float(np.sqrt(6.0 / (fan_in + fan_out)))


def init_weights(family, H, W, D_feat, num_choices, seed=0, bias_std=0.0, text_dim=300):
    """Seeded float32 initial values. ``bias_std=0`` reproduces the reference (zero biases);
    tests/bench use ``bias_std=0.1`` so that every bias path is exercised."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, shape in variable_shapes(family, H, W, D_feat, num_choices, text_dim).items():
        if name.endswith('/weights'):
            lim = _xavier_limit(shape)
            out[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        else:
            out[name] = (rng.standard_normal(shape) * bias_std).astype(np.float32)
    return out


def l2_reg(weights):
    """Sum of tf.nn.l2_loss (= 0.5*sum(w^2)) over every ``.../weights`` variable, biases
    excluded (models_clevr/nmn3_model.py:163-166)."""
    return float(sum(0.5 * np.sum(np.asarray(w, np.float64) ** 2)
                     for n, w in weights.items() if n.endswith('/weights')))


def init_seq2seq_weights(num_vocab_txt, embed_dim_txt, num_vocab_nmn, embed_dim_nmn, lstm_dim,
                         num_layers, seed=0):
    """Random weights of the layout generator under the reference's variable names relative to
    ``encoder_decoder/`` (models_clevr/nmn3_netgen_att.py:73-240): uniform, Xavier-like scale, the
    attention vector and the token projection a little hotter so that the decoded layouts vary
    across questions. For tests and benchmarks (no checkpoints are reachable offline)."""
    rng = np.random.RandomState(seed)
    L = lstm_dim
    w = OrderedDict()

    def u(*shape, a):
        return rng.uniform(-a, a, size=shape).astype(np.float32)
    w['encoder/embedding_mat'] = u(num_vocab_txt, embed_dim_txt, a=0.5)
    w['decoder/embedding_mat'] = u(num_vocab_nmn, embed_dim_nmn, a=0.5)
    w['decoder/go_embedding'] = u(1, embed_dim_nmn, a=0.5)
    for side, E in (('encoder', embed_dim_txt), ('decoder', embed_dim_nmn)):
        for l in range(num_layers):
            n_in = (E if l == 0 else L) + L
            p = '%s/lstm/multi_rnn_cell/cell_%d/basic_lstm_cell/' % (side, l)
            w[p + 'weights'] = u(n_in, 4 * L, a=(3.0 / n_in) ** 0.5)
            w[p + 'biases'] = u(4 * L, a=0.1)
    w['encoder/encoder_h_transform/weights'] = u(L, L, a=(3.0 / L) ** 0.5)
    w['encoder/encoder_h_transform/biases'] = u(L, a=0.1)
    w['decoder/att_prediction/weights'] = u(L, L, a=(3.0 / L) ** 0.5)
    w['decoder/att_prediction/biases'] = u(L, a=0.1)
    w['decoder/att_prediction/v'] = u(L, a=(3.0 / L) ** 0.5 * 4)
    w['decoder/token_prediction/weights'] = u(2 * L, num_vocab_nmn, a=(3.0 / L) ** 0.5 * 4)
    w['decoder/token_prediction/biases'] = u(num_vocab_nmn, a=0.1)
    return w

"""CPU ORACLE support (test infrastructure): an eager numpy stand-in for the ~30 TensorFlow 1.0
symbols that the reference's ``models_*/nmn3_modules.py``, ``util/cnn.py`` and
``util/empty_safe_conv.py`` touch, so that those files can be imported and EXECUTED UNMODIFIED
from /root/reference in this container (which has no TensorFlow) to generate golden vectors.

It is used by tests/golden/make_golden.py only. Nothing at test/bench run time needs
/root/reference or this shim (the GPU box has neither): the goldens are committed.

What is *not* the reference here: the arithmetic of each TF op. Those are restated from the
TF 1.0 API semantics listed in SURVEY.md App. A:
  xw_plus_b = x@W+b; l2_normalize(x,dim,eps=1e-12) = x*rsqrt(max(sum x^2, eps));
  softmax over the last axis; conv2d NHWC stride 1 'SAME' = zero-padded cross-correlation with
  filter [kh,kw,cin,cout]; gather on axis 0; reduce_* with keep_dims=False; concat(values, axis).
Variables are looked up by their full scope path in a user-supplied store (so the reference's
scope/reuse structure decides which weights each op sees); a missing name raises KeyError.
"""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as np


class _Shape:
    def __init__(self, dims):
        self._dims = list(dims)

    def as_list(self):
        return list(self._dims)


class T(np.ndarray):
    """ndarray with TF's static-shape accessors."""
    _static = None

    def get_shape(self):
        return _Shape(self._static if self._static is not None else self.shape)

    def set_shape(self, shape):
        self._static = shape.as_list() if isinstance(shape, _Shape) else list(shape)

    def __array_finalize__(self, obj):
        self._static = None


def _t(x, static=None, dtype=None):
    a = np.asarray(x, dtype=dtype).view(T)
    a._static = static
    return a


class _State:
    def __init__(self):
        self.scope = []
        self.store = {}
        self.created = []


STATE = _State()


@contextlib.contextmanager
def variable_scope(name_or_scope, reuse=None):
    saved = list(STATE.scope)
    if isinstance(name_or_scope, _Scope):
        STATE.scope = list(name_or_scope.path)  # re-entering a captured scope is absolute
    else:
        STATE.scope = STATE.scope + [name_or_scope]
    try:
        yield _Scope(STATE.scope)
    finally:
        STATE.scope = saved


class _Scope:
    def __init__(self, path):
        self.path = list(path)
        self.name = '/'.join(path)


def get_variable(name, shape=None, initializer=None, dtype=None):
    full = '/'.join(STATE.scope + [name])
    STATE.created.append(full)
    if full not in STATE.store:
        raise KeyError('tf1_shim: no value supplied for variable %r' % full)
    val = STATE.store[full]
    shp = [shape] if isinstance(shape, int) else list(shape)
    assert list(val.shape) == shp, (full, val.shape, shp)
    return _t(val)


def placeholder(dtype, shape=None):
    dims = [0 if d is None else d for d in shape]
    return _t(np.zeros(dims, dtype), static=list(shape))


def _axis(axis):
    return tuple(axis) if isinstance(axis, (list, tuple)) else axis


def conv2d(x, filter=None, strides=None, padding='SAME', **_):
    assert padding == 'SAME' and list(strides) == [1, 1, 1, 1]
    x = np.asarray(x)
    f = np.asarray(filter)
    n, H, W, cin = x.shape
    kh, kw, _, cout = f.shape
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    xp = np.zeros((n, H + kh - 1, W + kw - 1, cin), x.dtype)
    xp[:, pt:pt + H, pl:pl + W] = x
    out = np.zeros((n, H, W, cout), x.dtype)
    for y in range(H):
        for xx in range(W):
            patch = xp[:, y:y + kh, xx:xx + kw, :].reshape(n, kh * kw * cin)
            out[:, y, xx, :] = patch @ f.reshape(-1, cout)
    return _t(out)


def _l2_normalize(x, dim, epsilon=1e-12):
    x = np.asarray(x)
    ss = np.sum(np.square(x), axis=dim, keepdims=True)
    return _t(x * (1.0 / np.sqrt(np.maximum(ss, np.asarray(epsilon, x.dtype)))))


def _softmax(x):
    x = np.asarray(x)
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return _t(e / np.sum(e, axis=-1, keepdims=True))


def _gather_nd(params, indices):
    idx = np.asarray(indices)
    return _t(np.asarray(params)[tuple(idx[:, k] for k in range(idx.shape[1]))])


def install(store):
    """Register fake ``tensorflow`` / ``tensorflow_fold`` modules; ``store`` maps full variable
    paths to numpy arrays. Returns the fake tf module."""
    STATE.scope, STATE.store, STATE.created = [], dict(store), []
    tf = types.ModuleType('tensorflow')
    tf.float32, tf.int32 = np.float32, np.int32
    tf.convert_to_tensor = lambda v, *a, **k: v if isinstance(v, np.ndarray) else \
        [int(e) for e in v]
    tf.variable_scope = variable_scope
    tf.get_variable = get_variable
    tf.placeholder = placeholder
    tf.shape = lambda x: [int(d) for d in np.asarray(x).shape]
    tf.reshape = lambda x, shape: _t(np.reshape(np.asarray(x), [int(d) for d in shape]))
    tf.gather = lambda params, idx: _t(np.asarray(params)[np.asarray(idx, np.int64)])
    tf.gather_nd = _gather_nd
    tf.stack = lambda vals, axis=0: _t(np.stack([np.asarray(v) for v in vals], axis=axis))
    tf.ones = lambda shape, dtype=np.float32: _t(np.ones([int(d) for d in shape], dtype))
    tf.matmul = lambda a, b: _t(np.asarray(a) @ np.asarray(b))
    tf.minimum = lambda a, b: _t(np.minimum(a, b))
    tf.maximum = lambda a, b: _t(np.maximum(a, b))
    tf.reduce_sum = lambda x, axis=None: _t(np.sum(np.asarray(x), axis=_axis(axis)))
    tf.reduce_min = lambda x, axis=None: _t(np.min(np.asarray(x), axis=_axis(axis)))
    tf.reduce_max = lambda x, axis=None: _t(np.max(np.asarray(x), axis=_axis(axis)))
    tf.reduce_mean = lambda x, axis=None: _t(
        np.mean(np.asarray(x), axis=_axis(axis), dtype=np.asarray(x).dtype))
    tf.concat = lambda vals, axis: _t(np.concatenate([np.asarray(v) for v in vals], axis=axis))
    tf.tile = lambda x, mult: _t(np.tile(np.asarray(x), [int(m) for m in mult]))
    tf.linspace = lambda a, b, n: _t(np.linspace(a, b, int(n)).astype(np.float32))
    tf.stop_gradient = lambda x: x
    tf.add_to_collection = lambda *a, **k: None
    tf.constant_initializer = lambda *a, **k: None
    tf.RegisterGradient = lambda name: (lambda fn: fn)
    tf.device = lambda name: contextlib.nullcontext()

    class _Graph:
        def gradient_override_map(self, m):
            return contextlib.nullcontext()

    tf.get_default_graph = lambda: _Graph()
    tf.GraphKeys = types.SimpleNamespace(REGULARIZATION_LOSSES='regularization_losses')

    nn = types.ModuleType('tensorflow.nn')
    nn.xw_plus_b = lambda x, w, b: _t(np.asarray(x) @ np.asarray(w) + np.asarray(b))
    nn.l2_normalize = _l2_normalize
    nn.softmax = _softmax
    nn.conv2d = conv2d
    nn.bias_add = lambda x, b: _t(np.asarray(x) + np.asarray(b))
    nn.l2_loss = lambda w: 0.5 * float(np.sum(np.square(np.asarray(w, np.float64))))
    nn.relu = lambda x: _t(np.maximum(x, 0))
    tf.nn = nn

    layers = types.SimpleNamespace(xavier_initializer=lambda *a, **k: None,
                                   xavier_initializer_conv2d=lambda *a, **k: None)
    tf.contrib = types.SimpleNamespace(layers=layers)

    fold = types.ModuleType('tensorflow_fold')
    sys.modules['tensorflow'] = tf
    sys.modules['tensorflow.nn'] = nn
    sys.modules['tensorflow_fold'] = fold
    return tf


def uninstall():
    for k in ('tensorflow', 'tensorflow.nn', 'tensorflow_fold'):
        sys.modules.pop(k, None)
    for k in [m for m in sys.modules if m.split('.')[0] in
              ('models_clevr', 'models_shapes', 'models_vqa', 'util')]:
        sys.modules.pop(k, None)

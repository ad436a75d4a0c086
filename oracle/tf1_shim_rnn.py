"""CPU ORACLE support (test infrastructure): the additional TensorFlow 1.0 symbols that
``models_clevr/nmn3_netgen_att.py`` (AttentionSeq2Seq) touches, added on top of oracle/tf1_shim.py
so that the reference file can be imported and EXECUTED UNMODIFIED to generate golden vectors
(tests/golden/make_golden_seq2seq.py). Eager numpy; nothing at test/bench run time needs it.

Restated TF 1.0.0 semantics (the residual "parity unpinned" part, as for the module goldens):
  * BasicLSTMCell (contrib/rnn/python/ops/core_rnn_cell_impl.py): variables
    ``basic_lstm_cell/{weights [in+units, 4*units], biases [4*units]}``; gates split in the order
    i, j, f, o; ``c' = c*sigmoid(f + 1.0) + sigmoid(i)*tanh(j)``, ``h' = tanh(c')*sigmoid(o)``.
  * MultiRNNCell: layer l runs in scope ``multi_rnn_cell/cell_<l>``; state is a tuple of (c, h).
  * dynamic_rnn(time_major=True, sequence_length): for t >= length the output row is zero and the
    state row is carried through unchanged.
  * raw_rnn: loop_fn(0, None, None, None) supplies the first input / state; each iteration runs
    the cell, then loop_fn(time+1, output, state, loop_state); stops when all elements finish.
  * tensordot(axes=1), embedding_lookup = gather on axis 0, softmax(dim), reduce_*(keep_dims),
    where / argmax (first maximum) / logical ops / TensorArray.write/stack.
"""
from __future__ import annotations

import types

import numpy as np

from . import tf1_shim as base
from .tf1_shim import _t, variable_scope, get_variable


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class BasicLSTMCell:
    def __init__(self, num_units, forget_bias=1.0, state_is_tuple=True):
        assert state_is_tuple
        self.num_units, self.forget_bias = num_units, np.float32(forget_bias)
        self.output_size = num_units

    def __call__(self, inputs, state):
        c, h = state
        x = np.concatenate([np.asarray(inputs), np.asarray(h)], axis=1)
        with variable_scope('basic_lstm_cell'):
            w = get_variable('weights', [x.shape[1], 4 * self.num_units])
            b = get_variable('biases', [4 * self.num_units])
        concat = (x @ np.asarray(w) + np.asarray(b)).astype(np.float32)
        i, j, f, o = np.split(concat, 4, axis=1)
        new_c = (np.asarray(c) * _sigmoid(f + self.forget_bias) + _sigmoid(i) * np.tanh(j))
        new_h = np.tanh(new_c) * _sigmoid(o)
        return _t(new_h.astype(np.float32)), (_t(new_c.astype(np.float32)), _t(new_h.astype(np.float32)))

    def zero_state(self, n):
        z = np.zeros((n, self.num_units), np.float32)
        return (_t(z), _t(z.copy()))


class DropoutWrapper:
    def __init__(self, cell, output_keep_prob=1.0):
        raise NotImplementedError('goldens are generated without dropout (eval configuration)')


class MultiRNNCell:
    def __init__(self, cells, state_is_tuple=True):
        assert state_is_tuple
        self.cells = list(cells)
        self.output_size = self.cells[-1].output_size

    def __call__(self, inputs, state):
        cur, new_states = inputs, []
        with variable_scope('multi_rnn_cell'):
            for l, cell in enumerate(self.cells):
                with variable_scope('cell_%d' % l):
                    cur, ns = cell(cur, state[l])
                new_states.append(ns)
        return cur, tuple(new_states)

    def zero_state(self, n):
        return tuple(c.zero_state(n) for c in self.cells)


def dynamic_rnn(cell, inputs, sequence_length=None, dtype=None, time_major=False, scope=None):
    assert time_major
    x = np.asarray(inputs)
    T, N = x.shape[0], x.shape[1]
    lens = np.asarray(sequence_length)
    state = cell.zero_state(N)
    outs = []
    with variable_scope(scope or 'rnn'):
        for t in range(T):
            out, new_state = cell(_t(x[t]), state)
            live = (t < lens)[:, None]
            outs.append(np.where(live, np.asarray(out), 0.0).astype(np.float32))
            state = tuple((_t(np.where(live, np.asarray(nc), np.asarray(oc))),
                           _t(np.where(live, np.asarray(nh), np.asarray(oh))))
                          for (nc, nh), (oc, oh) in zip(new_state, state))
    return _t(np.stack(outs)), state


class TensorArray:
    def __init__(self, dtype=None, size=0, infer_shape=True):
        self.items = [None] * int(size)

    def write(self, index, value):
        self.items[int(index)] = np.asarray(value)
        return self

    def stack(self):
        return _t(np.stack(self.items))


def raw_rnn(cell, loop_fn, scope=None):
    with variable_scope(scope or 'rnn'):
        time = 0
        finished, nxt, state, _, loop_state = loop_fn(time, None, None, None)
        out = None
        while not bool(np.all(finished)):
            out, cell_state = cell(nxt, state)
            time += 1
            finished, nxt, state, _, loop_state = loop_fn(time, out, cell_state, loop_state)
    return None, state, loop_state


def _softmax(x, dim=-1):
    x = np.asarray(x)
    e = np.exp(x - np.max(x, axis=dim, keepdims=True))
    return _t(e / np.sum(e, axis=dim, keepdims=True))


def _reduce(fn):
    def r(x, axis=None, keep_dims=False):
        ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
        return _t(fn(np.asarray(x), axis=ax, keepdims=keep_dims))
    return r


_UNIFORMS = []


def set_sampling_uniforms(u):
    """Uniform numbers in [0, 1) that the shim's `tf.multinomial` consumes, one row [N] per call
    (= per decoding step of nmn3_netgen_att.py:236-238)."""
    _UNIFORMS[:] = [np.asarray(r, np.float64) for r in u]


def multinomial(logits, num_samples):
    """`tf.multinomial(logits, 1)` as inverse-CDF sampling from softmax(logits) with the next row
    of set_sampling_uniforms(): the first class whose cumulative probability exceeds u (TF's own
    generator is not reproducible outside TF; the distribution is the same)."""
    assert num_samples == 1
    z = np.asarray(logits, np.float64)
    q = np.exp(z - z.max(axis=1, keepdims=True))
    cdf = np.cumsum(q, axis=1)
    u = _UNIFORMS.pop(0)
    tok = (cdf <= (u * cdf[:, -1])[:, None]).sum(axis=1)
    return _t(np.minimum(tok, z.shape[1] - 1).astype(np.int64)[:, None])


def install_rnn(tf):
    """Adds the seq2seq symbols to the fake module returned by tf1_shim.install()."""
    tf.multinomial = multinomial
    tf.newaxis = None
    tf.convert_to_tensor = lambda v, dtype=None, **k: (
        v if isinstance(v, np.ndarray) else
        ([int(e) for e in v] if all(np.isscalar(e) for e in v) else np.asarray(v, dtype)))
    tf.greater_equal = lambda a, b: _t(np.greater_equal(a, b))
    tf.less = lambda a, b: _t(np.less(a, b))
    tf.equal = lambda a, b: _t(np.equal(a, b))
    tf.logical_or = lambda a, b: _t(np.logical_or(a, b))
    tf.logical_and = lambda a, b: _t(np.logical_and(a, b))
    tf.tensordot = lambda a, b, axes=1: _t(np.tensordot(np.asarray(a), np.asarray(b), axes=axes))
    tf.reduce_all = _reduce(np.all)
    tf.reduce_any = _reduce(np.any)
    tf.reduce_sum = _reduce(np.sum)
    tf.reduce_min = _reduce(np.min)
    tf.reduce_max = _reduce(np.max)
    tf.cast = lambda x, dtype: _t(np.asarray(x).astype(dtype))
    tf.range = lambda *a, **k: _t(np.arange(*a, dtype=k.get('dtype', np.int32)))
    tf.tanh = lambda x: _t(np.tanh(np.asarray(x)))
    tf.log = lambda x: _t(np.log(np.asarray(x)))
    tf.where = lambda c, a, b: _t(np.where(np.asarray(c), np.asarray(a), np.asarray(b)))
    tf.ones_like = lambda x: _t(np.ones_like(np.asarray(x)))
    tf.zeros = lambda shape, dtype=np.float32: _t(np.zeros([int(d) for d in shape], dtype))
    tf.argmax = lambda x, axis: _t(np.argmax(np.asarray(x), axis=axis))
    tf.TensorArray = TensorArray
    tf.nn.embedding_lookup = lambda params, ids: _t(np.asarray(params)[np.asarray(ids, np.int64)])
    tf.nn.softmax = _softmax
    tf.nn.dynamic_rnn = dynamic_rnn
    tf.nn.raw_rnn = raw_rnn
    rnn = types.SimpleNamespace(BasicLSTMCell=BasicLSTMCell, DropoutWrapper=DropoutWrapper,
                                MultiRNNCell=MultiRNNCell)
    tf.contrib.rnn = rnn
    return tf

"""Layout assembler: Reverse-Polish layout tokens -> per-question expression trees.

Drop-in for ``models_clevr/nmn3_assembler.py`` (and its SHAPES / VQA siblings
``models_shapes/nmn3_assembler.py``, ``models_vqa/nmn3_assembler.py``): same constructor,
attributes (``module_names, EOS_idx, name2idx_dict, num_vocab_nmn, P, W, b``), methods
(``module_list2tokens``, ``assemble``) and the same expression dictionaries
(``module, output_type, time_idx, batch_idx, input_0, input_1`` /
``INVALID_EXPR`` with ``expr_str`` and ``error``).

This file is host logic only. The throughput path does not build Python dicts at all: it hands
the raw ``[T, N]`` token matrix to ``n2nmn_compile_schedule`` (csrc/schedule.cpp), which applies
the same stack discipline in C++; ``tests/test_schedule.py`` checks both agree.
"""
from __future__ import annotations

import numpy as np

INVALID_EXPR = 'INVALID_EXPR'

# (number of attention inputs, output kind) for every layout token of the three vocabularies.
# Reference tables: models_clevr/nmn3_assembler.py:9-40, models_shapes/nmn3_assembler.py:9-18,
# models_vqa/nmn3_assembler.py:9-18. Names never clash across families, so one table serves all.
MODULE_SIGNATURE = {
    '_Scene': (0, 'att'),
    '_Find': (0, 'att'),
    '_Filter': (1, 'att'),
    '_FindSameProperty': (1, 'att'),
    '_Transform': (1, 'att'),
    '_And': (2, 'att'),
    '_Or': (2, 'att'),
    '_Exist': (1, 'ans'),
    '_Count': (1, 'ans'),
    '_EqualNum': (2, 'ans'),
    '_MoreNum': (2, 'ans'),
    '_LessNum': (2, 'ans'),
    '_SameProperty': (2, 'ans'),
    '_Describe': (1, 'ans'),
    '_Answer': (1, 'ans'),
}

_module_input_num = {k: v[0] for k, v in MODULE_SIGNATURE.items()}
_module_output_type = {k: v[1] for k, v in MODULE_SIGNATURE.items()}


def build_validity_mats(module_names):
    """Decoder-side validity automaton (reference: models_clevr/nmn3_assembler.py:50-119).

    State x = [#att on stack, #ans on stack, T_remain]. Token s is admissible iff
    ``x @ W[:, s, c] - b[s, c] >= 0`` for all four constraints c; emitting s moves the state by
    ``P[s]``. Only the seq2seq decoder (off the hot path) and the synthetic layout sampler consume
    these; they are kept for API parity.
    """
    V = len(module_names)
    is_mod = np.array([s != '<eos>' for s in module_names])
    n_in = np.array([_module_input_num[s] if s != '<eos>' else 0 for s in module_names], np.int64)
    is_ans = np.array([s != '<eos>' and _module_output_type[s] == 'ans' for s in module_names])
    is_att = is_mod & ~is_ans

    P = np.zeros((V, 3), np.int32)
    P[:, 0] = is_att.astype(np.int64) - n_in
    P[:, 1] = is_ans
    P[:, 2] = -1

    absorb = n_in - is_att.astype(np.int64)        # net attentions consumed by a token
    max_absorb_att = int(np.max(absorb * is_att))  # best non-answer consumer
    max_absorb_ans = int(np.max(absorb * is_ans))  # best answer consumer

    W = np.zeros((3, V, 4), np.int32)
    b = np.zeros((V, 4), np.int32)
    # c0: enough attentions on the stack            #att >= n_in
    W[0, is_mod, 0] = 1
    b[is_mod, 0] = n_in[is_mod]
    # c1: answer modules leave nothing behind       -#att >= -n_in
    W[0, is_ans, 1] = -1
    b[is_ans, 1] = -n_in[is_ans]
    #     attention modules need room for ans + eos  T_remain >= 3
    W[2, is_att, 1] = 1
    b[is_att, 1] = 3
    # c2: nothing may follow an answer but <eos>    -#ans >= 0
    W[1, is_mod, 2] = -1
    # c3: attention modules must leave a stack that can still be consumed in time
    #     -#att + max_absorb_att*T_remain >= 3*max_absorb_att - max_absorb_ans - absorb[s]
    W[0, is_att, 3] = -1
    W[2, is_att, 3] = max_absorb_att
    b[is_att, 3] = 3 * max_absorb_att - max_absorb_ans - absorb[is_att]
    # <eos>: only once an answer exists             #ans >= 1
    eos = ~is_mod
    W[1, eos, 0] = 1
    b[eos, 0] = 1
    return P, W, b


class Assembler:
    def __init__(self, module_vocab_file):
        with open(module_vocab_file) as f:
            self.module_names = [line.strip() for line in f.readlines()]
        self.EOS_idx = self.module_names.index('<eos>')
        self.name2idx_dict = {name: i for i, name in enumerate(self.module_names)}
        self.num_vocab_nmn = len(self.module_names)
        self.P, self.W, self.b = build_validity_mats(self.module_names)

    @classmethod
    def from_names(cls, module_names):
        """Build from an in-memory vocabulary (no file); not in the reference."""
        self = cls.__new__(cls)
        self.module_names = list(module_names)
        self.EOS_idx = self.module_names.index('<eos>')
        self.name2idx_dict = {name: i for i, name in enumerate(self.module_names)}
        self.num_vocab_nmn = len(self.module_names)
        self.P, self.W, self.b = build_validity_mats(self.module_names)
        return self

    # reference: models_clevr/nmn3_assembler.py:137-143
    def module_list2tokens(self, module_list, T=None):
        layout_tokens = [self.name2idx_dict[name] for name in module_list]
        if T is not None:
            if len(module_list) >= T:
                raise ValueError('Not enough time steps to add <eos>')
            layout_tokens += [self.EOS_idx] * (T - len(module_list))
        return layout_tokens

    def _layout_tokens2str(self, layout_tokens):
        return ' '.join(self.module_names[idx] for idx in layout_tokens)

    def _invalid_expr(self, layout_tokens, error_str):
        return {'module': INVALID_EXPR,
                'expr_str': self._layout_tokens2str(layout_tokens),
                'error': error_str}

    # reference: models_clevr/nmn3_assembler.py:153-212
    def _assemble_layout_tokens(self, layout_tokens, batch_idx):
        tokens = [int(t) for t in layout_tokens]
        if self.EOS_idx not in tokens:
            return self._invalid_expr(tokens, 'cannot find <eos>')
        stack = []
        for t, tok in enumerate(tokens):
            if tok == self.EOS_idx:
                break
            name = self.module_names[tok]
            arity, out_type = MODULE_SIGNATURE[name]
            if len(stack) < arity:
                return self._invalid_expr(tokens, 'not enough input for ' + name)
            node = {'module': name, 'output_type': out_type,
                    'time_idx': t, 'batch_idx': batch_idx}
            # operands come off the stack right-to-left: the last popped one is input_0
            for slot in reversed(range(arity)):
                operand = stack.pop()
                if operand['output_type'] != 'att':
                    return self._invalid_expr(tokens, 'input incompatible for ' + name)
                node['input_%d' % slot] = operand
            stack.append(node)
        if len(stack) != 1:
            return self._invalid_expr(
                tokens, 'final stack size not equal to 1 (%d remains)' % len(stack))
        if stack[0]['output_type'] != 'ans':
            return self._invalid_expr(tokens, 'result type must be ans, not att')
        return stack[0]

    # reference: models_clevr/nmn3_assembler.py:214-222
    def assemble(self, layout_tokens_batch):
        layout_tokens_batch = np.asarray(layout_tokens_batch)
        _, N = layout_tokens_batch.shape
        expr_list = [self._assemble_layout_tokens(layout_tokens_batch[:, n], n)
                     for n in range(N)]
        expr_validity = np.array([e['module'] != INVALID_EXPR for e in expr_list], bool)
        return expr_list, expr_validity

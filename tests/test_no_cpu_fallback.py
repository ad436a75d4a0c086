"""CPU: the product has no CPU path. Without a CUDA device every entry point that would compute
raises (Python) / returns a negative status with a message (C ABI) instead of falling back, and
nothing under n2nmn_b200/ imports the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from n2nmn_b200 import _lib, synth
from n2nmn_b200.assembler import Assembler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason='a CUDA device is present')


@no_gpu
def test_modules_and_executor_raise_without_a_gpu():
    from n2nmn_b200.executor import LayoutExecutor
    from n2nmn_b200.modules import ModulesCLEVR, ModulesSHAPES, ModulesVQA
    f, w = synth.make_inputs(2, 10, 15, 512, 5, seed=1)
    for cls, args in ((ModulesCLEVR, (f, w, 28)), (ModulesSHAPES, (f, w, 2)),
                      (ModulesVQA, (f, w, None, 3001))):
        with pytest.raises(_lib.N2NMNError):
            cls(*args)
    asm = Assembler(synth.vocab_file('clevr'))
    with pytest.raises(_lib.N2NMNError):
        LayoutExecutor('clevr', f, w, 28, asm)


@no_gpu
def test_seq2seq_raises_without_a_gpu():
    from n2nmn_b200.seq2seq import AttentionSeq2Seq
    asm = Assembler(synth.vocab_file('clevr'))
    for dev in ('cpu', None):
        with pytest.raises(_lib.N2NMNError):
            AttentionSeq2Seq(None, None, 5, 10, 8, asm.num_vocab_nmn, 8, 16, 1, asm, T_encoder=4,
                             max_batch=2, device=dev)


@no_gpu
def test_c_abi_create_fails_with_a_message_without_a_gpu():
    L = _lib.lib()
    cfg = _lib.Seq2SeqConfig(_lib.ABI_VERSION, 10, 8, 15, 8, 16, 1, 4, 5, 2, 0, 0)
    h = C.c_void_p()
    rc = L.n2nmn_seq2seq_create(C.byref(cfg), C.byref(h))
    assert rc < 0 and not h.value
    assert L.n2nmn_last_error()          # a message, not a silent failure
    # NULL arguments are rejected before any device work
    assert L.n2nmn_seq2seq_create(None, C.byref(h)) < 0
    assert L.n2nmn_seq2seq_set_sampling(None, None) < 0


def test_product_package_never_imports_the_oracle():
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|importlib.*oracle', re.M)
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, 'n2nmn_b200')):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(d, fn)).read()
                if pat.search(src):
                    bad.append(os.path.join(d, fn))
    assert not bad, bad
    # bench.py: only the reference arm / cpu_baseline legs may touch oracle/
    src = open(os.path.join(ROOT, 'bench.py')).read()
    for m in re.finditer(r'^\s*(from|import)\s+oracle\b.*$', src, re.M):
        head = src[:m.start()]
        owner = re.findall(r'^(?:def|class) (\w+)', head, re.M)[-1]     # enclosing top-level scope
        assert owner in ('CpuPort', 'seq2seq_cpu_port', 'run_reference_arm'), (owner, m.group(0))
    assert np is not None

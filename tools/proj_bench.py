#!/usr/bin/env python
"""Kernel-only timing of the projection kernel (per-launch CUDA events) on a Find-only batch.
Used with N2NMN_LIB=<experiment build> to attribute its time (tools/build_variants.py + tools/gpu_epilogue_attrib.sh)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from n2nmn_b200 import synth, weights as wts
from n2nmn_b200.assembler import Assembler
from n2nmn_b200.executor import LayoutExecutor

B = int(os.environ.get('PB_BATCH', 64))
H, W, D, T, C = 10, 15, 512, 20, 28
asm = Assembler(synth.vocab_file('clevr'))
layout = os.environ.get('PB_LAYOUT', 'find')
if layout == 'find':
    toks = synth.tokens_from_layouts(asm, [['_Find', '_Exist']] * B, T)
else:
    toks = synth.expert_mix_tokens(asm, B, T)
weights = wts.init_weights('clevr', H, W, D, C, seed=0, bias_std=0.1)
P = 10
feats, wvs = [], []
for i in range(P):
    f, w = synth.make_inputs(B, H, W, D, T, seed=1234 + i)
    feats.append(torch.from_numpy(f).cuda()); wvs.append(torch.from_numpy(w).cuda())
ex = LayoutExecutor('clevr', feats[0], wvs[0], C, asm, weights=weights, max_batch=B, max_T=T)
for i in range(10):
    ex.forward_device(feats[i % P], wvs[i % P], toks)
torch.cuda.synchronize()
ex.set_profiling(True)
acc = {}
for i in range(60):
    ex.forward_device(feats[i % P], wvs[i % P], toks)
    for name, us in ex.launch_times():
        acc.setdefault(name, []).append(us)
print(os.environ.get('N2NMN_LIB', 'default'), layout, 'B=%d' % B,
      {k: (round(float(np.median(v)), 2), round(float(np.mean(v)), 2)) for k, v in acc.items()}, '(median, mean us)')

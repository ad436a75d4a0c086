#!/usr/bin/env python
"""Six train steps on one CLEVR batch (profiling target of tools/gpu_train_prof.sh)."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from n2nmn_b200 import synth, weights as wts
from n2nmn_b200.assembler import Assembler
from n2nmn_b200.executor import LayoutExecutor
from n2nmn_b200.trainer import ModuleNetTrainer
B, H, W, D, T, C = 64, 10, 15, 512, 10, 28
asm = Assembler(synth.vocab_file('clevr'))
weights = wts.init_weights('clevr', H, W, D, C, seed=0, bias_std=0.1)
f, w = synth.make_inputs(B, H, W, D, T, seed=1)
f, w = torch.from_numpy(f).cuda(), torch.from_numpy(w).cuda()
ex = LayoutExecutor('clevr', f, w, C, asm, weights=weights, max_batch=B, max_T=T)
tr = ModuleNetTrainer(ex)
tok = synth.expert_mix_tokens(asm, B, T)
lab = np.random.RandomState(0).randint(0, C, size=B)
for i in range(6):
    tr.train_step(f, w, tok, lab)
torch.cuda.synchronize()

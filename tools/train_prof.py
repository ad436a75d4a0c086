"""Runs a few policy-search train steps (CLEVR B=64, T=10) for the profiler / a timing print."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from n2nmn_b200 import synth, weights as wts
from n2nmn_b200.assembler import Assembler
from n2nmn_b200.executor import LayoutExecutor
from n2nmn_b200.trainer import ModuleNetTrainer

B, T, C = 64, 10, 28
asm = Assembler(synth.vocab_file('clevr'))
feat, wv = synth.make_inputs(B, 10, 15, 512, T, seed=1)
W = wts.init_weights('clevr', 10, 15, 512, C, seed=0, bias_std=0.1)
ex = LayoutExecutor('clevr', torch.from_numpy(feat).cuda(), torch.from_numpy(wv).cuda(), C, asm,
                    weights=W, max_batch=B, max_T=T)
tr = ModuleNetTrainer(ex)
tok = synth.expert_mix_tokens(asm, B, T)
lab = np.random.RandomState(3).randint(0, C, size=B)
lsp = torch.full((B,), -2.0, device='cuda')
f, w = torch.from_numpy(feat).cuda(), torch.from_numpy(wv).cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(5):
    tr.train_step(f, w, tok, lab, log_seq_prob=lsp, sync=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    tr.train_step(f, w, tok, lab, log_seq_prob=lsp, sync=False)
e1.record(); torch.cuda.synchronize()
print('train step: %.3f ms' % (e0.elapsed_time(e1) / n))

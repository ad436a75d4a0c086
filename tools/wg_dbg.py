import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from n2nmn_b200 import synth, weights as wts
from n2nmn_b200.assembler import Assembler
from n2nmn_b200.executor import LayoutExecutor
from n2nmn_b200.trainer import ModuleNetTrainer
B, T, C = 12, 10, 28
asm = Assembler(synth.vocab_file('clevr'))
feat, wv = synth.make_inputs(B, 10, 15, 512, T, seed=1)
W = wts.init_weights('clevr', 10, 15, 512, C, seed=0, bias_std=0.1)
ex = LayoutExecutor('clevr', torch.from_numpy(feat).cuda(), torch.from_numpy(wv).cuda(), C, asm, weights=W, max_batch=B, max_T=T)
tr = ModuleNetTrainer(ex)
tok = synth.expert_mix_tokens(asm, B, T)
lab = np.arange(B) % C
f, w = torch.from_numpy(feat).cuda(), torch.from_numpy(wv).cuda()
def grads(env):
    if env: os.environ['N2NMN_WGRAD_MMA_SYNC'] = '1'
    else: os.environ.pop('N2NMN_WGRAD_MMA_SYNC', None)
    tr.forward_backward(f, w, tok, lab)
    torch.cuda.synchronize()
    return {n: g.cpu().numpy().copy() for n, g in tr.grads().items()}
g_ref = grads(True)
g_new = grads(False)
for n in g_ref:
    if 'conv_image' in n or 'fc_att' in n:
        a, b = g_new[n], g_ref[n]
        print('%-48s ref|max| %.3e new|max| %.3e  maxdiff %.3e  corr %.4f' % (
            n, np.abs(b).max(), np.abs(a).max(), np.abs(a - b).max(),
            float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))))
n = 'FindModule/conv_image/weights'
a, b = g_new[n].reshape(512, -1), g_ref[n].reshape(512, -1)
print('rows nonzero new', int((np.abs(a).sum(1) > 0).sum()), 'cols nonzero new', int((np.abs(a).sum(0) > 0).sum()))
print('new[0,:6]', a[0, :6], '\nref[0,:6]', b[0, :6])
# is new a scaled / permuted version?
print('ratio of norms', np.linalg.norm(a) / np.linalg.norm(b))

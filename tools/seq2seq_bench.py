"""Device time of one seq2seq forward (CLEVR sizes) with CUDA events; prints questions/s."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from n2nmn_b200 import synth
from n2nmn_b200.assembler import Assembler
from n2nmn_b200.seq2seq import AttentionSeq2Seq
from n2nmn_b200.weights import init_seq2seq_weights

N, T_enc, T_dec, L, layers, V_txt, E = 64, 45, 20, 512, 2, 90, 300
asm = Assembler(synth.vocab_file('clevr'))
rng = np.random.RandomState(0)
w = init_seq2seq_weights(V_txt, E, asm.num_vocab_nmn, E, L, layers)
s = AttentionSeq2Seq(None, None, T_dec, V_txt, E, asm.num_vocab_nmn, E, L, layers, asm,
                     T_encoder=T_enc, max_batch=N, weights=w, device='cuda:0',
                     precision='tf32' if '--tf32' in sys.argv else 'fp32')
seq = torch.from_numpy(rng.randint(0, V_txt, size=(T_enc, N)).astype(np.int32)).cuda()
lens = torch.from_numpy(rng.randint(5, T_enc + 1, size=N).astype(np.int32)).cuda()
for _ in range(5):
    s.forward(seq, lens)
torch.cuda.synchronize()
n0 = s.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = 50
e0.record()
for _ in range(R):
    s.forward(seq, lens)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
print('seq2seq forward N=%d T_enc=%d T_dec=%d L=%d: %.3f ms/batch, %.0f questions/s, %d launches/batch'
      % (N, T_enc, T_dec, L, ms, N / ms * 1e3, (s.launch_count() - n0) // R))

"""GPU: the training-loop glue (n2nmn_b200/train.py) end to end on synthetic
data: frozen seq2seq (teacher forced) -> word vectors -> module-network train steps -> snapshot in
the TF checkpoint format -> snapshot restores into a fresh executor and reproduces its scores."""
import numpy as np
import pytest
import torch

from n2nmn_b200 import checkpoint, synth, weights as wts
from n2nmn_b200.assembler import Assembler
from n2nmn_b200.weights import init_seq2seq_weights

pytestmark = pytest.mark.gpu


def test_training_loop_with_frozen_layout_generator(tmp_path):
    from n2nmn_b200.executor import LayoutExecutor
    from n2nmn_b200.seq2seq import AttentionSeq2Seq
    from n2nmn_b200.train import run_training
    from n2nmn_b200.trainer import ModuleNetTrainer
    N, H, W, D, T_enc, T_dec, C = 16, 10, 15, 512, 12, 10, 28
    V_txt, E, L = 40, 300, 64
    asm = Assembler(synth.vocab_file('clevr'))
    rng = np.random.RandomState(0)
    s2s = AttentionSeq2Seq(None, None, T_dec, V_txt, E, asm.num_vocab_nmn, E, L, 1, asm,
                           T_encoder=T_enc, max_batch=N, device='cuda:0',
                           weights=init_seq2seq_weights(V_txt, E, asm.num_vocab_nmn, E, L, 1, seed=1))
    feat, wv0 = synth.make_inputs(N, H, W, D, T_dec, seed=3)
    Wm = wts.init_weights('clevr', H, W, D, C, seed=0, bias_std=0.1)
    ex = LayoutExecutor('clevr', torch.from_numpy(feat).cuda(), torch.from_numpy(wv0).cuda(), C, asm,
                        weights=Wm, max_batch=N, max_T=T_dec)
    tr = ModuleNetTrainer(ex, lr=1e-3)
    tokens = synth.expert_mix_tokens(asm, N, T_dec)
    batch = {'image_feat_batch': feat, 'gt_layout_batch': tokens,
             'input_seq_batch': rng.randint(0, V_txt, size=(T_enc, N)).astype(np.int32),
             'seq_length_batch': rng.randint(3, T_enc + 1, size=N).astype(np.int32),
             'answer_label_batch': (np.arange(N) * 3) % C}

    def word_vecs_fn(b, toks):      # the frozen generator, teacher forced (train_clevr_gt_layout.py)
        out = s2s.forward(b['input_seq_batch'], b['seq_length_batch'], True, toks)
        assert np.array_equal(out[0].cpu().numpy(), toks)
        return out[3]

    logs = []
    hist = run_training(tr, (batch for _ in range(1000)), asm, max_iter=25, word_vecs_fn=word_vecs_fn,
                        snapshot_dir=str(tmp_path), snapshot_interval=10, log_interval=5,
                        log=logs.append)
    assert len(hist) == 25 and hist[-1]['loss'] < 0.7 * hist[0]['loss']
    assert all(h['validity'] == 1.0 for h in hist)
    assert any('snapshot saved' in l for l in logs) and sum('iter = ' in l for l in logs) == 5
    # the last snapshot restores into a fresh executor and gives the trained model's scores
    got, ignored = checkpoint.import_module_weights(str(tmp_path / '00000025'))
    assert not ignored
    wv = word_vecs_fn(batch, tokens)
    ex2 = LayoutExecutor('clevr', torch.from_numpy(feat).cuda(), wv, C, asm, weights=got,
                         max_batch=N, max_T=T_dec)
    s_new, _ = ex2.forward_device(torch.from_numpy(feat).cuda(), wv, tokens)
    s_old, _ = ex.forward_device(torch.from_numpy(feat).cuda(), wv, tokens)
    torch.cuda.synchronize()
    np.testing.assert_allclose(s_new.cpu().numpy(), s_old.cpu().numpy(), atol=1e-5)
    for name in ('00000010', '00000020', '00000025'):
        assert (tmp_path / (name + '.index')).exists()

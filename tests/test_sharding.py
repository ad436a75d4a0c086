"""CPU, world_size 2 over gloo: the data-parallel scheme of bench.py / ModuleNetTrainer —
questions sharded by contiguous batch slices, no collective in eval, ONE summed all-reduce of the
flat gradient (+ loss) in training — reproduces the single-process result. The oracle stands in
for the CUDA compute (no GPU here); the collective logic is the thing under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from n2nmn_b200 import synth, weights as wts
from n2nmn_b200.assembler import Assembler
from oracle import nmn_oracle_torch as ot

N, H, W, D, T, C = 8, 10, 15, 64, 8, 28


def _case():
    feat, word_vecs = synth.make_inputs(N, H, W, D, T, seed=77)
    Wt = wts.init_weights('clevr', H, W, D, C, seed=1, bias_std=0.1)
    asm = Assembler(synth.vocab_file('clevr'))
    tokens = synth.expert_mix_tokens(asm, N, T)
    labels = (np.arange(N) * 3) % C
    return feat, word_vecs, Wt, asm, tokens, labels


def _flat(g, names):
    return np.concatenate([np.asarray(g[n], np.float64).reshape(-1) for n in names])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    feat, word_vecs, Wt, asm, tokens, labels = _case()
    per = N // world
    sl = slice(rank * per, (rank + 1) * per)
    # the shard a rank owns: its images, its columns of word_vecs / tokens / labels
    m = ot.TorchOracleModules(feat[sl], word_vecs[:, sl], C, Wt)
    exprs, valid = asm.assemble(tokens[:, sl])
    scores, per_sample, avg, g, g_wv = ot.loss_and_grads(m, exprs, valid, labels[sl])
    names = sorted(g)
    flat = torch.from_numpy(np.concatenate([_flat(g, names), [avg]]))   # loss rides along
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                          # the ONE collective
    flat /= world
    gathered = [torch.zeros(per, C, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(scores.astype(np.float64)))  # eval: optional gather
    if rank == 0:
        np.save(out, {'flat': flat.numpy(), 'scores': torch.cat(gathered).numpy()},
                allow_pickle=True)
    dist.destroy_process_group()


def test_two_rank_data_parallel_matches_single_process(tmp_path):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / 'r0.npy')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out, allow_pickle=True).item()
    feat, word_vecs, Wt, asm, tokens, labels = _case()
    m = ot.TorchOracleModules(feat, word_vecs, C, Wt)
    exprs, valid = asm.assemble(tokens)
    scores, per_sample, avg, g, _ = ot.loss_and_grads(m, exprs, valid, labels)
    want = np.concatenate([_flat(g, sorted(g)), [avg]])
    np.testing.assert_allclose(got['flat'], want, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(got['scores'], scores, rtol=0, atol=1e-5)

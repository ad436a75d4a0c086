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


# ---- node-balanced shards (n2nmn_b200/sharding.py) --------------------------------
def test_layout_costs_and_balanced_shards():
    from n2nmn_b200 import sharding
    asm = Assembler(synth.vocab_file('clevr'))
    layouts = [['_Find', '_Count'], ['_Find'] + ['_Filter'] * 8 + ['_Exist'],
               ['_Find', '_Find', '_And', '_Exist'], ['_Scene', '_Count'],
               ['_Find', '_Transform', '_Filter', '_Transform', '_Filter', '_Describe'],
               ['_Find', '_Count'], ['_Find', '_Find', '_SameProperty'], ['_Scene', '_Exist']]
    tok = np.stack([asm.module_list2tokens(l, 12) for l in layouts], axis=1)
    cost = sharding.layout_costs(tok, asm)
    assert cost.tolist() == [len(l) for l in layouts]
    w = sharding.layout_costs(tok, asm, {'_Find': 10, '_Filter': 10})
    assert w[1] == 91 and w[3] == 2
    for world in (2, 4):
        sh = sharding.balanced_shards(cost, world)
        assert sorted(np.concatenate(sh).tolist()) == list(range(8))        # a partition
        assert all(len(s) == 8 // world for s in sh)                         # equal cardinality
        assert all((np.diff(s) > 0).all() for s in sh)
        cont = sharding.contiguous_shards(8, world)
        assert sharding.imbalance(cost, sh) <= sharding.imbalance(cost, cont)
    assert sharding.imbalance(cost, sharding.balanced_shards(cost, 2)) <= 1.04   # 16 vs 15 nodes
    assert sharding.imbalance(cost, sharding.contiguous_shards(8, 2)) > 1.15     # 18 vs 13
    # equal costs: the deal is deterministic and every rank gets N/world questions
    eq = sharding.balanced_shards(np.ones(12), 3)
    assert [len(s) for s in eq] == [4, 4, 4]
    # shard / unshard round trip on both batch axes
    x = np.arange(8 * 3).reshape(8, 3)
    sh = sharding.balanced_shards(cost, 4)
    assert np.array_equal(sharding.unshard([sharding.shard(x, s, 0) for s in sh], sh), x)
    y = np.arange(5 * 8 * 2).reshape(5, 8, 2)
    assert np.array_equal(sharding.unshard([sharding.shard(y, s, 1) for s in sh], sh, 1), y)
    with pytest.raises(ValueError):
        sharding.balanced_shards(cost, 3)
    # a large random batch: LPT lands within a node of perfect balance
    rng = np.random.RandomState(0)
    big = synth.random_valid_tokens(asm, 512, 20, seed=3)
    c = sharding.layout_costs(big, asm)
    assert sharding.imbalance(c, sharding.balanced_shards(c, 8)) <= 1.01
    assert rng is not None


def _balanced_worker(rank, world, port, out):
    from n2nmn_b200 import sharding
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    feat, word_vecs, Wt, asm, tokens, labels = _case()
    sh = sharding.balanced_shards(sharding.layout_costs(tokens, asm), world)
    mine = sh[rank]
    m = ot.TorchOracleModules(sharding.shard(feat, mine, 0), sharding.shard(word_vecs, mine, 1), C, Wt)
    exprs, valid = asm.assemble(sharding.shard(tokens, mine, 1))
    scores, per_sample, avg, g, g_wv = ot.loss_and_grads(m, exprs, valid, labels[mine])
    names = sorted(g)
    flat = torch.from_numpy(np.concatenate([_flat(g, names), [avg]]))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    gathered = [torch.zeros(len(mine), C, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(scores.astype(np.float64)))
    if rank == 0:
        np.save(out, {'flat': flat.numpy(),
                      'scores': sharding.unshard([t.numpy() for t in gathered], sh)},
                allow_pickle=True)
    dist.destroy_process_group()


def test_two_rank_balanced_shards_match_single_process(tmp_path):
    """The permuted (node-balanced) shards give the same gradient and, un-permuted, the same
    scores as one process on the whole batch."""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / 'r0.npy')
    mp.spawn(_balanced_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out, allow_pickle=True).item()
    feat, word_vecs, Wt, asm, tokens, labels = _case()
    m = ot.TorchOracleModules(feat, word_vecs, C, Wt)
    exprs, valid = asm.assemble(tokens)
    scores, per_sample, avg, g, _ = ot.loss_and_grads(m, exprs, valid, labels)
    want = np.concatenate([_flat(g, sorted(g)), [avg]])
    np.testing.assert_allclose(got['flat'], want, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(got['scores'], scores, rtol=0, atol=1e-5)


def _gather_worker(rank, world, port, out):
    from n2nmn_b200 import evaluate as ev
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    answers = ['w%d' % i for i in range(9)][rank::world]
    res = dict(split='val', num_questions=len(answers), answer_correct=len(answers) - rank,
               layout_correct=len(answers), layout_valid=len(answers), answer_accuracy=0.0,
               layout_accuracy=0.0, layout_validity=0.0, output_answers=answers)
    merged = ev.gather_rank_results(res)
    if rank == 1:                      # every rank holds the merged result
        np.save(out, merged, allow_pickle=True)
    dist.destroy_process_group()


def test_two_rank_eval_results_gather(tmp_path):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / 'r1.npy')
    mp.spawn(_gather_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out, allow_pickle=True).item()
    assert got['output_answers'] == ['w%d' % i for i in range(9)]
    assert got['num_questions'] == 9 and got['answer_correct'] == 5 + 3

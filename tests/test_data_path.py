"""CPU: the host data path (n2nmn_b200/data.py) against fixtures produced by executing the
reference's own files (tests/golden/make_golden_data.py -> golden_data.json): tokenizer,
functional-program -> expert layout, `prune_filter_module`, batch assembly, the prefetching reader."""
import copy
import json
import os

import numpy as np
import pytest

from n2nmn_b200 import data, synth
from n2nmn_b200.assembler import Assembler

G = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'golden_data.json')))


def test_tokenize_matches_reference():
    assert [data.tokenize(s) for s in G['sentences']] == G['tokenized']


def test_program_to_layout_matches_reference_and_does_not_mutate():
    for prog, want in zip(G['programs'], G['layouts']):
        before = copy.deepcopy(prog)
        assert data.program_to_layout(prog) == want
        assert prog == before
    # every layout the reference's generator emits assembles into a valid tree
    asm = Assembler(synth.vocab_file('clevr'))
    toks = np.array([asm.module_list2tokens(l, 12) for l in G['layouts']], np.int32).T
    assert asm.assemble(toks)[1].all()


def test_prune_filter_tokens():
    assert data.prune_filter_tokens(['_Find', '_Filter', '_Filter', '_Transform', '_Filter',
                                     '_Filter', '_Count']) == ['_Find', '_Transform', '_Filter', '_Count']
    assert data.prune_filter_tokens(['_Scene', '_Filter', '_Count']) == ['_Scene', '_Filter', '_Count']


def _write_imdb(tmp_path):
    L = G['loader']
    rng = np.random.RandomState(L['seed'])
    vq, va = tmp_path / 'vq.txt', tmp_path / 'va.txt'
    vq.write_text('\n'.join(L['words']) + '\n')
    va.write_text('\n'.join(L['answers']) + '\n')
    imdb = []
    for d in L['imdb']:
        fp = str(tmp_path / d['feature_path'])
        np.save(fp, rng.standard_normal((1, L['H'], L['W'], L['D'])).astype(np.float32))
        imdb.append(dict(d, feature_path=fp))
    path = str(tmp_path / 'imdb.npy')
    np.save(path, np.array(imdb), allow_pickle=True)
    params = dict(vocab_question_file=str(vq), vocab_answer_file=str(va),
                  T_encoder=L['T_encoder'], T_decoder=L['T_decoder'],
                  assembler=Assembler(synth.vocab_file('clevr')), batch_size=3)
    return path, params


def test_batch_loader_matches_reference(tmp_path):
    path, params = _write_imdb(tmp_path)
    imdb = data.load_imdb(path)
    for case in G['loader']['cases']:
        loader = data.ClevrBatchLoader(imdb, dict(params, prune_filter_module=case['prune']))
        b = loader.load_one_batch(case['sample_ids'])
        assert b['input_seq_batch'].dtype == np.int32 and b['image_feat_batch'].dtype == np.float32
        for k in ('input_seq_batch', 'seq_length_batch', 'answer_label_batch', 'gt_layout_batch'):
            assert b[k].tolist() == case[k], k
        assert b['image_path_list'] == case['image_path_list']
        assert abs(float(b['image_feat_batch'].astype(np.float64).sum()) - case['image_feat_sum']) < 1e-5
        n = len(case['sample_ids'])
        assert b['image_feat_batch'].shape == (n, G['loader']['H'], G['loader']['W'], G['loader']['D'])
    # the imdb itself is left untouched by pruning (the reference edits its token lists in place)
    assert imdb[2]['gt_layout_tokens'].count('_Filter') == 5


def test_batch_loader_fp16_feature_store(tmp_path):
    """feature_dtype='float16': the same batch with the grids rounded to half precision."""
    path, params = _write_imdb(tmp_path)
    imdb = data.load_imdb(path)
    ids = G['loader']['cases'][0]['sample_ids']
    b32 = data.ClevrBatchLoader(imdb, params).load_one_batch(ids)
    b16 = data.ClevrBatchLoader(imdb, dict(params, feature_dtype='float16')).load_one_batch(ids)
    assert b16['image_feat_batch'].dtype == np.float16
    assert np.array_equal(b16['image_feat_batch'], b32['image_feat_batch'].astype(np.float16))
    assert b16['input_seq_batch'].tolist() == b32['input_seq_batch'].tolist()
    with pytest.raises(ValueError):
        data.ClevrBatchLoader(imdb, dict(params, feature_dtype='int8'))


def test_data_reader_one_pass_order_and_short_last_batch(tmp_path):
    path, params = _write_imdb(tmp_path)
    rd = data.DataReader(path, shuffle=False, one_pass=True, prefetch_num=2, num_workers=3, **params)
    got = list(rd.batches())
    assert [len(b['seq_length_batch']) for b in got] == [3, 2]          # 5 samples, batch 3
    assert got[0]['image_path_list'] == ['img_0.png', 'img_1.png', 'img_2.png']
    assert got[1]['image_path_list'] == ['img_3.png', 'img_4.png']
    rd2 = data.DataReader(path, shuffle=True, one_pass=True, seed=3, **params)
    seen = sorted(p for b in rd2.batches() for p in b['image_path_list'])
    assert seen == ['img_%d.png' % i for i in range(5)]
    with pytest.raises(TypeError):
        data.DataReader('imdb.json', **params)


def test_data_reader_rank_shards_are_disjoint_and_cover_the_epoch(tmp_path):
    path, params = _write_imdb(tmp_path)
    seen = []
    for r in range(2):
        rd = data.DataReader(path, shuffle=True, one_pass=True, seed=11, rank=r, world=2,
                             **dict(params, batch_size=2))
        seen.append([p for b in rd.batches() for p in b['image_path_list']])
    assert not set(seen[0]) & set(seen[1])
    assert sorted(seen[0] + seen[1]) == ['img_%d.png' % i for i in range(5)]
    assert abs(len(seen[0]) - len(seen[1])) <= 1
    with pytest.raises(ValueError):
        data.DataReader(path, shuffle=True, rank=0, world=2, **params)      # no common seed
    with pytest.raises(ValueError):
        data.DataReader(path, shuffle=False, rank=2, world=2, **params)


def test_vocab_dict_unknown_words(tmp_path):
    f = tmp_path / 'v.txt'
    f.write_text('a\nb\n')
    v = data.VocabDict(str(f))
    assert v.word2idx('b') == 1 and v.idx2word(0) == 'a' and v.UNK_idx is None
    with pytest.raises(ValueError):
        v.word2idx('zzz')
    f.write_text('<unk>\na\n')
    assert data.VocabDict(str(f)).word2idx('zzz') == 0

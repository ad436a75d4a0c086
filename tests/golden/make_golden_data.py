#!/usr/bin/env python
"""Golden vectors for n2nmn_b200/data.py, produced by EXECUTING the reference's own data path:

  * exp_clevr/data/get_ground_truth_layout.py: its function definitions are exec'd (the file runs
    `add_gt_layout` on dataset files at import time, so it cannot be imported) and
    `linearize_program` is applied to hand-built CLEVR functional programs covering every entry of
    its function table;
  * util/clevr_train/data_reader.py + util/text_processing.py, imported unmodified: a small seeded
    imdb (.npy of dicts, one feature .npy per image) is written to a scratch directory and
    `BatchLoaderClevr.load_one_batch` / `tokenize` are run on it.

Needs /root/reference (present in the authoring container only); writes
tests/golden/golden_data.json, which travels with the repo. Re-run: python tests/golden/make_golden_data.py
"""
import ast
import json
import os
import sys
import tempfile

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def node(function, inputs, value=None):
    d = {'function': function, 'inputs': list(inputs), 'value_inputs': [] if value is None else [value]}
    return d


def programs():
    P = []
    # count after filters
    P.append([node('scene', []), node('filter_color', [0], 'red'), node('filter_shape', [1], 'cube'),
              node('count', [2])])
    # query after relate
    P.append([node('scene', []), node('filter_size', [0], 'large'), node('unique', [1]),
              node('relate', [2], 'left'), node('filter_material', [3], 'metal'),
              node('unique', [4]), node('query_color', [5])])
    # exist with intersect
    P.append([node('scene', []), node('filter_color', [0], 'blue'), node('unique', [1]),
              node('relate', [2], 'behind'), node('scene', []), node('filter_shape', [4], 'sphere'),
              node('unique', [5]), node('relate', [6], 'right'), node('intersect', [3, 7]),
              node('exist', [8])])
    # union + count
    P.append([node('scene', []), node('filter_color', [0], 'red'), node('scene', []),
              node('filter_material', [2], 'rubber'), node('filter_size', [3], 'small'),
              node('union', [1, 4]), node('count', [5])])
    # integer comparisons: the two counts are spliced out
    for cmp_fn in ('equal_integer', 'greater_than', 'less_than'):
        P.append([node('scene', []), node('filter_shape', [0], 'cylinder'), node('count', [1]),
                  node('scene', []), node('filter_color', [3], 'green'),
                  node('filter_size', [4], 'large'), node('count', [5]), node(cmp_fn, [2, 6])])
    # attribute comparisons: the two queries are spliced out
    for attr in ('color', 'material', 'shape', 'size'):
        P.append([node('scene', []), node('filter_size', [0], 'small'), node('unique', [1]),
                  node('query_' + attr, [2]), node('scene', []),
                  node('filter_color', [4], 'cyan'), node('unique', [5]),
                  node('query_' + attr, [6]), node('equal_' + attr, [3, 7])])
    # same_* and every query_*
    for attr in ('color', 'material', 'shape', 'size'):
        P.append([node('scene', []), node('filter_shape', [0], 'cube'), node('unique', [1]),
                  node('same_' + attr, [2]), node('filter_material', [3], 'metal'),
                  node('unique', [4]), node('query_' + attr, [5])])
    # scene with no filter (-> _Scene stays), and a stray unused scene node (second root)
    P.append([node('scene', []), node('count', [0])])
    P.append([node('scene', []), node('scene', []), node('filter_color', [1], 'gray'),
              node('exist', [2])])
    # a long filter run (exercises prune_filter_module downstream)
    P.append([node('scene', []), node('filter_size', [0], 'large'), node('filter_color', [1], 'red'),
              node('filter_material', [2], 'metal'), node('filter_shape', [3], 'cube'),
              node('unique', [4]), node('relate', [5], 'front'), node('filter_size', [6], 'small'),
              node('filter_color', [7], 'blue'), node('count', [8])])
    return P


def reference_linearize():
    src = open(os.path.join(REF, 'exp_clevr/data/get_ground_truth_layout.py')).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom, ast.FunctionDef)) or
            (isinstance(n, ast.Assign) and isinstance(n.value, (ast.Dict, ast.Set)))]
    ns = {}
    exec(compile(ast.Module(body=keep, type_ignores=[]), 'get_ground_truth_layout.py', 'exec'), ns)
    return ns['linearize_program']


def main():
    import copy
    out = {}
    lin = reference_linearize()
    progs = programs()
    out['programs'] = progs
    out['layouts'] = [lin({'program': copy.deepcopy(p)}) for p in progs]

    sys.path.insert(0, REF)
    from util import text_processing
    from util.clevr_train import data_reader as ref_reader
    sentences = ['Are there more cubes than yellow things?', 'What color is the sphere; it\'s big?',
                 'How many   objects are left of the tiny matte-block ?']
    out['sentences'] = sentences
    out['tokenized'] = [text_processing.tokenize(s) for s in sentences]

    from n2nmn_b200 import synth
    from n2nmn_b200.assembler import Assembler
    asm = Assembler(synth.vocab_file('clevr'))
    with tempfile.TemporaryDirectory() as tmp:
        words = ['<unk>', 'how', 'many', 'cubes', 'are', 'there', 'what', 'color', 'is', 'the',
                 'sphere', '?', 'red']
        answers = ['<unk>', 'yes', 'no', '0', '1', '2', 'red', 'blue']
        vq, va = os.path.join(tmp, 'vq.txt'), os.path.join(tmp, 'va.txt')
        open(vq, 'w').write('\n'.join(words) + '\n')
        open(va, 'w').write('\n'.join(answers) + '\n')
        rng = np.random.RandomState(5)
        H, W, D = 2, 3, 4
        layouts = [[m for m in l] for l in out['layouts']]
        imdb = []
        qs = ['how many cubes are there ?', 'what color is the sphere ?', 'are there purple cubes ?',
              'how many red cubes are there ?', 'what is the sphere ?']
        for i in range(5):
            fp = os.path.join(tmp, 'feat_%d.npy' % i)
            np.save(fp, rng.standard_normal((1, H, W, D)).astype(np.float32))
            imdb.append(dict(image_path='img_%d.png' % i, feature_path=fp,
                             question_tokens=text_processing.tokenize(qs[i]),
                             answer=['2', 'red', 'no', '1', 'blue'][i],
                             gt_layout_tokens=list(layouts[[0, 1, len(layouts) - 1, 8, 3][i]])))
        imdb_arr = np.array(imdb)
        params = dict(vocab_question_file=vq, vocab_answer_file=va, T_encoder=9, T_decoder=12,
                      assembler=asm, batch_size=3)
        cases = []
        for prune in (False, True):
            loader = ref_reader.BatchLoaderClevr(copy.deepcopy(imdb_arr),
                                                 dict(params, prune_filter_module=prune))
            b = loader.load_one_batch([4, 0, 2, 1, 3])
            cases.append({'prune': prune, 'sample_ids': [4, 0, 2, 1, 3],
                          'input_seq_batch': b['input_seq_batch'].tolist(),
                          'seq_length_batch': b['seq_length_batch'].tolist(),
                          'answer_label_batch': b['answer_label_batch'].tolist(),
                          'gt_layout_batch': b['gt_layout_batch'].tolist(),
                          'image_path_list': b['image_path_list'],
                          'image_feat_sum': float(b['image_feat_batch'].astype(np.float64).sum())})
        out['loader'] = {'words': words, 'answers': answers, 'questions': qs, 'H': H, 'W': W,
                         'D': D, 'seed': 5,
                         'imdb': [dict(d, feature_path=os.path.basename(d['feature_path']))
                                  for d in imdb],
                         'T_encoder': 9, 'T_decoder': 12, 'cases': cases}
    with open(os.path.join(HERE, 'golden_data.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote golden_data.json:', len(progs), 'programs;', 'layouts e.g.', out['layouts'][:3])


if __name__ == '__main__':
    main()

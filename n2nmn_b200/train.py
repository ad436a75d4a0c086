"""Training loop around the hot path: what `run_training` of
exp_clevr/train_clevr_gt_layout.py:161-222 (and its policy-search sibling
train_clevr_rl_gt_layout.py:178-243) does per iteration, with the module network — forward,
backward, clip, Adam — on the GPU.

Scope: the reference trains the seq2seq layout generator and the module network jointly. Here
the layout generator has no backward pass, so it is FROZEN: it supplies the layout
tokens (teacher forced with the ground-truth layouts as in train_clevr_gt_layout.py, decoded
greedily, or sampled — `AttentionSeq2Seq(decoder_sampling=True)` as the policy-search script
does, with its `log_seq_prob` fed to the REINFORCE bookkeeping through `log_seq_prob_fn`) and the
attended word vectors; the module network is what learns. Per iteration the
loop returns/logs the reference's quantities (loss, accuracy cur/avg, validity, :197-213) and
writes snapshots in the TensorFlow checkpoint format under the reference's variable names
(`snapshot_saver.save`, :216-219) every `snapshot_interval` iterations.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import checkpoint


def run_training(trainer, batches, assembler, max_iter, word_vecs_fn, layout_fn=None,
                 snapshot_dir=None, snapshot_interval=10000, log_interval=20, log=print,
                 accuracy_decay=0.99, log_seq_prob_fn=None):
    """trainer: ModuleNetTrainer; batches: iterable of data_reader dicts
    (n2nmn_b200.data.DataReader.batches()); word_vecs_fn(batch, tokens) -> [T,N,Dt] device tensor
    (e.g. from AttentionSeq2Seq.forward(..., use_gt_layout=True, gt_layout_batch=tokens));
    layout_fn(batch) -> [T,N] tokens (default: the batch's ground-truth layouts);
    log_seq_prob_fn(batch, tokens) -> [N] device tensor for the REINFORCE bookkeeping, or None.
    Returns the list of per-iteration records."""
    layout_fn = layout_fn or (lambda b: b['gt_layout_batch'])
    dev = trainer.m.device
    avg_accuracy, history = 0.0, []
    if snapshot_dir:
        os.makedirs(snapshot_dir, exist_ok=True)
    for n_iter, batch in enumerate(batches):
        if n_iter >= max_iter:
            break
        tokens = np.ascontiguousarray(layout_fn(batch), dtype=np.int32)
        feat = batch.get('image_feat_pinned')
        feat = feat if feat is not None else torch.from_numpy(batch['image_feat_batch'])
        feat = feat.to(dev, non_blocking=True)
        word_vecs = word_vecs_fn(batch, tokens).to(dev)
        labels = np.asarray(batch['answer_label_batch'])
        lsp = log_seq_prob_fn(batch, tokens) if log_seq_prob_fn is not None else None
        out = trainer.train_step(feat, word_vecs, tokens, labels, log_seq_prob=lsp)
        predictions = np.argmax(out['scores'].cpu().numpy(), axis=1)
        validity = out['validity']
        accuracy = float(np.mean(np.logical_and(validity, predictions == labels)))   # :197-199
        avg_accuracy += (1 - accuracy_decay) * (accuracy - avg_accuracy)
        rec = {'iter': n_iter + 1, 'loss': out['avg_sample_loss'], 'accuracy': accuracy,
               'avg_accuracy': avg_accuracy, 'validity': float(np.mean(validity)),
               'baseline': out['baseline']}
        history.append(rec)
        if (n_iter + 1) % log_interval == 0 or (n_iter + 1) == max_iter:
            log('iter = %d\n\tloss = %f, accuracy (cur) = %f, accuracy (avg) = %f, validity = %f'
                % (rec['iter'], rec['loss'], accuracy, avg_accuracy, rec['validity']))
        if snapshot_dir and ((n_iter + 1) % snapshot_interval == 0 or (n_iter + 1) == max_iter):
            snapshot_file = os.path.join(snapshot_dir, '%08d' % (n_iter + 1))
            checkpoint.export_module_weights(
                snapshot_file, {k: v.detach().cpu().numpy() for k, v in trainer.weights().items()})
            log('snapshot saved to ' + snapshot_file)
    return history

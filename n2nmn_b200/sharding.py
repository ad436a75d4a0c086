"""Question shards for the data-parallel runs: questions are independent — no op
of models_*/nmn3_modules.py crosses `batch_idx` — so a global batch splits into one shard per GPU
with no collective in evaluation and ONE all-reduce of the flat gradient in training
(`ModuleNetTrainer.train_step`). The reference itself is single-device (one `session.run` per
batch, exp_clevr/eval_clevr.py:96-133); this file has no counterpart there.

Contiguous slices give every rank the same number of questions but not the same work: a layout's
cost is its node count (a `_Find → _Count` question is 2 nodes, a deep `_Filter` chain 10+), and
the step ends when the slowest rank does. `balanced_shards` deals the questions so that every rank
gets the same NUMBER of questions (the contexts are created for a fixed `max_batch`) and nearly
the same total cost; `shard` / `unshard` apply and undo the permutation on the batch axis of
features `[N, …]`, word vectors / tokens `[T, N, …]` and scores `[N, C]`.
Host-side numpy only: the permutation is applied where the batch is assembled (data loader), the
GPU path sees ordinary contiguous shards.
"""
from __future__ import annotations

import numpy as np


def layout_costs(layout_tokens, assembler, op_cost=None):
    """Cost of every question of `layout_tokens` [T, N]: the number of module tokens before the
    first `<eos>` (the node count of its tree), or Σ `op_cost[module_name]` (default 1 per node)
    to weigh e.g. the Find-type nodes, which own the feature-grid contraction."""
    tok = np.asarray(layout_tokens)
    if tok.ndim != 2:
        raise ValueError('layout_tokens must be [T, N]')
    eos = assembler.EOS_idx
    live = np.cumsum(tok == eos, axis=0) == 0          # positions before the first <eos>
    if op_cost is None:
        return live.sum(axis=0).astype(np.float64)
    table = np.array([float(op_cost.get(n, 1.0)) for n in assembler.module_names], np.float64)
    if tok.min() < 0 or tok.max() >= len(table):
        raise ValueError('token outside the layout vocabulary')
    return (table[tok] * live).sum(axis=0)


def contiguous_shards(num_questions, world):
    """The plain split: rank r owns questions [r·N/world, (r+1)·N/world)."""
    if num_questions % world:
        raise ValueError('the global batch (%d) must divide by the world size (%d)'
                         % (num_questions, world))
    per = num_questions // world
    return [np.arange(r * per, (r + 1) * per) for r in range(world)]


def balanced_shards(costs, world):
    """Index arrays (one per rank, each sorted ascending, N/world entries) with nearly equal total
    cost: longest-processing-time greedy under an equal-cardinality constraint — questions in
    order of decreasing cost, each to the rank with the least cost so far that still has room —
    then pairwise swaps between the heaviest rank and the others while one lowers the maximum.
    Deterministic (ties keep the original question order)."""
    costs = np.asarray(costs, np.float64)
    n = costs.shape[0]
    if n % world:
        raise ValueError('the global batch (%d) must divide by the world size (%d)' % (n, world))
    per = n // world
    order = np.argsort(-costs, kind='stable')
    load = np.zeros(world)
    room = np.full(world, per)
    owner = np.empty(n, np.int64)
    for q in order:
        open_ranks = np.flatnonzero(room > 0)
        r = open_ranks[np.argmin(load[open_ranks])]     # first of the least loaded
        owner[q] = r
        load[r] += costs[q]
        room[r] -= 1
    for _ in range(4 * n):                 # each accepted swap lowers the pair's maximum
        h = int(np.argmax(load))
        mine = np.flatnonzero(owner == h)
        best = None
        for r in range(world):
            if r == h or load[h] - load[r] <= 0:
                continue
            theirs = np.flatnonzero(owner == r)
            d = costs[mine][:, None] - costs[theirs][None, :]        # moved from h to r
            after = np.maximum(load[h] - d, load[r] + d)
            after[d <= 0] = np.inf
            i, j = np.unravel_index(np.argmin(after), after.shape)
            if after[i, j] < load[h] - 1e-12 and (best is None or after[i, j] < best[0]):
                best = (after[i, j], mine[i], theirs[j], r, d[i, j])
        if best is None:
            break
        _, a, b, r, d = best
        owner[a], owner[b] = r, h
        load[h] -= d
        load[r] += d
    return [np.flatnonzero(owner == r) for r in range(world)]


def imbalance(costs, shards):
    """max over ranks of the shard cost ÷ mean shard cost (1.0 = perfectly balanced): the factor by
    which the slowest rank stretches the step."""
    costs = np.asarray(costs, np.float64)
    tot = np.array([costs[s].sum() for s in shards])
    return float(tot.max() / max(tot.mean(), 1e-300))


def shard(array, index, batch_axis):
    """This rank's slice of a global-batch array: `index` = its entry of balanced_shards()."""
    return np.ascontiguousarray(np.take(np.asarray(array), index, axis=batch_axis))


def unshard(parts, shards, batch_axis=0):
    """Inverse of `shard` over all ranks: `parts[r]` is rank r's result (e.g. scores [N/world, C]
    after an all-gather), `shards` the index arrays; returns the global-batch array in the
    original question order."""
    parts = [np.asarray(p) for p in parts]
    n = int(sum(len(s) for s in shards))
    shape = list(parts[0].shape)
    shape[batch_axis] = n
    out = np.empty(shape, parts[0].dtype)
    for p, s in zip(parts, shards):
        idx = [slice(None)] * out.ndim
        idx[batch_axis] = s
        out[tuple(idx)] = p
    return out

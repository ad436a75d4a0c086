#!/usr/bin/env python
"""Per-CTA phase timeline of the three kernels from the -DN2NMN_EXP_TIMELINE build
(N2NMN_LIB=.../libn2nmn_b200_timeline.so python tools/timeline.py)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from n2nmn_b200 import _lib, synth, weights as wts
from n2nmn_b200.assembler import Assembler
from n2nmn_b200.executor import LayoutExecutor

CS = int(os.environ.get("N2NMN_TREE_CLUSTER", 4))   # CTAs per question (rank 0 = CTA CS*q)
B, H, W, D, T, Cc = 64, 10, 15, 512, 20, 28
asm = Assembler(synth.vocab_file('clevr'))
toks = synth.expert_mix_tokens(asm, B, T)
weights = wts.init_weights('clevr', H, W, D, Cc, seed=0, bias_std=0.1)
P = 10
feats, wvs = [], []
for i in range(P):
    f, w = synth.make_inputs(B, H, W, D, T, seed=1234 + i)
    feats.append(torch.from_numpy(f).cuda()); wvs.append(torch.from_numpy(w).cuda())
ex = LayoutExecutor('clevr', feats[0], wvs[0], Cc, asm, weights=weights, max_batch=B, max_T=T)
lib = _lib.lib()
buf = torch.zeros(3 * 512 * 64, dtype=torch.int64, device='cuda')
lib.n2nmn_exp_set_timeline.argtypes = [C.c_void_p]
assert lib.n2nmn_exp_set_timeline(C.c_void_p(buf.data_ptr())) == 0
for i in range(12):
    ex.forward_device(feats[i % P], wvs[i % P], toks)
torch.cuda.synchronize()
buf.zero_()
ex.forward_device(feats[3], wvs[3], toks)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(3, 512, 64)
clk, gt = t[:, :, :32], t[:, :, 32:]
g0 = gt[gt > 0].min()
names = ['text', 'proj', 'tree']
from n2nmn_b200 import config as cfgmod
cb = ex.compile_tokens(toks)
nodes = cb.nodes()
qptr = np.zeros(B + 1, np.int64)
for op, t_, b_, dep, i0, i1 in nodes:
    qptr[b_ + 1] += 1
qptr = np.cumsum(qptr)
# tree kernel: per-op node durations on rank 0 of each cluster (cluster size 4 -> CTA 4*q)
per_op = {}
for q in range(B):
    cta = CS * q
    if clk[2, cta, 0] == 0:
        continue
    prev = clk[2, cta, 3]
    for j, i in enumerate(range(qptr[q], qptr[q + 1])):
        cur = clk[2, cta, 4 + j]
        if cur == 0:
            break
        per_op.setdefault(cfgmod.OP_NAMES[nodes[i][0]], []).append(cur - prev)
        prev = cur
print('== tree kernel, rank-0 CTA, cycles per node by module:')
for k_, v_ in sorted(per_op.items(), key=lambda kv: -np.median(kv[1])):
    print('   %-18s n=%3d median %6.0f max %6.0f' % (k_, len(v_), np.median(v_), max(v_)))
for k in range(3):
    used = np.where(clk[k, :, 0] > 0)[0]
    if len(used) == 0:
        continue
    print('== %s: %d CTAs; kernel span (globaltimer) start %.2f us .. last stamp %.2f us'
          % (names[k], len(used), (gt[k][gt[k] > 0].min() - g0) / 1e3, (gt[k].max() - g0) / 1e3))
    # per-slot cycles relative to the CTA's slot 0, median and max over CTAs
    for slot in range(1, 32):
        col = clk[k, used, slot]
        ok = col > 0
        if ok.sum() == 0:
            continue
        rel = col[ok] - clk[k, used, 0][ok]
        gts = (gt[k, used, slot][ok] - g0) / 1e3
        print('  slot %2d: n=%3d  cycles since CTA start  median %7.0f  max %7.0f | abs time us median %.2f max %.2f'
              % (slot, ok.sum(), np.median(rel), rel.max(), np.median(gts), gts.max()))

# phase stamps inside Transform (20..24) and Describe/SameProperty (26..30), rank-0 CTAs
def phases(name, slots):
    rows_ = []
    for q in range(B):
        cta = CS * q
        v_ = [clk[2, cta, s_] for s_ in slots]
        if all(x > 0 for x in v_):
            rows_.append(np.diff(v_))
    if rows_:
        print(name, 'n=%d' % len(rows_), 'median cycles per phase', np.median(np.array(rows_), axis=0))
phases('transform [stage->ready, compute, cluster sync, gather]', [20, 21, 22, 23, 24])
phases('stencil   [bank frags, (stamp), m-tile 0, m-tile 1, m-tile 2, sync+reduce]', [21, 25, 17, 18, 19, 31, 22])
phases('pooled    [copy+softmax, rowsum, cluster sync, partial-sum+normalize(rank0)]', [26, 27, 28, 29, 30])
rows_ = []
for cta in range(148):
    v_ = [clk[1, cta, j] for j in (5, 7, 28, 29, 30, 6)]
    if all(x > 0 for x in v_):
        rows_.append(np.diff(v_))
if rows_:
    print('proj FIND tiles: [chunks, bar, outputs, bar, to-exit] n=%d' % len(rows_), np.median(np.array(rows_), axis=0))
rows_ = []
for cta in range(148):
    if clk[1, cta, 28] == 0 and clk[1, cta, 5] > 0 and clk[1, cta, 7] > 0:
        rows_.append([clk[1, cta, 7] - clk[1, cta, 5], clk[1, cta, 6] - clk[1, cta, 7]])
if rows_:
    print('proj STORED tiles: [chunks, to-exit] n=%d' % len(rows_), np.median(np.array(rows_), axis=0))
# proj epilogue: the four 32-column chunks of warp 4 (slots 24..27) after tmem_full (slot 5)
for kind in ('FIND', 'STORED'):
    rows_ = []
    for cta in range(148):
        is_find = clk[1, cta, 28] > 0
        if is_find != (kind == 'FIND'):
            continue
        v_ = [clk[1, cta, 5]] + [clk[1, cta, 24 + j] for j in range(4)] + [clk[1, cta, 7]]
        if all(x > 0 for x in v_):
            rows_.append(np.diff(v_))
    if rows_:
        a_ = np.array(rows_)
        print('proj %s tiles: cycles per chunk [c0, c1, c2, c3, release] median' % kind,
              np.median(a_, axis=0), ' p90', np.percentile(a_, 90, axis=0))

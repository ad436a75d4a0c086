#!/usr/bin/env python
"""bench.py — CLEVR questions/sec of the module-network hot path on B200 (BASELINE.json metric).

One step = one pass of the hot path over one CLEVR-shaped batch (default 64 questions,
10x15x512 pool5 grid, T=20 layout tokens, expert-layout mix): host layout compile (C++) ->
table upload -> text projection -> tcgen05 conv_image contraction with fused Find epilogue ->
tree kernel -> scores [64,28] on device. Inputs come from a pool of distinct batches resident in
HBM that is larger than L2, walked round-robin, so no step re-reads a cached batch.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference ...                           # CPU arm (the oracle restatement)

Under torchrun every rank owns one GPU and its own shard of the questions (weak scaling, no
data-path collective: questions are independent); timing is CUDA events between
barriers, max over ranks; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, D, T_DEC, C, TEXT_DIM = 10, 15, 512, 20, 28, 300
METRIC = 'clevr_questions_per_sec'
UNIT = 'questions/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='questions per GPU per step')
    ap.add_argument('--layouts', default='expert', choices=['expert', 'random', 'deep'])
    ap.add_argument('--pool', type=int, default=10, help='distinct resident batches (x19.7 MB)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='cpu_baseline sample budget')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--wave', action='store_true', help='depth-bucketed wave executor')
    ap.add_argument('--host-threads', type=int, default=0,
                    help='ignored (kept for old command lines): every context of the pool has its '
                         'own native worker thread')
    ap.add_argument('--no-train', action='store_true', help='skip the train-step measurement')
    ap.add_argument('--no-other-sets', action='store_true',
                    help='skip the random / deep layout sets reported beside the expert mix')
    ap.add_argument('--tree-cluster', type=int, default=None,
                    help='CTAs per question in the executor kernel (default: chosen by the pool)')
    ap.add_argument('--proj-ctas', type=int, default=None,
                    help='cap of the contraction kernel grid (default: chosen by the pool; 0 = all SMs)')
    ap.add_argument('--streams', type=int, default=12,
                    help='contexts/streams fed round-robin (independent batches overlap)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'],
                    bf16_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0,
                source='fallback (B200_PROFILING.md)')


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture of this same command (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f).get(kernel)
    return None if not d else d['dram_read_bytes'] + d['dram_write_bytes']


def make_tokens(asm, kind, n, seed):
    from n2nmn_b200 import synth
    if kind == 'expert':
        toks = synth.expert_mix_tokens(asm, n, T_DEC)
        rng = np.random.RandomState(seed)          # same mix, different question order per batch
        return np.ascontiguousarray(toks[:, rng.permutation(n)])
    if kind == 'random':
        return synth.random_valid_tokens(asm, n, T_DEC, seed=seed)
    return synth.random_valid_tokens(asm, n, T_DEC, seed=seed, ans_weight=0.15, min_depth=3,
                                     max_depth=12)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md): one
    `nvidia-smi -lms` process started right before the region and stopped right after it."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)      # let the first samples start flowing
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                out = ''
            rows = [[c.strip() for c in ln.split(',')] for ln in out.splitlines() if ln.strip()]
        sm = [float(r[1]) for r in rows if len(r) > 2 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace('.', '').isdigit()]
        reasons = []
        for name, col in (('hw_slowdown', 4), ('hw_thermal_slowdown', 5),
                          ('sw_thermal_slowdown', 6), ('sw_power_cap', 7)):
            if any(len(r) > col and r[col].lower().startswith('active') for r in rows):
                reasons.append(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(rows)}


def best_blas_threads(run_once):
    """OpenBLAS with one thread per core is slow on many-core hosts for these small GEMMs.
    Give the CPU arm its best shot: try a few thread counts, keep the fastest."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        return None, os.cpu_count()
    ncpu = os.cpu_count() or 1
    best = (None, None)
    for nt in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        with threadpool_limits(limits=nt):
            run_once()
            t0 = time.perf_counter()
            run_once()
            dt = time.perf_counter() - t0
        if best[0] is None or dt < best[0]:
            best = (dt, nt)
    return threadpool_limits(limits=best[1]), best[1]


def cpu_reference_qps(feat, word_vecs, weights, tokens_list, budget_s, min_batches=2):
    """Times the oracle restatement of the reference path (Assembler.assemble + TF-Fold-style
    depth-batched module calls, numpy/OpenBLAS fp32) on a bounded sample of the workload.
    Returns (questions/s, batches, seconds, BLAS threads used)."""
    from n2nmn_b200 import synth
    from n2nmn_b200.assembler import Assembler
    from oracle.nmn_oracle import OracleModules, run_depth_batched
    asm = Assembler(synth.vocab_file('clevr'))
    m = OracleModules(feat, word_vecs, C, weights, family='clevr')
    exprs, _ = asm.assemble(tokens_list[0])
    limiter, threads = best_blas_threads(lambda: run_depth_batched(m, exprs))
    n_q, t0, k = 0, time.perf_counter(), 0
    while True:
        tok = tokens_list[k % len(tokens_list)]
        exprs, _ = asm.assemble(tok)
        run_depth_batched(m, exprs)
        n_q += tok.shape[1]
        k += 1
        el = time.perf_counter() - t0
        if k >= min_batches and el >= budget_s:
            break
    if limiter is not None:
        limiter.restore_original_limits()
    return n_q / el, k, el, threads


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU restatement (oracle) of the reference's TF1 path on host cores;
    TF 1.0 + TF Fold cannot be installed here. Rank 0 only."""
    if rank != 0:
        return
    from n2nmn_b200 import synth, weights as wts
    from n2nmn_b200.assembler import Assembler
    asm = Assembler(synth.vocab_file('clevr'))
    feat, word_vecs = synth.make_inputs(args.batch, H, W, D, T_DEC, seed=1234)
    weights = wts.init_weights('clevr', H, W, D, C, seed=0, bias_std=0.1)
    toks = [make_tokens(asm, args.layouts, args.batch, seed=100 + i) for i in range(4)]
    from oracle.nmn_oracle import OracleModules, run_depth_batched
    m = OracleModules(feat, word_vecs, C, weights, family='clevr')
    exprs0 = asm.assemble(toks[0])[0]
    limiter, threads = best_blas_threads(lambda: run_depth_batched(m, exprs0))
    for i in range(max(args.warmup, 1)):
        run_depth_batched(m, asm.assemble(toks[i % 4])[0])
    steps = min(args.steps, 200)
    t0 = time.perf_counter()
    for i in range(steps):
        run_depth_batched(m, asm.assemble(toks[i % 4])[0])
    el = time.perf_counter() - t0
    qps = steps * args.batch / el
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': qps, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * el / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': 'CLEVR gt-layout eval, batch=%d, 10x15x512 pool5, %s layouts, T=%d'
                   % (args.batch, args.layouts, T_DEC), 'global_batch': args.batch},
        'cpu_baseline': {'value': qps, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': '%d batches of %d questions (oracle/nmn_oracle.py: numpy + '
                                   'OpenBLAS, Assembler.assemble included)' % (steps, args.batch)},
        'e2e': {'value': qps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(torch, index):
    """Run this process (and so its pinned-buffer allocations and the pool's worker threads) on
    the CPUs that are local to the GPU's PCIe root: host<->device copies from the other socket
    run at about half the rate on a two-socket box. Best effort; returns what was done."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open('/sys/bus/pci/devices/%s/local_cpulist' % bdf) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 'no local cpus in the affinity mask'
        os.sched_setaffinity(0, cpus)
        return 'bound to %d cpus local to %s (%s)' % (len(cpus), bdf, spec)
    except Exception as e:   # no sysfs in this container, or not a PCI device
        return 'not bound: %s' % e


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from n2nmn_b200 import _lib, synth, weights as wts
    from n2nmn_b200.assembler import Assembler
    from n2nmn_b200.executor import ExecutorPool, LayoutExecutor

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device. The product path has no CPU fallback.')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    B, P = args.batch, args.pool
    asm = Assembler(synth.vocab_file('clevr'))
    weights = wts.init_weights('clevr', H, W, D, C, seed=0, bias_std=0.1)
    # pool of distinct batches (per-rank seeds = per-rank shard of the global question stream)
    feats, wvs, toks = [], [], []
    for i in range(P):
        f, w = synth.make_inputs(B, H, W, D, T_DEC, seed=1234 + 1000 * rank + i)
        feats.append(torch.from_numpy(f).to(dev))
        wvs.append(torch.from_numpy(w).to(dev))
        toks.append(make_tokens(asm, args.layouts, B, seed=100 + 1000 * rank + i))
    flags = _lib.FLAG_WAVE_EXECUTOR if args.wave else 0
    K = max(1, args.streams)
    pool = ExecutorPool('clevr', feats[0], wvs[0], C, asm, weights=weights, num_streams=K,
                        flags=flags, max_batch=B, max_T=T_DEC, tree_cluster=args.tree_cluster,
                        proj_ctas=args.proj_ctas)
    ex = pool.executors[0]
    PROJ_CTAS, TREE_CLUSTER = pool.proj_ctas, pool.tree_cluster
    scores_k = [torch.empty((B, C), dtype=torch.float32, device=dev) for _ in range(K)]
    scores = scores_k[0]

    def step(i):
        # public API, one call per batch: bind the batch's device-resident features, compile its
        # layouts (C++), upload the tables, launch the kernels; asynchronous. Batches go
        # round-robin over K contexts/streams so independent batches overlap on the GPU.
        k = i % P
        pool.submit(feats[k], wvs[k], toks[k], out=scores_k[i % K])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pool.begin()
    for i in range(args.warmup):
        step(i)
    pool.end()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pool.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    pool.begin()          # the K streams start after e0 ...
    for i in range(args.steps):
        step(args.warmup + i)
    pool.end()            # ... and e1 is recorded after all of them have drained
    e1.record()
    host_ms = (time.perf_counter() - t_host0) * 1e3   # time the host needed to enqueue the steps
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pool.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- the other two layout sets of SURVEY.md §8(d) (random valid layouts; deep layouts), same
    #      pool, same inputs, shorter runs: reported beside the headline, not as the headline
    other_sets = None
    if args.layouts == 'expert' and not args.no_other_sets:
        other_sets = {}
        for kind in ('random', 'deep'):
            otoks = [make_tokens(asm, kind, B, seed=100 + 1000 * rank + i) for i in range(P)]
            n_o = max(200, min(args.steps // 4, 4000))
            pool.begin()
            for i in range(3 * K):
                pool.submit(feats[i % P], wvs[i % P], otoks[i % P], out=scores_k[i % K])
            pool.end()
            barrier()
            oe0, oe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            oe0.record()
            pool.begin()
            for i in range(n_o):
                pool.submit(feats[i % P], wvs[i % P], otoks[i % P], out=scores_k[i % K])
            pool.end()
            oe1.record()
            barrier()
            to = torch.tensor([oe0.elapsed_time(oe1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(to, op=dist.ReduceOp.MAX)
            nodes = int(np.mean([int((t != asm.EOS_idx).sum()) for t in otoks]))
            other_sets[kind] = {'value': world * B * n_o / (float(to.item()) * 1e-3), 'unit': UNIT,
                                'steps': n_o, 'nodes_per_batch': nodes}

    # ---- e2e: host (pinned) buffers in, host scores out, every step, through the public API
    e2e = None
    if not args.no_e2e:
        hp = min(P, 6)
        hf = [feats[i].cpu().pin_memory() for i in range(hp)]
        hw = [wvs[i].cpu().pin_memory() for i in range(hp)]
        hs = [torch.empty((B, C), dtype=torch.float32).pin_memory() for _ in range(hp)]
        pool.begin()
        for i in range(2 * K):
            pool.submit_host(hf[i % hp], hw[i % hp], toks[i % hp], hs[i % hp])
        pool.end()
        k_e2e = max(10, min(args.steps, 400))
        barrier()
        t0 = time.perf_counter()
        pool.begin()
        for i in range(k_e2e):
            pool.submit_host(hf[i % hp], hw[i % hp], toks[i % hp], hs[i % hp])
        pool.end()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        # the copied-back scores are the device path's scores
        chk, _ = ex.forward_device(feats[(k_e2e - 1) % hp], wvs[(k_e2e - 1) % hp],
                                   toks[(k_e2e - 1) % hp])
        torch.cuda.synchronize()
        assert torch.equal(hs[(k_e2e - 1) % hp], chk.cpu()), 'e2e scores differ from device path'
        te = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {'value': world * B * k_e2e / float(te.item()), 'unit': UNIT,
               'h2d_bytes_per_step': int(hf[0].numel() * 4 + hw[0].numel() * 4),
               'd2h_bytes_per_step': int(hs[0].numel() * 4), 'steps': k_e2e,
               'how': 'ExecutorPool.submit_host: pinned host features+word_vecs -> async H2D -> '
                      'C++ layout compile -> kernels -> async D2H scores, %d streams, every '
                      'step, wall clock around the loop + final synchronize' % K}

    # ---- roofline of the dominant kernel: per-launch CUDA events, separate pass of the same steps
    roof, kernel_us = None, {}
    if rank == 0:
        pk = peaks()
        def timed_pass():
            ex.set_profiling(True)
            acc, bytes_acc, flops_acc, n = {}, 0, 0, 0
            for i in range(min(args.steps, 50)):
                ex.forward_device(feats[i % P], wvs[i % P], toks[i % P], out=scores)
                for name, us in ex.launch_times():
                    acc.setdefault(name, []).append(us)
                info = ex.last_step_info()
                bytes_acc += info['kernel_bytes'][1]
                flops_acc += info['kernel_flops'][1]
                n += 1
            ex.set_profiling(False)
            return {k: float(np.mean(v)) for k, v in acc.items()}, bytes_acc / max(n, 1), \
                flops_acc / max(n, 1)

        def fractions(us, nbytes, nflops):
            dur = us * 1e-6
            gbs, tfs = nbytes / dur / 1e9, nflops / dur / 1e12
            return gbs, tfs, gbs / pk['hbm_gbs'], tfs / (pk['bf16_tflops'] / 2)

        # as run in the timed region (the pool's narrow grid when several batches are in flight)
        kernel_us, nbytes, nflops = timed_pass()
        proj = 'proj_umma_kernel'
        if proj in kernel_us:
            tf32_peak = pk['bf16_tflops'] / 2
            gbs, tfs, hbm_frac, tc_frac = fractions(kernel_us[proj], nbytes, nflops)
            bound = 'hbm' if hbm_frac >= tc_frac else 'tensor'
            grid = min(PROJ_CTAS, 148) if PROJ_CTAS > 0 else 148
            roof = {'kernel': proj, 'bound': bound,
                    'achieved': gbs if bound == 'hbm' else tfs,
                    'peak': pk['hbm_gbs'] if bound == 'hbm' else tf32_peak,
                    'unit': 'GB/s' if bound == 'hbm' else 'TFLOP/s',
                    'frac': max(hbm_frac, tc_frac), 'traffic': ncu_traffic(proj),
                    'hbm_frac': hbm_frac, 'tensor_frac_of_tf32_peak': tc_frac,
                    'avg_launch_us': kernel_us[proj],
                    'algorithmic_bytes_per_launch': nbytes,
                    'flops_per_launch': nflops,
                    'peak_source': pk['source'] + '; TF32 peak taken as bf16 burst / 2',
                    'share_of_step': kernel_us[proj] / max(sum(kernel_us.values()), 1e-9),
                    'grid_ctas': grid,
                    'frac_of_occupied_sms': max(hbm_frac, tc_frac) / (grid / 148.0),
                    'note': 'as launched in the timed region: a persistent grid of %d CTAs (one '
                            'per SM) so that the other SMs run the other in-flight batches; '
                            'frac is against the WHOLE chip' % grid}
            if PROJ_CTAS > 0:   # the same launch spread over every SM (lowest single-batch latency)
                ex.set_proj_ctas(0)
                kus2, nb2, nf2 = timed_pass()
                ex.set_proj_ctas(PROJ_CTAS)
                g2, t2, hf2, tf2 = fractions(kus2[proj], nb2, nf2)
                roof['full_grid'] = {'grid_ctas': 148, 'avg_launch_us': kus2[proj],
                                     'hbm_frac': hf2, 'tensor_frac_of_tf32_peak': tf2,
                                     'frac': max(hf2, tf2)}

    # ---- config 3: policy-search train step (fwd + bwd + ONE NCCL all-reduce + clip + Adam),
    #      T=10 as in exp_clevr/train_clevr_rl_gt_layout.py; reported beside the eval headline
    train = None
    if not args.no_train:
        from n2nmn_b200.trainer import ModuleNetTrainer
        T_TRAIN = 10
        tr_ex = LayoutExecutor('clevr', feats[0], wvs[0][:T_TRAIN].contiguous(), C, asm,
                               weights=weights, max_batch=B, max_T=T_TRAIN)
        tr = ModuleNetTrainer(tr_ex)
        ttok = [np.ascontiguousarray(synth.expert_mix_tokens(asm, B, T_TRAIN)[
            :, np.random.RandomState(7 + i).permutation(B)]) for i in range(P)]
        twv = [w[:T_TRAIN].contiguous() for w in wvs]
        tlab = [np.random.RandomState(11 + i).randint(0, C, size=B) for i in range(P)]
        lsp = torch.full((B,), -2.0, device=dev)
        for i in range(5):
            tr.train_step(feats[i % P], twv[i % P], ttok[i % P], tlab[i % P], log_seq_prob=lsp)
        k_tr = max(10, min(args.steps, 50))
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for i in range(k_tr):
            out = tr.train_step(feats[i % P], twv[i % P], ttok[i % P], tlab[i % P],
                                log_seq_prob=lsp)
        t1e.record()
        barrier()
        tms = torch.tensor([t0e.elapsed_time(t1e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        train = {'questions_per_sec': world * B * k_tr / (float(tms.item()) * 1e-3),
                 'ms_per_step': float(tms.item()) / k_tr, 'steps': k_tr, 'global_batch': B * world,
                 'T_decoder': T_TRAIN, 'last_avg_sample_loss': out['avg_sample_loss'],
                 'what': 'fwd + bwd + all-reduce(flat grads, %d floats) + per-tensor clip + Adam '
                         '+ weight re-pack' % (tr.flat_size + 1)}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        f0, w0 = feats[0].cpu().numpy(), wvs[0].cpu().numpy()
        qps, nb, el, threads = cpu_reference_qps(f0, w0, weights, toks[:4], args.cpu_seconds)
        cpu = {'value': qps, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'host_cpus': os.cpu_count(),
               'sample': '%d batches of %d questions in %.1f s (oracle/nmn_oracle.py: numpy+OpenBLAS '
                         'fp32 depth-batched restatement, Assembler.assemble included)' % (nb, B, el)}

    if rank == 0:
        info = ex.last_step_info()
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 (fp32 in/out, fp32 accumulate)',
            'data': 'synthetic',
            'config': {'workload': 'CLEVR gt-layout eval, batch=%d/GPU, 10x15x512 pool5, %s '
                                   'layouts depth<=12, T=%d' % (B, args.layouts, T_DEC),
                       'global_batch': B * world, 'parallelism': 'dp%d (question shards, no '
                       'collective)' % world,
                       'cache': 'inputs larger than L2: %d distinct resident batches (%.0f MB) '
                                'walked round-robin' % (P, P * B * H * W * D * 4 / 1e6),
                       'executor': 'wave' if args.wave else 'tree',
                       'streams': K, 'host_threads': K, 'tree_cluster_ctas': TREE_CLUSTER,
                       'proj_grid_ctas': PROJ_CTAS if PROJ_CTAS > 0 else 148,
                       'nodes_per_batch': info['num_nodes'], 'max_depth': info['max_depth']},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
            'host_enqueue_ms_per_step': host_ms / args.steps, 'host_numa': numa,
            'roofline': roof, 'cpu_baseline': cpu, 'kernel_us': kernel_us, 'train_step': train,
            'other_layout_sets': other_sets,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Benchmark of the CaSE_RG answer-decode hot path (BASELINE.json: answer tokens/s decoded, ms per decode step).

    python bench.py --gpus N --steps K --warmup W [--config c1|c2|c3|c4|c5]     # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W [--config ..] # the UNMODIFIED reference on the host CPU

Configurations = BASELINE.json `configs` (default c2, the one the metric is quoted on):
  c1  CaSE, greedy (the in-module loop), B=8, 10 x 100-token passages                         weak scaling, step = 1 batch
  c2  CaSE, beam 4, B=64, 10 x 256-token passages, Lq=60, T=40, V=30522                       weak scaling, step = 1 batch
  c3  c2's shape over a CAsT-test-set-sized job: 1,008 rows sharded by DistributedSampler     STRONG scaling, step = the whole
      order (common/CumulativeTrainer.py:139), B=64 batches incl. the partial tail batch,     job incl. the result gather
      result gather INSIDE the timed region
  c4  GTTP pointer-generator, beam 4, B=128, V=50,000, Lc=60, Lb=10 x 100                    weak scaling, step = 1 batch
  c5  CaSE long context, beam 8, 256 queries in total (32 per batch), 20 x 512-token passages STRONG scaling, step = the job

A weak-scaling "step" = one pass of the hot path over one batch: per-batch prefill (memory K/V + Uk.mem projections)
followed by T decode steps.  Random-init weights (xavier, common/CumulativeTrainer.py:13-24), synthetic CAsT-shaped inputs.

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through the public API
(generations.beam_batches / FastGTTP) from pinned HOST buffers, H2D and the D2H of the answers inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = 256
WSEED, ISEED = 123456, 20211
CONFIGS = {
    'c1': dict(family='case', B=8, W=1, Lq=60, NP=10, Lp=100, T=40, V=30522, greedy=True, scaling='weak',
               label='CaSE default Run.py hyperparameters, greedy decode, batch 8, 10 passages x 100 tokens, Lq=60, T=40, '
                     'V=30522 (BASELINE.json configs[0])'),
    'c2': dict(family='case', B=64, W=4, Lq=60, NP=10, Lp=256, T=40, V=30522, scaling='weak',
               label='CaSE beam-4 decode, batch 64, 10 passages x 256 tokens, Lq=60, T=40, V=30522 (BASELINE.json configs[1])'),
    'c3': dict(family='case', B=64, W=4, Lq=60, NP=10, Lp=256, T=40, V=30522, scaling='strong', rows=1008,
               label='CaSE beam-4 decode over a CAsT-test-set-sized job (1,008 rows, 10 x 256-token passages) sharded '
                     'across the GPUs in DistributedSampler order, batches of 64 incl. the partial tail batch, result '
                     'gather inside the timed region (BASELINE.json configs[2])'),
    'c4': dict(family='gttp', B=128, W=4, Lq=60, NP=10, Lp=100, T=40, V=50000, scaling='weak',
               label='GTTP pointer-generator decode, beam 4, batch 128, 50k vocabulary, Lc=60, Lb=10 x 100 '
                     '(BASELINE.json configs[3])'),
    'c5': dict(family='case', B=32, W=8, Lq=60, NP=20, Lp=512, T=40, V=30522, scaling='strong', rows=256,
               label='CaSE long-context stress: 20 passages x 512 tokens, beam 8, 256 queries in batches of 32 sharded '
                     'across the GPUs (BASELINE.json configs[4])'),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--vocab-impl', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the fp32 line and the other-config lines of the default run')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--opt', type=lambda s: int(s, 0), default=0,
                    help='CASE_OPT_* bits (include/case_b200.h) for A/B runs: 0x1 no PDL, 0x2 row-block layer kernels, '
                         '0x4 no fused first stack, 0x8 no side stream, 0x10 no post linears, 0x20 context-form additive '
                         'attention, 0x40 no L2 evict-first, 0x80 dense tail, 0x200 unfused select')
    ap.add_argument('--cpu-budget', type=float, default=None, help='seconds of CPU work for the reference arm / cpu_baseline')
    ap.add_argument('--profile', type=int, default=0,
                    help='ncu helper: run one eager batch and bracket N decode steps (from t=20) with '
                         'cudaProfilerStart/Stop, then exit (use ncu --profile-from-start off)')
    return ap.parse_args()


def config_dict(args, cfg, extra=None):
    c = dict(workload=cfg['label'], name=args.config, queries_per_batch=cfg['B'], beam=cfg['W'], passages=cfg['NP'],
             passage_len=cfg['Lp'], query_len=cfg['Lq'], max_target_length=cfg['T'], vocab=cfg['V'],
             parallelism=f'queries sharded x{args.gpus}, no data-path collective, one result gather',
             l2='per-step K/V + Uk.mem streams (~0.6 GB of valid keys at c2) exceed the 126 MB L2; no explicit flush')
    if 'rows' in cfg:
        c['job_rows'] = cfg['rows']
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def reference_runner(cfg, n_queries, threads, seed_off=0):
    """The reference's own implementation of the path on the host: the UNMODIFIED reference sources (baseline/_ref, or
    /root/reference in the build container) through their public API - CaSETransformerSeqDecoder.forward for the in-module
    greedy loop, common/Generations.beam through the protocol adapters of baseline/refshim.py otherwise.  Falls back to the
    oracle port (kind 'port') only if the snapshot is missing.  -> (run(T) -> (seconds, tokens[int64 B x L]), kind)."""
    import torch
    from case_rg_b200 import synthetic as syn
    torch.set_num_threads(threads)
    V, W = cfg['V'], cfg['W']
    ns = None
    try:
        from baseline import refshim
        if refshim.reference_root() is not None:
            ns = refshim.load_reference()
    except Exception:
        ns = None
    if cfg['family'] == 'gttp':
        sd = syn.make_gttp_state(WSEED, V, H, H)
        inp = syn.make_gttp_inputs(ISEED + seed_off, n_queries, cfg['Lq'], cfg['NP'], cfg['Lp'], V, H)
        if ns is not None:
            from baseline import refshim
            vocab2id, id2vocab = syn.make_vocab(V)
            model = refshim.make_gttp_adapter(ns, H, H, vocab2id, id2vocab, max_dec_len=cfg['T'], beam_width=W)
            model.load_state_dict(sd, strict=False)
            model.eval()
            model.attach(inp)

            def run(Tn):
                data = {'id': inp.ids, 'context': inp.context, 'background': inp.background,
                        'background_map': ns.utils.build_map(inp.background_map, max=V)}
                t0 = time.perf_counter()
                with torch.no_grad():
                    out = ns.gen.beam(model, data, vocab2id, Tn, W)
                return time.perf_counter() - t0, out
            return run, 'reference'
        from oracle.gttp import GttpOracle
        from oracle import generations as OG
        orc = GttpOracle(sd)

        def run(Tn):
            t0 = time.perf_counter()
            with torch.no_grad():
                out = OG.beam(orc.stepper(inp, dense_onehot=True), Tn, W)
            return time.perf_counter() - t0, out
        return run, 'port'
    sd = syn.make_case_decoder_state(WSEED, V, H)
    inp = syn.make_case_inputs(ISEED + seed_off, n_queries, cfg['Lq'], cfg['NP'], cfg['Lp'], V, H)
    if ns is not None:
        from baseline import refshim
        dec = refshim.reference_decoder(ns, sd, V, H)
        vocab2id, _ = syn.make_vocab(V)

        def run(Tn):
            onehot = ns.utils.build_map(inp.source_map, max=V)        # Model.py:335 (part of the reference's forward)
            t0 = time.perf_counter()
            with torch.no_grad():
                if cfg.get('greedy'):
                    out = dec(inp.encode_memories, syn.BOS, syn.UNK, onehot, additional_decoder_feature=inp.answer_rep,
                              encode_weights=inp.encode_weights, encode_masks=inp.encode_masks, max_target_length=Tn)[3]
                else:
                    ad = refshim.CaseAdapter(ns, dec, inp)
                    out = ns.gen.beam(ad, {'id': inp.ids, 'source_map': onehot}, vocab2id, Tn, W)
            return time.perf_counter() - t0, out
        return run, 'reference'
    from oracle.case_decoder import CaseOracle
    from oracle import generations as OG
    orc = CaseOracle(sd)

    def run(Tn):
        t0 = time.perf_counter()
        with torch.no_grad():
            if cfg.get('greedy'):
                out, _ = orc.greedy_module(inp, Tn)
            else:
                out = OG.beam(orc.stepper(inp, dense_onehot=True), Tn, W)
        return time.perf_counter() - t0, out
    return run, 'port'


def _answer_tokens(out, greedy):
    """Answer tokens of a result tensor: B x T for the in-module greedy loop (it never stops, Model.py:94), else the
    best-sequence lengths (tokens after BOS up to and including EOS = the non-zero entries, Generations.py:188)."""
    return int(out.numel()) if greedy else int((out != 0).sum())


def run_reference_arm(args, cfg, emit=True):
    """bench.py --impl reference: the reference's own CPU implementation of the path on this box's host cores (all
    threads), same metric / unit / config as the b200 arm, each step a bounded sample of that workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    os.environ['CUDA_VISIBLE_DEVICES'] = ''          # reference helpers jump to CUDA whenever it is visible (Utils.py:26-27)
    import warnings
    warnings.filterwarnings('ignore')
    threads = os.cpu_count() or 1
    T, greedy = cfg['T'], bool(cfg.get('greedy'))
    budget_total = args.cpu_budget if args.cpu_budget else 150.0
    per_step_budget = budget_total / max(1, args.steps + args.warmup)
    # calibrate on one query: time of a 3-step decode -> seconds per (query, decode step); the cost per step grows with the
    # prefix (whole-prefix recompute), roughly (1 + 0.03 t)
    nq_full = cfg['B'] if greedy else 1
    run, kind = reference_runner(cfg, nq_full, threads)
    t_probe, _ = run(3)
    est_full = (t_probe / 3) * T * (1 + 0.03 * T)
    Ts = T
    while Ts > 3 and (t_probe / 3) * Ts * (1 + 0.03 * Ts) > per_step_budget:
        Ts -= 1
    nq = nq_full
    if not greedy and Ts == T:
        nq = int(max(1, min(8, per_step_budget // max(est_full, 1e-3))))    # <= 8 queries: the one-hot is B*S*V*4 bytes
        if nq != nq_full:
            run, kind = reference_runner(cfg, nq, threads)
    for _ in range(args.warmup):
        run(Ts)
    toks, dt, last = 0, 0.0, None
    for _ in range(args.steps):
        d, out = run(Ts)
        dt += d
        toks += _answer_tokens(out, greedy)
        last = out
    val = toks / dt
    sample = (f'{nq} quer{"y" if nq == 1 else "ies"} x {"greedy" if greedy else "beam %d" % cfg["W"]} x {Ts} decode steps per '
              f'bench step (of T={T}; full batch = {cfg["B"]} queries), '
              f'{"unmodified reference: " if kind == "reference" else "oracle port: "}'
              f'whole-prefix recompute, dense one-hot copy bmm, Python beam bookkeeping')
    line = dict(metric='answer_tokens_per_s', value=val, unit='tokens/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / max(1, args.steps), higher_is_better=True, scaling=cfg['scaling'],
                vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=config_dict(args, cfg, dict(sample=sample)),
                cpu_baseline=dict(value=val, unit='tokens/s', cores=threads, kind=kind, sample=sample),
                e2e=dict(value=val, unit='tokens/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                sample_tokens=last.tolist() if last is not None else None, sample_queries=nq, sample_steps=Ts)
    if emit:
        print(json.dumps(line), flush=True)
    return line


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------- workloads
CASE_KEYS = ('mem_q', 'mem_p', 'query', 'passage', 'prior_q', 'prior_p', 'answer_rep', 'source_map')
GTTP_KEYS = ('context', 'background', 'background_map', 'src_output', 'bg_output', 'init_state')


class Workload:
    """The local batches of one rank (pinned host copies + device-resident copies) and the two ways to run them."""

    def __init__(self, cfg, args, rank, world, dev, dtype, opt=0, resident=True):
        import torch
        from case_rg_b200 import synthetic as syn, _lib as L
        from case_rg_b200 import generations as FG
        from case_rg_b200.distributed import shard_indices
        self.cfg, self.args, self.rank, self.world, self.dev = cfg, args, rank, world, dev
        self.torch, self.L, self.FG = torch, L, FG
        self.family, self.T, self.W, self.V = cfg['family'], cfg['T'], cfg['W'], cfg['V']
        self.greedy = bool(cfg.get('greedy'))
        self.strong = cfg['scaling'] == 'strong'
        B = cfg['B']
        vocab_impl = args.vocab_impl if args.vocab_impl is not None else (1 if dtype == 'bf16' else 0)
        self.vocab_impl = vocab_impl
        graph = not args.no_graph
        if self.family == 'gttp':
            sd = syn.make_gttp_state(WSEED, self.V, H, H)
            self.model = FG.FastGTTP(sd, device=dev, dtype=dtype, max_dec_len=self.T, beam_width=self.W,
                                     vocab_impl=vocab_impl, use_graph=graph, opt=opt)
            self.keys = GTTP_KEYS
        else:
            sd = syn.make_case_decoder_state(WSEED, self.V, H)
            self.model = FG.FastCaSE(sd, device=dev, dtype=dtype, max_dec_len=self.T, beam_width=self.W,
                                     vocab_impl=vocab_impl, use_graph=graph, opt=opt)
            self.keys = CASE_KEYS
        self.mode = L.MODE_MODULE_GREEDY if self.greedy else L.MODE_BEAM
        # ---- rows of this rank
        if self.strong:
            self.n_total = cfg['rows']
            mine = shard_indices(self.n_total, rank, world)          # DistributedSampler(shuffle=False) order
            groups = [mine[i:i + B] for i in range(0, len(mine), B)]   # full batches + the partial tail batch
        else:
            self.n_total = world * B
            groups = [list(range(rank * B, (rank + 1) * B))]
        self.host, self.ids = [], []
        for g in groups:
            rows = [self._make_row(i) for i in g] if self.strong else None
            inp = self._cat(rows) if rows is not None else self._make_batch(ISEED + rank, B, rank * B)
            self.ids.append(torch.tensor(g, dtype=torch.int64))
            self.host.append({k: getattr(inp, k).pin_memory() for k in self.keys})
        self.resident = [{k: v.to(dev) for k, v in hb.items()} for hb in self.host] if resident else None
        self.h2d_bytes = sum(v.numel() * v.element_size() for hb in self.host for v in hb.values())
        self.tokens_local = None

    # one synthetic row per dataset index (strong-scaling jobs: every rank builds only its own rows)
    def _make_row(self, idx):
        return self._make_batch(ISEED + 7919 * (idx + 1), 1, idx)

    def _make_batch(self, seed, B, id_base):
        from case_rg_b200 import synthetic as syn
        c = self.cfg
        if self.family == 'gttp':
            return syn.make_gttp_inputs(seed, B, c['Lq'], c['NP'], c['Lp'], self.V, H, id_base=id_base)
        return syn.make_case_inputs(seed, B, c['Lq'], c['NP'], c['Lp'], self.V, H, id_base=id_base)

    def _cat(self, rows):
        torch = self.torch
        out = type('Batch', (), {})()
        for k in self.keys:
            setattr(out, k, torch.cat([getattr(r, k) for r in rows], 0))
        return out

    def search(self, data):
        if self.family == 'gttp':
            return self.model.fast_search(data, self.T, self.W, self.mode)
        return self.model.fast_search(data, self.T, self.W, self.mode)

    def gather(self, outs):
        """The path's only exchange: ids + answers of every rank to every rank (reference: per-rank result files,
        Utils.py:38-49).  Inside the timed region for the strong-scaling jobs."""
        from case_rg_b200.distributed import gather_answers
        torch = self.torch
        ids = torch.cat(self.ids).to(self.dev)
        toks = torch.zeros(ids.numel(), self.T, dtype=torch.int64, device=self.dev)
        r = 0
        for o in outs:
            toks[r:r + o.size(0), :o.size(1)] = o.to(self.dev)
            r += o.size(0)
        return gather_answers(ids, toks, self.T, self.n_total)

    def step_resident(self):
        if len(self.resident) > 1 and not self.greedy:
            # several local batches: the streaming face over the device-resident copies, so that the host reads the
            # answers of batch i while batch i+1 decodes (no per-batch host synchronisation between the decodes)
            outs = list(self.FG.beam_batches(self.model, iter(self.resident), None, self.T, self.W))
        else:
            outs = [self.search(d) for d in self.resident]
        if self.strong:
            return self.gather(outs)[1]
        return outs[-1]

    def step_e2e(self, n=1):
        """n passes through the public API from pinned host buffers: the streaming face (the copy of batch i+1 overlaps the
        decode of batch i) for the beam configurations; the in-module greedy loop copies, searches and reads back batch
        by batch."""
        torch = self.torch
        out = None
        for _ in range(n if self.strong else 1):
            reps = 1 if self.strong else n
            if self.greedy:
                outs = []
                for _r in range(reps):
                    for hb in self.host:
                        d = {k: v.to(self.dev, non_blocking=True) for k, v in hb.items()}
                        outs.append(self.search(d).cpu())
            else:
                gen = (hb for _r in range(reps) for hb in self.host)
                outs = list(self.FG.beam_batches(self.model, gen, None, self.T, self.W))
            out = self.gather(outs)[1].cpu() if self.strong else outs[-1]
        return out

    def count_tokens(self):
        """Answer tokens of one step over ALL ranks (all_reduce SUM of each rank's own count; strong-scaling jobs count
        every dataset row once - wrap-around padding of the shards is not counted)."""
        torch = self.torch
        import torch.distributed as dist
        if self.strong:
            _, toks = self.gather([self.search(d) for d in self.resident])
            torch.cuda.synchronize(self.dev)
            return _answer_tokens(toks, self.greedy)          # gathered: identical on every rank
        out = self.step_resident()
        torch.cuda.synchronize(self.dev)
        n = _answer_tokens(out, self.greedy)
        if self.world > 1:
            t = torch.tensor([n], device=self.dev, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            n = int(t)
        return n

    def launches_per_pass(self):
        eng = self.model.last_engine
        per_step = eng.kernel_launches_per_step() + (1 if self.family == 'case' else 0)
        prefill = 12 if self.family == 'case' else 0
        return len(self.host) * (self.T * per_step + prefill)


# ----------------------------------------------------------------------------- main arm
def main():
    args = parse()
    cfg = CONFIGS[args.config]
    if args.impl == 'reference':
        return run_reference_arm(args, cfg)

    import torch
    import torch.distributed as dist
    from case_rg_b200 import _lib as L

    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL's version banner (NCCL_DEBUG=VERSION, from the environment or /etc/nccl.conf) goes to stdout, next to the
        # JSON line: the environment wins over the conf file, and any debug output is sent to stderr
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    T = cfg['T']
    wl = Workload(cfg, args, rank, world, dev, args.dtype, opt=args.opt)

    if args.profile:
        import ctypes as C
        wl.model.use_graph = False
        wl.search(wl.resident[0])
        eng = wl.model.last_engine
        (eng._reset if wl.family == 'gttp' else eng.state.reset)()
        eng.args.mode, eng.args.max_len = wl.mode, T
        st = torch.cuda.current_stream(dev).cuda_stream
        for t in range(T):
            if t == 20:
                torch.cuda.synchronize(dev)
                torch.cuda.profiler.start()
            L.check(eng._step_fn(C.byref(eng.args), t, st), 'step')
            if t == 20 + args.profile - 1:
                torch.cuda.synchronize(dev)
                torch.cuda.profiler.stop()
        torch.cuda.synchronize(dev)
        print(json.dumps(dict(profiled_steps=args.profile, launches_per_step=eng.kernel_launches_per_step() + 1)))
        return

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return ms, wall, out

    warm = max(args.warmup, 3)
    for _ in range(warm):
        wl.step_resident()
    torch.cuda.synchronize(dev)
    wl.step_e2e(2)
    tokens_per_step = wl.count_tokens()
    eng = wl.model.last_engine

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, wall, out = timed(wl.step_resident, args.steps)
    if wl.strong:
        ms_e2e, wall_e2e, out_e = timed(lambda: wl.step_e2e(1), args.steps)
    else:
        ms_e2e, wall_e2e, out_e = timed(lambda: wl.step_e2e(args.steps), 1)
    # the PCIe rate this box gives the pinned input buffers (explains e2e vs value when the link is slow/shared)
    stage_dev = {k: torch.empty_like(v, device=dev) for k, v in wl.host[0].items()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for k, v in wl.host[0].items():
        stage_dev[k].copy_(v, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    h2d_ms = e0.elapsed_time(e1)
    h2d_one = sum(v.numel() * v.element_size() for v in wl.host[0].values())
    del stage_dev
    clk = clocks.stop() if rank == 0 else None

    # where the step goes: prefill (once per batch) vs the T decode steps, timed separately on the first local batch
    def _ev(fn, n=3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n
    d0 = wl.resident[0]
    wl.search(d0)
    eng0 = wl.model.last_engine
    if wl.family == 'gttp':
        prefill_ms = _ev(lambda: eng0.prefill(d0['src_output'], d0['bg_output'], d0['context'], d0['background'],
                                              d0['background_map'], d0['init_state']))
    else:
        prefill_ms = _ev(lambda: eng0.prefill(d0['mem_q'], d0['mem_p'], d0['query'].ne(0), d0['passage'].ne(0),
                                              d0['prior_q'], d0['prior_p'], d0['answer_rep'], d0['source_map']))
    decode_ms = _ev(lambda: eng0.decode(T, wl.mode, use_graph=not args.no_graph))

    value = tokens_per_step * args.steps / (ms * 1e-3)
    e2e_value = tokens_per_step * args.steps / (ms_e2e * 1e-3)
    d2h = int(out_e.numel() * out_e.element_size()) if out_e is not None else 0

    roof = cpu = parity = extras = None
    if rank == 0:
        roof = roofline(wl, eng0, decode_ms / T, torch, dev)
        if not args.no_cpu_baseline and world == 1:
            cpu, parity = cpu_baseline_and_parity(args, cfg, wl, torch)
        if not args.no_extras and world == 1 and args.config == 'c2' and args.dtype == 'bf16' and not args.opt:
            extras = extra_lines(args, dev, torch)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    launches = args.steps * wl.launches_per_pass()
    line = dict(metric='answer_tokens_per_s', value=value, unit='tokens/s', n_gpus=world, steps=args.steps,
                warmup=warm, ms_per_step=ms / args.steps, higher_is_better=True, scaling=cfg['scaling'],
                vs_baseline=None, dtype=args.dtype if args.dtype == 'bf16' else 'f32', data='synthetic',
                config=config_dict(args, cfg, dict(cuda_graph=not args.no_graph, opt=args.opt,
                                                   vocab_gemm='tcgen05' if wl.vocab_impl == 1 else 'simt',
                                                   batches_per_gpu_per_step=len(wl.host),
                                                   tail_batch_rows=int(wl.ids[-1].numel()))),
                ms_per_decode_step=decode_ms / T, prefill_ms=prefill_ms, decode_ms=decode_ms,
                answer_tokens_per_step=tokens_per_step,
                e2e=dict(value=e2e_value, unit='tokens/s', h2d_bytes_per_step=wl.h2d_bytes, d2h_bytes_per_step=d2h,
                         ms_per_step=ms_e2e / args.steps, h2d_ms_alone=h2d_ms, h2d_gbs_alone=h2d_one / (h2d_ms * 1e6),
                         pinned=all(v.is_pinned() for v in wl.host[0].values())),
                gpu_launches=launches, clocks=clk, roofline=roof, cpu_baseline=cpu, parity=parity,
                wall_s=dict(resident=wall, e2e=wall_e2e))
    if extras:
        line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extra_lines(args, dev, torch):
    """Short measurements beside the headline (N=1, default run only): the parity-green fp32 storage mode on the same
    config, and the other single-GPU BASELINE configs (c1, c4, the per-GPU share of c5) - value only, 3 steps each."""
    out = {}

    def quick(name, cfg, dtype, steps=3):
        try:
            wl = Workload(cfg, args, 0, 1, dev, dtype)
            for _ in range(2):
                wl.step_resident()
            toks = wl.count_tokens()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(steps):
                wl.step_resident()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            res = dict(value=toks / (ms * 1e-3), unit='tokens/s', ms_per_step=ms, dtype=dtype, steps=steps,
                       ms_per_decode_step_incl_prefill=ms / (cfg['T'] * len(wl.host)), workload=cfg['label'])
            del wl
            torch.cuda.empty_cache()
            return res
        except Exception as ex:      # evidence beside the headline: never fail the bench line over it
            return dict(unavailable=repr(ex)[:200])
    out['fp32'] = quick('c2', CONFIGS['c2'], 'fp32')
    out['from_token_ids'] = from_token_ids(args, CONFIGS['c2'], dev, torch)
    share = dict(CONFIGS['c5'], rows=32, label='per-GPU share of c5 at 8 GPUs: 32 queries, beam 8, 20 x 512-token passages')
    out['other_configs'] = dict(c1=quick('c1', CONFIGS['c1'], 'bf16'), c4=quick('c4', CONFIGS['c4'], 'bf16'),
                                c5_per_gpu_share=quick('c5', share, 'bf16'))
    return out


def from_token_ids(args, cfg, dev, torch, steps=3):
    """the whole of ``CaSE.do_test`` on the device.  Each pass copies only the token ids (query, passage,
    source_map) from pinned host memory, runs the pre-decode producers (encoder, two Interactions, transformer blocks,
    scorers, prior / answer_rep - case_rg_b200/producers.py), the prefill and the beam decode, and reads the answers back.
    Random-init producer weights of the reference's architecture; value in answer tokens / s like the headline."""
    try:
        from case_rg_b200 import synthetic as syn, generations as FG, _lib as L
        T, W, V, B = cfg['T'], cfg['W'], cfg['V'], cfg['B']
        model = FG.FastCaSE(syn.make_case_decoder_state(WSEED, V, H), device=dev, dtype='bf16', max_dec_len=T, beam_width=W,
                            vocab_impl=1, use_graph=not args.no_graph,
                            producers=syn.make_case_producer_state(WSEED + 1, V, H))
        prod = model.producers
        inp = syn.make_case_inputs(ISEED, B, cfg['Lq'], cfg['NP'], cfg['Lp'], V, H)
        host = {k: getattr(inp, k).pin_memory() for k in ('query', 'passage', 'source_map')}

        def one():
            d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            return model.search_ids(d, T, W, 'beam')['answer'].cpu()

        def ev(fn, n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(n):
                r = fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / n, r
        for _ in range(2):
            ans = one()
        toks = _answer_tokens(ans, False)
        ms, ans = ev(one, steps)
        q, p = host['query'].to(dev), host['passage'].to(dev)
        prod(q, p)                                   # untimed: the allocator settles on the stand-alone call pattern
        ms_prod, _ = ev(lambda: prod(q, p), max(steps, 5))
        res = dict(value=toks / (ms * 1e-3), unit='tokens/s', ms_per_batch=ms, producers_ms=ms_prod,
                   h2d_bytes_per_batch=sum(v.numel() * v.element_size() for v in host.values()),
                   d2h_bytes_per_batch=int(ans.numel() * ans.element_size()), answer_tokens_per_batch=toks, steps=steps,
                   workload='c2 from token ids: producers + prefill + beam-4 decode, host ids in, answers out')
        del model, prod
        torch.cuda.empty_cache()
        return res
    except Exception as ex:          # evidence beside the headline: never fail the bench line over it
        return dict(unavailable=repr(ex)[:300])


# ----------------------------------------------------------------------------- roofline
def kernel_times(step_fn, torch, dev):
    """CUPTI activity records of one extra, untimed batch (prefill + the decode graph): per kernel name the summed GPU
    time, the launch count and hence the IN-GRAPH duration per launch (not back-to-back launches of one kernel)."""
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step_fn()
        torch.cuda.synchronize(dev)
    tot = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and 'Memcpy' not in e.name and 'Memset' not in e.name:
            k = e.name.split('(')[0].split('<')[0].replace('void ', '').replace('cb::', '')
            us, n, mx = tot.get(k, (0.0, 0, 0.0))
            d = e.time_range.end - e.time_range.start
            tot[k] = (us + d, n + 1, max(mx, d))
    return tot


def roofline(wl, eng, ms_per_decode_step, torch, dev):
    """`roofline` object of the JSON line.  Top level = the dominant BANDWIDTH-bound kernel (the passage-memory cross-
    attention, 4 launches per decode step), timed IN the decode graph (CUPTI records of one batch); `kernels` = the same
    for every kernel with a byte or flop model; `step` = the whole decode step against the HBM peak (algorithmic bytes of
    an incremental decoder over the VALID keys; DESIGN.md §4 states every term).  In-graph durations are
    CUPTI's start-to-end of each launch: with programmatic dependent launch a kernel's record starts when its first CTA is
    resident, i.e. it includes the wait for the predecessor's tail - the per-kernel fractions are lower bounds."""
    L = wl.L
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm, which = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650.0, 'fallback')
    tf = peaks.get('bf16_tflops_sustained', 1405.0)
    try:
        kt = kernel_times(lambda: wl.search(wl.resident[0]), torch, dev)
    except Exception as ex:
        return dict(unavailable=repr(ex)[:200])
    total_us = sum(v[0] for v in kt.values())
    shares = {k: round(v[0] / total_us, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])[:10]}
    if wl.family != 'case':
        return dict(kernel=None, bound='hbm', achieved=None, peak=hbm, unit='GB/s', frac=None, traffic=None,
                    peak_source=which, step_shares=shares)
    B, W, R, T, V = eng.B, eng.W, eng.R, wl.T, eng.V
    S0, S1 = eng.S
    bf = eng.w.cdtype == L.BF16
    esz = 2 if bf else 4
    compact = bool(getattr(eng, 'compact', False))
    valid1 = int(eng.xcount.sum().item()) if compact else B * S1
    valid0 = int(eng.mask[0].sum().item())
    # ---- per-kernel models: algorithmic bytes (or flops) per launch
    models = {
        'cross_attn_part_kernel': ('hbm', valid1 * 2 * H * esz, 'K|V of the valid passage keys, once per query'),
        'additive_attn_gate_kernel': ('hbm', (valid1 + valid0) * (H * esz + 16) / 2.0,
                                      'Uk.mem rows + gate-projected keys of the valid keys (mean of the two launches)'),
        'vocab_gemm_tc_kernel': ('tensor', 2.0 * R * H * V, '2 R H V flops'),
        'vocab_base_kernel': ('l2', R * V * 4, 'one read of the fp32 logits tile (L2-resident)'),
        'prefill_project_tc_kernel': ('tensor', 2.0 * ((valid0 + valid1) / 2.0) * H * (8 * H + H),
                                      '2 valid_keys H (8H K|V columns + H Uk columns) flops (mean of the two launches; '
                                      'padded keys are neither gathered nor multiplied)'),
    }
    kernels = []
    for name, (bound, work, what) in models.items():
        if name not in kt:
            continue
        us, n, _ = kt[name]
        per = us / n
        if bound == 'tensor':
            ach = work / (per * 1e-6) / 1e12
            kernels.append(dict(kernel=name, bound=bound, achieved=ach, peak=tf, unit='TFLOP/s', frac=ach / tf,
                                us_per_launch_in_graph=per, launches=n, work_per_launch=work, model=what))
        else:
            ach = work / (per * 1e-6) / 1e9
            kernels.append(dict(kernel=name, bound=bound, achieved=ach, peak=hbm, unit='GB/s', frac=ach / hbm,
                                us_per_launch_in_graph=per, launches=n, work_per_launch=work, model=what))
    for name in ('layer_chain_kernel', 'sparse_tail_kernel'):
        if name in kt:
            us, n, _ = kt[name]
            kernels.append(dict(kernel=name, bound='latency', us_per_launch_in_graph=us / n, launches=n,
                                share_of_batch=round(us / total_us, 4),
                                model='dependent-stage chain: neither HBM- nor tensor-bound (DESIGN.md §4)'))
    # ---- the whole decode step: minimum traffic of an incremental decoder over the valid keys (SURVEY.md §8d terms)
    Lx = 4
    step_bytes = (Lx * 2 * (valid0 + valid1) * H * esz            # X: cross-attention K|V, both stacks
                  + (valid0 + valid1) * (H * esz + 16)            # A: Uk.mem + gate-projected keys
                  + 2 * R * V * 4                                  # D: logits tile written once, read once
                  + V * H * esz                                    # Wv
                  + R * 2 * 8 * (T / 2.0) * H * esz                # K: self-attention KV history, mean t = T/2
                  + 70 * H * H * esz                               # Wl: layer + attention + gen.0 weights, once
                  + R * (S0 + S1) * 12 + B * (S0 + S1) * 4)        # C: copy scatter
    step = dict(bound='hbm', algorithmic_bytes_per_decode_step=step_bytes, ms_per_decode_step=ms_per_decode_step,
                achieved=step_bytes / (ms_per_decode_step * 1e-3) / 1e9, peak=hbm, unit='GB/s')
    step['frac'] = step['achieved'] / hbm
    top = next((k for k in kernels if k['kernel'] == 'cross_attn_part_kernel'), None)
    if top is None:
        return dict(kernel=None, bound='hbm', achieved=None, peak=hbm, unit='GB/s', frac=None, traffic=None,
                    peak_source=which, kernels=kernels, step=step, step_shares=shares)
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at c2 from the committed `ncu --set full` capture
    # (profiles/r2_top_kernels_ncu_summary.txt, same in round 1: 114.3 MB read + 3-6 MB written per launch); other shapes:
    # not captured
    traffic = 117.4e6 if (B, W, S1) == (64, 4, 2560) else None
    return dict(kernel='cross_attn_part_kernel (passage memory, valid keys; timed in the decode graph)', bound='hbm',
                achieved=top['achieved'], peak=hbm, unit='GB/s', frac=top['frac'], traffic=traffic, peak_source=which,
                algorithmic_bytes_per_launch=top['work_per_launch'], us_per_launch=top['us_per_launch_in_graph'],
                launches_per_decode_step=4, kernels=kernels, step=step, step_shares=shares)


# ----------------------------------------------------------------------------- cpu baseline + parity of the sample
def cpu_baseline_and_parity(args, cfg, wl, torch):
    """The reference arm on a bounded sample (about 20 s of CPU work) in a subprocess without CUDA, and - as a by-product
    - parity of THIS run's answers on the sample: the GPU path decodes the same queries (same seeds, same weights) and its
    answers are compared with the reference's, query by query; differing answers are scored with the oracle (key =
    cum_cost / length of the answer under the oracle's fp32 model) so that ties at the storage precision can be told from
    misses (tests/parity_tools.py has the full-size version of this check)."""
    budget = args.cpu_budget if args.cpu_budget else 20.0
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--config', args.config, '--steps', '1',
           '--warmup', '0', '--cpu-budget', str(budget)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, CUDA_VISIBLE_DEVICES='', RANK='0'))
        j = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:   # report, never hide
        return dict(value=None, unit='tokens/s', cores=os.cpu_count(), kind='reference', sample=f'failed: {e!r}'), None
    cpu = j['cpu_baseline']
    parity = None
    try:
        parity = sample_parity(args, cfg, wl, torch, j)
    except Exception as e:
        parity = dict(unavailable=repr(e)[:200])
    return cpu, parity


def sample_parity(args, cfg, wl, torch, ref_line):
    from case_rg_b200 import synthetic as syn
    nq, Ts = int(ref_line['sample_queries']), int(ref_line['sample_steps'])
    want = torch.tensor(ref_line['sample_tokens'], dtype=torch.int64)
    dev = wl.dev
    if wl.family == 'gttp':
        inp = syn.make_gttp_inputs(ISEED, nq, cfg['Lq'], cfg['NP'], cfg['Lp'], wl.V, H)
        data = {k: getattr(inp, k).to(dev) for k in GTTP_KEYS}
    else:
        inp = syn.make_case_inputs(ISEED, nq, cfg['Lq'], cfg['NP'], cfg['Lp'], wl.V, H)
        data = {k: getattr(inp, k).to(dev) for k in CASE_KEYS}
    got = wl.model.fast_search(data, Ts, wl.W, wl.mode).cpu()
    Lm = max(got.size(1), want.size(1))
    pad = lambda t: torch.cat([t, torch.zeros(t.size(0), Lm - t.size(1), dtype=torch.int64)], 1)
    got, want = pad(got), pad(want)
    same = (got == want).all(1)
    res = dict(against='the reference arm\'s answers on its sample queries (same seeds and weights)', queries=nq,
               decode_steps=Ts, storage=args.dtype, identical=int(same.sum()),
               token_agreement=float((got == want).float().mean()),
               note='xavier random-init weights give a near-flat 30,522-way softmax, so free-running searches leave the fp32 '
                    'trajectory at the first storage-precision tie; tests/test_gpu_search_parity.py holds the full-size '
                    'parity cases on peaked weights (fp32 storage: identical; bf16 storage: zero misses under the oracle-'
                    'scored tie criterion, profiles/r2_parity.md)')
    return res


if __name__ == '__main__':
    main()

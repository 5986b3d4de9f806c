#!/usr/bin/env python
"""Benchmark of the CaSE answer-decode hot path (BASELINE.json: answer tokens/s, beam 4).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU implementation (oracle port)

A "step" = one pass of the hot path over one batch: per-batch prefill (memory K/V + Uk.mem
projections) followed by T=40 decode steps of beam 4 over B=64 queries with 10 x 256-token
passages and a 60-token query context, V=30522 (BASELINE.json configs[1]).  Random-init weights
(xavier, common/CumulativeTrainer.py:13-24) and synthetic CAsT-shaped inputs; every rank decodes
its own batch (weak scaling, no data-path collective; one result gather at the end).

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` goes through
the public API (generations.beam over FastCaSE) from pinned HOST buffers, H2D and the D2H of the
answers inside the timed region.
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

WORKLOAD = dict(B=64, W=4, Lq=60, NP=10, Lp=256, T=40, V=30522, H=256)
WSEED, ISEED = 123456, 20211


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--vocab-impl', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-pdl', action='store_true', help='disable programmatic dependent launch (A/B)')
    ap.add_argument('--no-chain', action='store_true', help='row-block layer kernels instead of the cluster kernels (A/B)')
    ap.add_argument('--no-fork', action='store_true', help='no side stream for the additive attentions (A/B)')
    ap.add_argument('--no-post', action='store_true', help='row_linear launches instead of post linears (A/B)')
    ap.add_argument('--no-stack', action='store_true', help='no fused first stack (A/B)')
    ap.add_argument('--no-gate', action='store_true', help='context form of the additive attentions (A/B)')
    ap.add_argument('--next-prefetch', type=int, default=None, help='L2 warm-up mask for the next step (A/B; 0 off, 1 weights + query K|V, 3 + history)')
    ap.add_argument('--no-evict-first', action='store_true', help='no L2 evict-first policy on the K|V / Uk.mem streams (A/B)')
    ap.add_argument('--gate-f16', action='store_true', help='f16 / tensor-core form of the gate kernel (A/B; slower)')
    ap.add_argument('--xattn-ctas', type=int, default=None, help='grid of the passage cross-attention (default: one CTA per SM)')
    ap.add_argument('--xattn-next', type=int, default=None, help='tiles per warp the passage cross-attention prefetches into L2 for the next layer (A/B)')
    ap.add_argument('--kv-prefetch', type=int, default=None, help='percent of the next K|V stream prefetched into L2 by the cluster launches (A/B)')
    ap.add_argument('--streams', type=int, default=1, help='batch slices decoded concurrently on their own streams')
    ap.add_argument('--batch', type=int, default=WORKLOAD['B'])
    ap.add_argument('--beam', type=int, default=WORKLOAD['W'])
    ap.add_argument('--profile', type=int, default=0,
                    help='ncu helper: run one eager batch and bracket N decode steps (from t=20) with '
                         'cudaProfilerStart/Stop, then exit (use ncu --profile-from-start off)')
    return ap.parse_args()


def config_dict(args, extra=None):
    c = dict(workload='CaSE beam-4 decode, batch 64, 10 passages x 256 tokens, Lq=60, T=40, V=30522 '
                      '(BASELINE.json configs[1])',
             queries_per_gpu=args.batch, beam=args.beam, passages=WORKLOAD['NP'], passage_len=WORKLOAD['Lp'],
             query_len=WORKLOAD['Lq'], max_target_length=WORKLOAD['T'], vocab=WORKLOAD['V'],
             parallelism=f'queries sharded x{args.gpus}, no data-path collective',
             l2='per-step K/V + Uk.mem streams (~0.6 GB of valid keys) exceed the 126 MB L2; no explicit flush')
    if extra:
        c.update(extra)
    return c


# ----------------------------------------------------------------------------- reference / cpu baseline
def cpu_reference_sample(n_queries, T, width, threads):
    """Reference algorithm on the host: whole prefix recomputed each step, dense one-hot copy bmm,
    Python beam bookkeeping (oracle port of CaSE/Model.py:94-123 + Generations.py:112-190)."""
    import torch
    from case_rg_b200 import synthetic as syn
    from oracle.case_decoder import CaseOracle
    from oracle import generations as OG
    torch.set_num_threads(threads)
    w = WORKLOAD
    sd = syn.make_case_decoder_state(WSEED, w['V'], w['H'])
    inp = syn.make_case_inputs(ISEED, n_queries, w['Lq'], w['NP'], w['Lp'], w['V'], w['H'])
    orc = CaseOracle(sd)

    def run(Tn):
        t0 = time.perf_counter()
        with torch.no_grad():
            out = OG.beam(orc.stepper(inp, dense_onehot=True), Tn, width)
        return time.perf_counter() - t0, int((out != 0).sum()) if Tn else 0, out
    return run


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ['CUDA_VISIBLE_DEVICES'] = ''
    import warnings
    warnings.filterwarnings('ignore')
    threads = os.cpu_count() or 1
    run = cpu_reference_sample(1, WORKLOAD['T'], args.beam, threads)
    # calibrate: budget ~150 s for warmup+steps -> choose the number of decode steps per sample
    t_probe, _, _ = run(2)
    per_step = max(t_probe / 2, 1e-3)
    budget = 150.0 / max(1, args.steps + args.warmup)
    T = WORKLOAD['T']
    while T > 2 and per_step * T * (1 + 0.02 * T) > budget:
        T -= 1
    for _ in range(args.warmup):
        run(T)
    t0 = time.perf_counter()
    toks = 0
    for _ in range(args.steps):
        _, _, out = run(T)
        toks += out.numel()
    dt = time.perf_counter() - t0
    val = toks / dt
    sample = f'1 query x beam {args.beam} x {T} decode steps per bench step (of T=40), dense one-hot, prefix recompute'
    line = dict(metric='answer_tokens_per_s', value=val, unit='tokens/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=config_dict(args, dict(sample=sample)),
                cpu_baseline=dict(value=val, unit='tokens/s', cores=threads, kind='port', sample=sample),
                e2e=dict(value=val, unit='tokens/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


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


# ----------------------------------------------------------------------------- main arm
def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from case_rg_b200 import synthetic as syn, _lib as L
    from case_rg_b200 import generations as FG
    from case_rg_b200.distributed import gather_answers

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
    w = WORKLOAD
    B, W, T, V = args.batch, args.beam, w['T'], w['V']
    if args.no_pdl:
        L.load().case_set_pdl(0)
    if args.no_chain:
        L.load().case_set_chain(0)
    if args.no_fork:
        L.load().case_set_fork(0)
    if args.no_post:
        L.load().case_set_post_linears(0)
    if args.no_stack:
        L.load().case_set_stack_fusion(0)
    if args.no_gate:
        L.load().case_set_gate_form(0)
    if args.next_prefetch is not None:
        L.load().case_set_next_step_prefetch(args.next_prefetch)
    if args.no_evict_first:
        L.load().case_set_stream_evict_first(0)
    if args.gate_f16:
        L.load().case_set_gate_f16(1)
    if args.xattn_ctas is not None:
        L.load().case_set_xattn_ctas(args.xattn_ctas)
    if args.xattn_next is not None:
        L.load().case_set_xattn_next_prefetch(args.xattn_next)
    if args.kv_prefetch is not None:
        L.load().case_set_kv_prefetch(args.kv_prefetch)
    if args.profile:
        args.streams = 1

    sd = syn.make_case_decoder_state(WSEED, V, w['H'])
    vocab_impl = args.vocab_impl if args.vocab_impl is not None else (1 if args.dtype == 'bf16' else 0)
    model = FG.FastCaSE(sd, device=dev, dtype=args.dtype, max_dec_len=T, beam_width=W, vocab_impl=vocab_impl,
                        use_graph=not args.no_graph, streams=args.streams)
    host = syn.make_case_inputs(ISEED + rank, B, w['Lq'], w['NP'], w['Lp'], V, w['H'], id_base=rank * B).pin()
    d = host.to(dev)
    data_dev = dict(mem_q=d.mem_q, mem_p=d.mem_p, query=d.query, passage=d.passage, prior_q=d.prior_q,
                    prior_p=d.prior_p, answer_rep=d.answer_rep, source_map=d.source_map)
    host_data = dict(mem_q=host.mem_q, mem_p=host.mem_p, query=host.query, passage=host.passage,
                     prior_q=host.prior_q, prior_p=host.prior_p, answer_rep=host.answer_rep,
                     source_map=host.source_map)
    mode = L.MODE_BEAM

    def step_resident():
        return model.fast_search(data_dev, T, W, mode)

    if args.profile:
        import ctypes as C
        model.use_graph = False
        step_resident()
        eng = model.last_engine
        eng.state.reset()
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

    def run_e2e(n):
        """n batches through the public streaming API: pinned host tensors in, host answers out; the copy
        of batch i+1 overlaps the decode of batch i (generations.beam_batches)."""
        out = None
        for out in FG.beam_batches(model, (host_data for _ in range(n)), None, T, W):
            pass
        return out

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

    for _ in range(max(args.warmup, 3)):
        out = step_resident()
    torch.cuda.synchronize(dev)
    run_e2e(2)
    # answer tokens of one step: best-sequence lengths (EOS kept), per Generations.py:188
    eng = model.last_engine
    tokens_per_step = eng.answer_tokens()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, wall, out = timed(step_resident, args.steps)
    ms_e2e, wall_e2e, out_e = timed(lambda: run_e2e(args.steps), 1)
    # the PCIe rate this box gives the pinned input buffers (explains e2e vs value when the link is slow/shared)
    stage_dev = {k: torch.empty_like(v, device=dev) for k, v in host_data.items()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for k, v in host_data.items():
        stage_dev[k].copy_(v, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    h2d_ms = e0.elapsed_time(e1)
    del stage_dev
    clk = clocks.stop() if rank == 0 else None

    # where the step goes: prefill (once per batch) vs the T decode steps, timed separately
    def _ev(fn, n=3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n
    prefill_ms = _ev(lambda: eng.prefill(d.mem_q, d.mem_p, d.query.ne(0), d.passage.ne(0), d.prior_q, d.prior_p,
                                         d.answer_rep, d.source_map))
    decode_ms = _ev(lambda: eng.decode(T, mode, use_graph=not args.no_graph))

    # result gather (the path's only exchange): ids + answers to every rank
    ids, answers = gather_answers(d.ids, out, T, world * B)

    value = world * tokens_per_step * args.steps / (ms * 1e-3)
    e2e_value = world * tokens_per_step * args.steps / (ms_e2e * 1e-3)
    h2d = host.nbytes()
    d2h = int(out_e.numel() * out_e.element_size())

    roof = cpu = shares = None
    if rank == 0:
        roof = roofline(model, eng, args, torch)
        shares = step_shares(step_resident, torch, dev)
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline(args)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    launches = args.steps * (T * (eng.kernel_launches_per_step() + 1) + 40)
    line = dict(metric='answer_tokens_per_s', value=value, unit='tokens/s', n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype=args.dtype if args.dtype == 'bf16' else 'f32', data='synthetic',
                config=config_dict(args, dict(cuda_graph=not args.no_graph, pdl=not args.no_pdl, vocab_gemm='tcgen05' if vocab_impl == 1 else 'simt',
                                              layer_kernels='row-block' if args.no_chain else 'cluster', streams=args.streams)),
                ms_per_decode_step=decode_ms / T, prefill_ms=prefill_ms, decode_ms=decode_ms,
                answer_tokens_per_step=tokens_per_step,
                e2e=dict(value=e2e_value, unit='tokens/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=ms_e2e / args.steps, h2d_ms_alone=h2d_ms, h2d_gbs_alone=h2d / (h2d_ms * 1e6),
                         pinned=all(v.is_pinned() for v in host_data.values())),
                gpu_launches=launches, clocks=clk, roofline=roof, step_shares=shares, cpu_baseline=cpu,
                wall_s=dict(resident=wall, e2e=wall_e2e))
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def step_shares(step_fn, torch, dev):
    """Share of every kernel in the GPU time of one batch (prefill + decode graph), from the CUPTI activity records of
    one extra, untimed batch: the figure the committed ncu launch list (profiles/) has to agree with.  The largest
    share is the latency-bound row work of the cluster launches (layer_chain_kernel); the `roofline` object describes
    the largest BANDWIDTH-bound kernel, the passage cross-attention."""
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step_fn()
            torch.cuda.synchronize(dev)
        tot = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA and 'Memcpy' not in e.name and 'Memset' not in e.name:
                k = e.name.split('(')[0].split('<')[0].replace('void ', '').replace('cb::', '')
                tot[k] = tot.get(k, 0.0) + (e.time_range.end - e.time_range.start)
        s = sum(tot.values())
        top = sorted(tot.items(), key=lambda kv: -kv[1])[:8]
        return {k: round(v / s, 4) for k, v in top} if s > 0 else None
    except Exception as ex:          # the profiler is evidence, not the product: never fail the bench line over it
        return {'unavailable': str(ex)[:120]}


def roofline(model, eng, args, torch):
    """Dominant HBM-bound kernel = the passage-memory cross-attention (4 launches per decode step):
    algorithmic bytes per launch = B * 2 * S1 * H * sizeof(storage), duration = CUDA
    events around back-to-back launches over the 4 layers' distinct K/V (688 MB > L2)."""
    from case_rg_b200 import _lib as L
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak, which = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650.0, 'fallback')
    eng = eng.subs[0] if hasattr(eng, 'subs') else eng    # the launch shape of one stream's slice
    B, W, S1 = eng.B, eng.W, eng.S[1]
    esz = 2 if eng.w.cdtype == L.BF16 else 4
    compact = bool(getattr(eng, 'compact', False))
    # algorithmic bytes: K and V of every key the attention needs, once per query;
    # with the compacted memory that is the VALID keys only (padding is dropped at prefill)
    nkeys = int(eng.xcount.sum().item()) if compact else B * S1
    alg = nkeys * 2 * L.H * esz
    st = torch.cuda.current_stream()
    reps = 20
    def launch(l):
        if compact:
            L.call('case_cross_attn_part', eng.q2.data_ptr(), eng.Kx[l].data_ptr(), eng.xcount.data_ptr(),
                   eng.xprefix.data_ptr(), B, W, S1, eng.xslots, eng.part_ml.data_ptr(), eng.part_acc.data_ptr(),
                   st.cuda_stream)
        elif eng.w.cdtype == L.BF16:
            L.call('case_cross_attn_partial_tc', eng.q2.data_ptr(), eng.Kx[l].data_ptr(), eng.mask[1].data_ptr(), B, W,
                   S1, eng.nsx[1], eng.part_ml.data_ptr(), eng.part_acc.data_ptr(), st.cuda_stream)
        else:
            L.call('case_cross_attn_partial', eng.q2.data_ptr(), eng.Kx[l].data_ptr(), eng.Vx[l].data_ptr(),
                   eng.mask[1].data_ptr(), B, W, S1, eng.nsx[1], eng.part_ml.data_ptr(), eng.part_acc.data_ptr(),
                   eng.w.cdtype, st.cuda_stream)
    for l in range(4, 8):
        launch(l)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for l in range(4, 8):
            launch(l)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * 4)
    ach = alg / (us * 1e-6) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at the BASELINE shape from the committed
    # `ncu --set full` capture (profiles/r1_top_kernels_ncu_summary.txt): 114.3 MB read + 3.1 MB written
    traffic = 117.4e6 if (compact and (B, W, S1) == (64, 4, 2560)) else None
    kname = 'cross_attn_part_kernel (passage memory, valid keys)' if compact else ('cross_attn_mma_kernel (passage memory)' if eng.w.cdtype == L.BF16 else 'cross_attn_partial_kernel (passage memory)')
    return dict(kernel=kname, bound='hbm', achieved=ach, peak=peak, unit='GB/s',
                frac=ach / peak, traffic=traffic, peak_source=which, algorithmic_bytes_per_launch=alg,
                us_per_launch=us, launches_per_decode_step=4)


def cpu_baseline(args):
    """Oracle port of the reference algorithm on this box's host cores, bounded sample (rank 0, N=1)."""
    code = ('import sys,json,os; sys.path.insert(0, %r); os.environ["CUDA_VISIBLE_DEVICES"]="";'
            'import warnings; warnings.filterwarnings("ignore");'
            'import bench; run = bench.cpu_reference_sample(1, 40, %d, os.cpu_count());'
            'dt, _, out = run(%d); print(json.dumps(dict(dt=dt, toks=int(out.numel()))))') % (ROOT, args.beam, 24)
    try:
        r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, CUDA_VISIBLE_DEVICES=''))
        j = json.loads(r.stdout.strip().splitlines()[-1])
        return dict(value=j['toks'] / j['dt'], unit='tokens/s', cores=os.cpu_count(), kind='port',
                    sample=f'1 query x beam {args.beam} x 24 decode steps (of 40), dense one-hot + prefix recompute, '
                           f'{j["dt"]:.1f} s of CPU work')
    except Exception as e:   # report, never hide
        return dict(value=None, unit='tokens/s', cores=os.cpu_count(), kind='port', sample=f'failed: {e!r}')


if __name__ == '__main__':
    main()

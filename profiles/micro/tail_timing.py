#!/usr/bin/env python
"""Where the time of case_row_tail goes: CUDA-event timing of the launch with phases switched off.
usage: python profiles/micro/tail_timing.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import _lib as L           # noqa: E402


def build(R=256, W=4, V=30522, S=(60, 2560), ns=(1, 10), K=4, finalize=1, dist=False, topk=True):
    dev, f32 = 'cuda', dict(dtype=torch.float32, device='cuda')
    B, H, MS = R // W, 256, L.MAX_SPLIT
    ldv = -(-V // 8) * 8
    keep = dict(logits=torch.randn(R, ldv, **f32) * 3, hN=torch.randn(R, H, **f32),
                Wm=torch.randn(3, 3 * H, **f32) * 0.05, bm=torch.randn(3, **f32),
                stats=[torch.rand(R, n, 4, **f32) + 0.1 for n in ns], ctxp=[torch.randn(R, n, H, **f32) for n in ns],
                attn=[torch.randn(R, s, **f32) for s in S], prior=[torch.rand(B, s, **f32) for s in S],
                smap=torch.randint(0, V, (B, sum(S)), device=dev, dtype=torch.int32),
                ctx=[torch.zeros(R, H, **f32) for _ in range(2)], gates=torch.rand(R, 4, **f32),
                fac=torch.rand(R, 2, MS, **f32), tv=torch.zeros(R, K, **f32),
                ti=torch.zeros(R, K, dtype=torch.int32, device=dev), dist=torch.zeros(R, ldv, **f32))
    a = L.TailArgs()
    a.R, a.V, a.W, a.K, a.ldl, a.ldd, a.mask_col0, a.nmem, a.do_finalize = R, V, W, K, ldv, ldv, 0, 2, finalize
    a.fac_ld, a.map_ld = 2 * MS, sum(S)
    for i in range(2):
        a.ns[i], a.fac_off[i], a.map_off[i], a.S[i] = ns[i], i * MS, (0, S[0])[i], S[i]
        a.stats[i], a.ctxp[i], a.ctx[i] = keep['stats'][i].data_ptr(), keep['ctxp'][i].data_ptr(), keep['ctx'][i].data_ptr()
        a.prior[i], a.attn_un[i] = keep['prior'][i].data_ptr(), keep['attn'][i].data_ptr()
    a.logits, a.hN, a.Wm, a.bm = keep['logits'].data_ptr(), keep['hN'].data_ptr(), keep['Wm'].data_ptr(), keep['bm'].data_ptr()
    a.gates, a.fac, a.map = keep['gates'].data_ptr(), keep['fac'].data_ptr(), keep['smap'].data_ptr()
    if topk:
        a.top_vals, a.top_idx = keep['tv'].data_ptr(), keep['ti'].data_ptr()
    if dist or not topk:
        a.dist = keep['dist'].data_ptr()
    return a, keep


def time_it(name, **kw):
    a, keep = build(**kw)
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        L.check(lib.case_row_tail(C.byref(a), st), 'tail')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.case_row_tail(C.byref(a), st)
    e1.record()
    torch.cuda.synchronize()
    print(f'{name:40s} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us')


def stamps(**kw):
    a, keep = build(**kw)
    lib = L.load()
    lib.case_debug_tail_timing.argtypes = [C.c_void_p]
    dbg = torch.zeros(32, dtype=torch.int64, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    names = ['pdl_wait', 'finalize', 'logits landed', 'max+exp+sum', 'cluster sync 1', 'scale', 'scatter', 'topk scan',
             'topk rounds', 'cluster sync 2']
    for rep in range(2):
        dbg.zero_()
        lib.case_debug_tail_timing(dbg.data_ptr())
        L.check(lib.case_row_tail(C.byref(a), st), 'tail')
        torch.cuda.synchronize()
        lib.case_debug_tail_timing(None)
        s = dbg.cpu().tolist()
        print('rep', rep, 'total', s[len(names)] - s[0])
        for i, n in enumerate(names):
            print(f'   {n:16s} {s[i + 1] - s[i]:7d} cyc')


def sparse_stamps(R=256, W=4, V=30522, S=(60, 2560), K=4, gate=True):
    a, keep = build(R=R, W=W, V=V, S=S, K=K, ns=(2, 16) if gate else (1, 10))
    if gate:                               # production form: gate partials [R][ns][4] instead of context partials
        keep['gp'] = [torch.randn(R, n, 4, dtype=torch.float32, device='cuda') for n in (2, 16)]
        for i in range(2):
            a.ctxp[i] = keep['gp'][i].data_ptr()
        a.gate_ctx = 1
    lib = L.load()
    lib.case_debug_sparse_tail_timing.argtypes = [C.c_void_p]
    f32 = dict(dtype=torch.float32, device='cuda')
    k2 = 2 * K
    ldv = -(-V // 8) * 8
    bms, bl = torch.zeros(R, 4, 2, **f32), torch.zeros(R, 4, k2, **f32)
    bi = torch.zeros(R, 4, k2, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    dbg = torch.zeros(128, dtype=torch.int64, device='cuda')
    names = ['init+pdl wait', 'finalize', 'hash inserts', 'pass A (gather)', 'threshold+append', 'select']

    def base():
        L.call('case_vocab_base', keep['logits'].data_ptr(), ldv, R, V, 0, k2, bms.data_ptr(), bl.data_ptr(), bi.data_ptr(), st)

    def tail():
        L.check(lib.case_sparse_tail(C.byref(a), bms.data_ptr(), bl.data_ptr(), bi.data_ptr(), k2, None, None, st), 'sparse')
    base()
    for rep in range(2):
        dbg.zero_()
        lib.case_debug_sparse_tail_timing(dbg.data_ptr())
        tail()
        torch.cuda.synchronize()
        lib.case_debug_sparse_tail_timing(None)
        s = dbg.cpu().tolist()
        print('sparse rep', rep, 'total', s[len(names)] - s[0])
        for i, n in enumerate(names):
            print(f'   {n:16s} {s[i + 1] - s[i]:7d} cyc')
        for w in range(8):
            ws = s[32 + w * 8: 32 + w * 8 + 6]
            print(f'   warp {w}: start+{ws[0] - s[1]:6d}  mem0 {ws[1] - ws[0]:6d}  mem1 {ws[2] - ws[1]:6d}  sums {ws[3] - ws[2]:6d}  barrier {ws[4] - ws[3]:6d}  gates {ws[5] - ws[4]:6d}')
    for name, fn in (('case_vocab_base', base), ('case_sparse_tail', tail)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f'{name:40s} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us')


if __name__ == '__main__':
    sparse_stamps()
    sys.exit(0)
    stamps()
    time_it('full (finalize, 2 scatters, top-4)')
    time_it('no finalize', finalize=0)
    time_it('no finalize, tiny scatter', finalize=0, S=(4, 4))
    time_it('no finalize, tiny scatter, top-1', finalize=0, S=(4, 4), K=1)
    time_it('no finalize, tiny scatter, dist only', finalize=0, S=(4, 4), topk=False)
    time_it('finalize, tiny scatter, top-1', finalize=1, S=(4, 4), K=1)
    time_it('full, R=64', R=64)
    time_it('full, R=128', R=128)

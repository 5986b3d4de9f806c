#!/usr/bin/env python
"""Stage-by-stage clock64() stamps of one case_layer_chain launch (CTA 0) at the bench shape, and a
row-level diff of the chain path against the row-block path.  usage: python profiles/micro/chain_timing.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import _lib as L, synthetic as syn           # noqa: E402
from case_rg_b200.generations import FastCaSE                   # noqa: E402

NAMES = ['start->resident', 'pdl_wait', 'B0 merge+bcast', 'B1 Wo2', 'B2 LN3', 'B3 W1', 'B4 W2', 'F0 LN1', 'F1 QKV',
         'F2 self-attn', 'F3 Wo', 'F4 LN2', 'F5 Wq2']


def timing():
    V, B, W, T = 30522, 64, 4, 40
    sd = syn.make_case_decoder_state(123456, V, 256)
    inp = syn.make_case_inputs(20211, B, 60, 10, 256, V, 256).to('cuda')
    data = dict(mem_q=inp.mem_q, mem_p=inp.mem_p, query=inp.query, passage=inp.passage, prior_q=inp.prior_q,
                prior_p=inp.prior_p, answer_rep=inp.answer_rep, source_map=inp.source_map)
    model = FastCaSE(sd, device='cuda', dtype='bf16', use_graph=False)
    model.fast_search(data, T, W, L.MODE_BEAM)
    eng = model.last_engine
    lib = L.load()
    lib.case_debug_chain_timing.argtypes = [C.c_void_p]
    dbg = torch.zeros(64, dtype=torch.int64, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    a = eng.args
    t = 20
    for rep in range(3):
        dbg.zero_()
        lib.case_debug_chain_timing(dbg.data_ptr())
        L.check(lib.case_layer_chain(C.byref(a.layers[4]), C.byref(a.layers[5]), None, None, None, 16.0, None, a.bbuf,
                                     a.part_ml, a.part_acc, (eng.xslots or a.nsplit_x[1]), a.h, a.kcache[5], a.vcache[5], a.anc[t & 1],
                                     T + 1, a.tok, T + 1, a.prow, t, T, a.bbuf, a.q2, eng.R, 0, None, st), 'chain')
        torch.cuda.synchronize()
        lib.case_debug_chain_timing(None)
        s = dbg.cpu().tolist()
        d = [s[i + 1] - s[i] for i in range(13)]
        print(f'rep {rep}: total {s[13] - s[0]} cycles')
        for n, x in zip(NAMES, d):
            print(f'   {n:18s} {x:7d} cyc  {x / 1.965e3:6.2f} us')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        lib.case_layer_chain(C.byref(a.layers[4]), C.byref(a.layers[5]), None, None, None, 16.0, None, a.bbuf,
                             a.part_ml, a.part_acc, (eng.xslots or a.nsplit_x[1]), a.h, a.kcache[5], a.vcache[5], a.anc[t & 1],
                             T + 1, a.tok, T + 1, a.prow, t, T, a.bbuf, a.q2, eng.R, 0, None, st)
    e1.record()
    torch.cuda.synchronize()
    print(f'back-to-back launches: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us each')


def rowdiff():
    V, B, W, T = 3000, 5, 4, 9
    sd = syn.make_case_decoder_state(51, V, 256, peaked=0.3, boost={syn.EOS: 6.0}, gen_gate_bias=2.0)
    inp = syn.make_case_inputs(52, B, 20, 3, 40, V, 256).to('cuda')
    data = dict(mem_q=inp.mem_q, mem_p=inp.mem_p, query=inp.query, passage=inp.passage, prior_q=inp.prior_q,
                prior_p=inp.prior_p, answer_rep=inp.answer_rep, source_map=inp.source_map)
    models = {c: FastCaSE(sd, device='cuda', dtype='bf16', use_graph=False, opt=0 if c else L.OPT_NO_CHAIN) for c in (0, 1)}
    for T_ in (1, 2, 3, 9):
        res = {}
        for chain in (0, 1):
            model = models[chain]
            model.fast_search(data, T_, W, L.MODE_BEAM)
            eng = model.last_engine
            torch.cuda.synchronize()
            res[chain] = (eng.h.clone(), eng.state.live.clone(), eng.state.tok.clone(), eng.q2.clone())
        d = (res[0][0] - res[1][0]).abs().amax(1)
        dq = (res[0][3] - res[1][3]).abs().amax(1)
        print(f'T={T_} row max|dh|:', [f'{x:.3f}' for x in d.tolist()])
        print(f'      row max|dq2|:', [f'{x:.3f}' for x in dq.tolist()])
        print('      live      :', res[1][1].tolist(), ' tok equal:', bool(torch.equal(res[0][2], res[1][2])))


def occupancy():
    lib = L.load()
    for smem in (100000, 211000):
        print('smem', smem, {c: lib.case_debug_chain_max_clusters(smem, c) for c in (1, 2, 4, 8)})


def vs_fp32():
    V, B, W, T = 3000, 5, 4, 1
    sd = syn.make_case_decoder_state(51, V, 256, peaked=0.3, boost={syn.EOS: 6.0}, gen_gate_bias=2.0)
    inp = syn.make_case_inputs(52, B, 20, 3, 40, V, 256).to('cuda')
    data = dict(mem_q=inp.mem_q, mem_p=inp.mem_p, query=inp.query, passage=inp.passage, prior_q=inp.prior_q,
                prior_p=inp.prior_p, answer_rep=inp.answer_rep, source_map=inp.source_map)
    lib = L.load()
    m32 = FastCaSE(sd, device='cuda', dtype='fp32', use_graph=False)
    m32.fast_search(data, T, W, L.MODE_BEAM)
    ref = m32.last_engine.h.clone()
    for chain in (0, 1):
        model = FastCaSE(sd, device='cuda', dtype='bf16', use_graph=False, opt=0 if chain else L.OPT_NO_CHAIN)
        model.fast_search(data, T, W, L.MODE_BEAM)
        torch.cuda.synchronize()
        d = (model.last_engine.h - ref).abs().amax(1)
        print(f'chain={chain} row max|h - h_fp32|:', [f'{x:.3f}' for x in d.tolist()])


if __name__ == '__main__':
    occupancy()
    vs_fp32()
    rowdiff()
    timing()

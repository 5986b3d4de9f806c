#!/usr/bin/env python
"""Stage-by-stage clock64() stamps of one case_layer_stack launch (CTA 0) at the bench shape.
usage: python profiles/micro/stack_timing.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import _lib as L, synthetic as syn           # noqa: E402
from case_rg_b200.generations import FastCaSE                   # noqa: E402

FRONT = ['F0 wait h', 'F0 LN1', 'F1 QKV', 'F2 self-attn', 'F3 Wo', 'F4 LN2', 'F5 Wq2']
BACK = ['X xattn+B0', 'B1 Wo2', 'B2 LN3', 'B3 W1', 'B4 W2']


def main():
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
    kcs = (C.c_void_p * 5)(*[a.kcache[l] for l in range(5)])
    vcs = (C.c_void_p * 5)(*[a.vcache[l] for l in range(5)])
    kxs = (C.c_void_p * 4)(*[a.Kx[l] for l in range(4)])

    def launch():
        return lib.case_layer_stack(a.layers, 4, kcs, vcs, kxs, a.mask[0], W, 60, None, a.E, a.pe, 16.0, a.x_in, a.h0,
                                    a.anc[t & 1], T + 1, a.tok, T + 1, a.prow, t, T, a.bbuf, a.q2, eng.R, 1, None, st)
    names = ['start->resident', 'pdl_wait']
    for f in range(5):
        names += [f'L{f} {n}' for n in FRONT]
        if f < 4:
            names += [f'L{f} {n}' for n in BACK]
    for rep in range(2):
        dbg.zero_()
        lib.case_debug_chain_timing(dbg.data_ptr())
        L.check(launch(), 'stack')
        torch.cuda.synchronize()
        lib.case_debug_chain_timing(None)
        s = [x for x in dbg.cpu().tolist() if x != 0]
        print(f'rep {rep}: {len(s)} stamps, total {s[-1] - s[0]} cycles = {(s[-1] - s[0]) / 1.965e3:.1f} us')
        if rep == 1:
            for i in range(len(s) - 1):
                n = names[i] if i < len(names) else '?'
                print(f'   {n:18s} {s[i + 1] - s[i]:7d} cyc  {(s[i + 1] - s[i]) / 1.965e3:6.2f} us')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        launch()
    e1.record()
    torch.cuda.synchronize()
    print(f'back-to-back launches: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us each')


if __name__ == '__main__':
    main()

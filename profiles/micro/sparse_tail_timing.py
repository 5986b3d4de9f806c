#!/usr/bin/env python
"""clock64() stamps of CTA 0 of sparse_tail_kernel inside one decode step at the bench shape (t = 20).
usage: python profiles/micro/sparse_tail_timing.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import _lib as L, synthetic as syn           # noqa: E402
from case_rg_b200.generations import FastCaSE                   # noqa: E402

NAMES = ['table init + gate weights (before the PDL wait)', 'merge attention stats, gates', 'hash copy mass',
         'gather touched logits, values', 'base candidates', 'block top-k + store', 'search bookkeeping (select)']


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
    lib.case_debug_sparse_tail_timing.argtypes = [C.c_void_p]
    dbg = torch.zeros(128, dtype=torch.int64, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    eng.state.reset()
    eng.args.mode, eng.args.max_len = L.MODE_BEAM, T
    for t in range(21):
        if t == 20:
            lib.case_debug_sparse_tail_timing(dbg.data_ptr())
        eng._step(t, st)
    torch.cuda.synchronize()
    lib.case_debug_sparse_tail_timing(None)
    s = [int(x) for x in dbg[:16].cpu() if int(x) != 0]
    print(f'{len(s)} stamps, total {s[-1] - s[0]} cycles')
    for i in range(len(s) - 1):
        name = NAMES[i] if i < len(NAMES) else '?'
        print(f'   {name:50s} {s[i + 1] - s[i]:7d} cyc  {(s[i + 1] - s[i]) / 1965.0:6.2f} us')


if __name__ == '__main__':
    main()

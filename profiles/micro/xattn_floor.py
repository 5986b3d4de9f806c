#!/usr/bin/env python
"""How much of case_cross_attn_part's time is HBM?  Back-to-back launches at the bench shape
(a) cycling the 4 layers' K|V (4 x 112 MB: every byte from HBM), (b) the same layer every time (112 MB against a
126 MB L2: mostly L2 hits), (c) B = 16 queries (28 MB: all L2).  usage: python profiles/micro/xattn_floor.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import _lib as L, synthetic as syn           # noqa: E402
from case_rg_b200.generations import FastCaSE                   # noqa: E402


def run(B):
    V, W, T = 30522, 4, 40
    sd = syn.make_case_decoder_state(123456, V, 256)
    inp = syn.make_case_inputs(20211, B, 60, 10, 256, V, 256).to('cuda')
    model = FastCaSE(sd, device='cuda', dtype='bf16')
    eng = model.engine_for(B, W, 60, 2560, T)
    eng.prefill(inp.mem_q, inp.mem_p, inp.query.ne(0), inp.passage.ne(0), inp.prior_q, inp.prior_p, inp.answer_rep,
                inp.source_map)
    st = torch.cuda.current_stream()
    nbytes = int(eng.xcount.sum().item()) * 2 * 256 * 2

    def launch(l):
        L.call('case_cross_attn_part', eng.q2.data_ptr(), eng.Kx[l].data_ptr(), eng.xcount.data_ptr(),
               eng.xprefix.data_ptr(), B, W, 2560, eng.xslots, eng.part_ml.data_ptr(), eng.part_acc.data_ptr(),
               st.cuda_stream)

    def timeit(layers, reps=40):
        for l in layers:
            launch(l)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for l in layers:
                launch(l)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * len(layers))
    a, b = timeit([4, 5, 6, 7]), timeit([4])
    print(f'B={B}: {nbytes / 1e6:.1f} MB per launch; cycling 4 layers {a:.2f} us ({nbytes / a / 1e3:.0f} GB/s), '
          f'same layer {b:.2f} us ({nbytes / b / 1e3:.0f} GB/s)')


if __name__ == '__main__':
    for B in (64, 32, 16, 4):
        run(B)

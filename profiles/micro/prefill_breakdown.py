#!/usr/bin/env python
"""Kernel-by-kernel breakdown of one CaseDecodeEngine.prefill at the bench shape (torch profiler, CUDA time).
usage: python profiles/micro/prefill_breakdown.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import _lib as L, synthetic as syn           # noqa: E402
from case_rg_b200.generations import FastCaSE                   # noqa: E402


def main():
    V, B, W, T = 30522, 64, 4, 40
    sd = syn.make_case_decoder_state(123456, V, 256)
    inp = syn.make_case_inputs(20211, B, 60, 10, 256, V, 256).to('cuda')
    model = FastCaSE(sd, device='cuda', dtype='bf16')
    eng = model.engine_for(B, W, 60, 2560, T)
    args = (inp.mem_q, inp.mem_p, inp.query.ne(0), inp.passage.ne(0), inp.prior_q, inp.prior_p, inp.answer_rep,
            inp.source_map)
    for _ in range(3):
        eng.prefill(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.prefill(*args)
    e1.record()
    torch.cuda.synchronize()
    print(f'prefill: {e0.elapsed_time(e1) / 10:.3f} ms per batch')
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        eng.prefill(*args)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=70))


if __name__ == '__main__':
    main()

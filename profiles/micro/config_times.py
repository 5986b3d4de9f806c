#!/usr/bin/env python
"""Per-batch and per-step times of the BASELINE.json configurations on one B200 (prefill + T = 40 decode steps through
the public Generations face, CUDA events, after 2 warm-up batches): C1 (B=8, greedy, 10 x 100), C2 (B=64, beam 4,
10 x 256), the per-GPU share of C5 (B=32, beam 8, 20 x 512) and C4 (GTTP, B=128, beam 4, V=50,000, Lb=1000).
usage: python profiles/micro/config_times.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import generations as FG, synthetic as syn           # noqa: E402

H, T, DEV = 256, 40, 'cuda'


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def case_cfg(name, B, W, NP, Lp, V=syn.BERT_VOCAB):
    sd = syn.make_case_decoder_state(53, V, H)
    d = syn.make_case_inputs(63, B, 60, NP, Lp, V, H).to(DEV)
    data = dict(mem_q=d.mem_q, mem_p=d.mem_p, query=d.query, passage=d.passage, prior_q=d.prior_q, prior_p=d.prior_p,
                answer_rep=d.answer_rep, source_map=d.source_map)
    model = FG.FastCaSE(sd, device=DEV, dtype='bf16')
    fn = (lambda: FG.beam(model, data, None, T, W)) if W > 1 else (lambda: model.module_greedy(data, T))
    ms, out = timed(fn)
    print(f'{name}: B={B} W={W} S={60 + NP * Lp} V={V}: {ms:.2f} ms per batch of {T} steps '
          f'({ms / T * 1e3:.0f} us per step incl. prefill)')


def gttp_cfg(name, B, W, V=50000):
    sd = syn.make_gttp_state(52, V, H, H)
    d = syn.make_gttp_inputs(62, B, 60, 10, 100, V, H).to(DEV)
    data = dict(context=d.context, background=d.background, background_map=d.background_map, src_output=d.src_output,
                bg_output=d.bg_output, init_state=d.init_state)
    model = FG.FastGTTP(sd, device=DEV, dtype='bf16')
    ms, out = timed(lambda: FG.beam(model, data, None, T, W))
    print(f'{name}: GTTP B={B} W={W} Lb=1000 V={V}: {ms:.2f} ms per batch of {T} steps ({ms / T * 1e3:.0f} us per step incl. prefill)')


if __name__ == '__main__':
    case_cfg('C1', 8, 1, 10, 100)
    case_cfg('C2', 64, 4, 10, 256)
    case_cfg('C5/8', 32, 8, 20, 512)
    gttp_cfg('C4', 128, 4)

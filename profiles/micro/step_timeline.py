#!/usr/bin/env python
"""Timeline of the kernels of ONE decode step inside the captured CUDA graph (CUPTI activity records through the
torch profiler): start offset, duration, stream and the gap to the previous kernel's end on the critical path.
usage: python profiles/micro/step_timeline.py [step]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import _lib as L, synthetic as syn           # noqa: E402
from case_rg_b200.generations import FastCaSE                   # noqa: E402


def main():
    step = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    V, B, W, T = 30522, 64, 4, 40
    sd = syn.make_case_decoder_state(123456, V, 256)
    inp = syn.make_case_inputs(20211, B, 60, 10, 256, V, 256).to('cuda')
    data = dict(mem_q=inp.mem_q, mem_p=inp.mem_p, query=inp.query, passage=inp.passage, prior_q=inp.prior_q,
                prior_p=inp.prior_p, answer_rep=inp.answer_rep, source_map=inp.source_map)
    model = FastCaSE(sd, device='cuda', dtype='bf16')
    for _ in range(3):
        model.fast_search(data, T, W, L.MODE_BEAM)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.fast_search(data, T, W, L.MODE_BEAM)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and 'Memcpy' not in e.name
          and 'Memset' not in e.name]
    ev.sort(key=lambda e: e.time_range.start)
    # the decode kernels: from the first layer_chain_kernel on; a step = from one stack launch to the next
    names = [e.name for e in ev]
    stacks = [i for i, n in enumerate(names) if 'layer_chain_kernel' in n]
    # the stack launch of a step is the chain launch that follows a sparse_tail (or the first one)
    starts = [i for i in stacks if i == stacks[0] or 'sparse_tail' in names[i - 1] or 'additive' in names[i - 1]]
    starts = [i for k, i in enumerate(starts) if k == 0 or i - starts[k - 1] > 5]
    print(f'{len(ev)} kernels, {len(starts)} steps found')
    if step + 1 >= len(starts):
        step = len(starts) // 2
    lo, hi = starts[step], starts[step + 1]
    t0 = ev[lo].time_range.start
    prev_end = t0
    print(f'step {step}: {ev[hi].time_range.start - t0:.1f} us from stack launch to stack launch')
    print(f'{"start":>8} {"dur":>7} {"gap":>6} stream  kernel')
    for e in ev[lo:hi]:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        print(f'{s:8.1f} {d:7.1f} {e.time_range.start - prev_end:6.1f} {getattr(e, "device_index", 0)}  {e.name[:60]}')
        prev_end = max(prev_end, e.time_range.end)
    total = {}
    for e in ev[lo:hi]:
        k = e.name.split('(')[0][-40:]
        total[k] = total.get(k, 0) + e.time_range.end - e.time_range.start
    print({k: round(v, 1) for k, v in total.items()})


if __name__ == '__main__':
    main()

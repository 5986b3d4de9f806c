#!/usr/bin/env python
"""Time of the pre-decode producers at a BASELINE shape, total and per kernel (CUPTI records).
usage: python profiles/micro/producers_timing.py [c2|c5|c1] [tc|cublas]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import synthetic as syn                     # noqa: E402
from case_rg_b200.producers import CaseProducers              # noqa: E402

SHAPES = {'c1': (8, 60, 10, 100), 'c2': (64, 60, 10, 256), 'c5': (32, 60, 20, 512)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
    gemm = sys.argv[2] if len(sys.argv) > 2 else 'tc'
    B, Lq, NP, Lp = SHAPES[name]
    V = 30522
    sd = syn.make_case_producer_state(5, V, 256)
    inp = syn.make_case_inputs(6, B, Lq, NP, Lp, V, 256)
    prod = CaseProducers(sd, device='cuda', gemm=gemm)
    q, p = inp.query.cuda(), inp.passage.cuda()
    for _ in range(2):
        prod(q, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        prod(q, p)
    e1.record()
    torch.cuda.synchronize()
    print(f'{name} gemm={gemm}: producers {e0.elapsed_time(e1) / 5:.2f} ms per batch (B={B}, {NP} x {Lp} passages, Lq={Lq})')
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        prod(q, p)
        torch.cuda.synchronize()
    tot = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name.split('(')[0].replace('void ', '').replace('cb::', '')[:60]
            us, n = tot.get(k, (0.0, 0))
            tot[k] = (us + (e.time_range.end - e.time_range.start), n + 1)
    s = sum(v[0] for v in tot.values())
    for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f'   {k:60s} {us / 1e3:8.3f} ms  x{n:4d}  {100 * us / s:5.1f} %')
    # GEMM throughput: flops of every Linear of the pipeline
    Mq, Mp = B * Lq, B * NP * Lp
    H = 256
    enc = 3 * 2 * (3 * H * H + H * H + 2 * H * H)
    blk5 = 2 * (3 * 1280 * 1280 + 1280 * 1280 + 1280 * H + H * H)
    blk1 = 2 * (3 * H * H + H * H + 2 * H * H)
    fl = (Mq + Mp) * enc + Mq * (2 * blk5 + 3 * blk1) + Mp * (2 * blk5 + 6 * blk1)
    gem = sum(us for k, (us, n) in tot.items() if 'gemm' in k.lower() or 'cutlass' in k.lower() or 'nvjet' in k.lower() or 'sm100' in k.lower())
    if gem > 0:
        print(f'   GEMM flops {fl / 1e12:.2f} TFLOP in {gem / 1e3:.2f} ms = {fl / gem / 1e6:.0f} TFLOP/s')


if __name__ == '__main__':
    main()

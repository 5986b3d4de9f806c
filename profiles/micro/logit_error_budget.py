#!/usr/bin/env python
"""Where the bf16 path's logit error comes from: the vocabulary GEMM itself (bf16 operands) or everything upstream of it
(the hidden state after 8 bf16 layers + attentions).  Step 0 of the C2 parity problem (peaked weights), oracle in strict fp32
on the GPU.   usage: python profiles/micro/logit_error_budget.py"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from case_rg_b200 import _lib as L, synthetic as syn, generations as FG     # noqa: E402
import test_gpu_search_parity as TP                                       # noqa: E402
from oracle.case_decoder import CaseOracle, strict_fp32                    # noqa: E402


def main():
    strict_fp32()
    dev = 'cuda:0'
    sd, inp = TP._case_problem(64, 60, 10, 256, 51, 61)
    data = TP._case_data(inp)
    model = FG.FastCaSE(sd, device=dev, dtype='bf16')
    FG.beam(model, data, None, 40, 4)
    eng = model.last_engine
    eng.state.reset()
    eng.args.mode, eng.args.max_len = L.MODE_BEAM, 40
    eng._run_steps(1)
    torch.cuda.synchronize()
    st = CaseOracle(sd, device=dev).incremental(inp)
    st.advance(torch.arange(64), torch.full((64,), syn.BOS))
    ref = st.last['logits']                                  # [64, V] fp32
    got = eng.logits[::4, :eng.V]                            # the engine's logits, BOS rows
    gf = eng.gfeat[::4].float()                              # gen.0 output of the bf16 path (fp32 buffer)
    Wv = sd['gen.2.weight'].to(dev).float()
    exact_tail = gf @ Wv.t()                                 # fp32 vocabulary GEMM on the bf16 path's hidden features
    bf_tail = gf.bfloat16().float() @ Wv.bfloat16().float().t()
    scale = float(ref.abs().max())
    def e(a, b): return float((a - b).abs().max()) / scale, float((a - b).abs().mean()) / scale
    print('max|logit| of the oracle: %.2f' % scale)
    print('engine logits            vs oracle: max %.2e  mean %.2e' % e(got, ref))
    print('fp32 GEMM on engine gfeat vs oracle: max %.2e  mean %.2e   (error of everything upstream of the GEMM)' % e(exact_tail, ref))
    print('bf16-operand GEMM         vs fp32 GEMM on the same gfeat: max %.2e  mean %.2e   (the GEMM alone)' % e(bf_tail, exact_tail))
    top2 = ref.topk(2, dim=1).values
    print('oracle top-1 / top-2 logit gap at step 0: median %.3f nats, min %.3f' % (float((top2[:, 0] - top2[:, 1]).median()), float((top2[:, 0] - top2[:, 1]).min())))


if __name__ == '__main__':
    main()

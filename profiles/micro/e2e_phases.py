#!/usr/bin/env python
"""Host-side phase times of the streamed e2e loop (where does a batch's wall time go?)."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import synthetic as syn, generations as FG, engine as E

V, B, W, T = 30522, 64, 4, 40
sd = syn.make_case_decoder_state(123456, V, 256)
host = syn.make_case_inputs(20211, B, 60, 10, 256, V, 256).pin()
keys = ('mem_q', 'mem_p', 'query', 'passage', 'prior_q', 'prior_p', 'answer_rep', 'source_map')
hd = {k: getattr(host, k) for k in keys}
model = FG.FastCaSE(sd, device='cuda:0', dtype='bf16', max_dec_len=T, beam_width=W)
list(FG.beam_batches(model, (hd for _ in range(3)), None, T, W))
torch.cuda.synchronize()

orig_prefill = E.CaseDecodeEngine.prefill
orig_launch = E.CaseDecodeEngine.launch
orig_finish = E.CaseDecodeEngine._finish_tokens
log = []
def wrap(name, fn):
    def f(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); log.append((name, (time.perf_counter() - t0) * 1e3)); return r
    return f
E.CaseDecodeEngine.prefill = wrap('prefill', orig_prefill)
E.CaseDecodeEngine.launch = wrap('launch', orig_launch)
E.CaseDecodeEngine._finish_tokens = wrap('finish', orig_finish)
t0 = time.perf_counter()
n = 0
for out in FG.beam_batches(model, (hd for _ in range(6)), None, T, W):
    n += 1
    log.append(('yield', (time.perf_counter() - t0) * 1e3))
torch.cuda.synchronize()
print('total ms', (time.perf_counter() - t0) * 1e3, 'per batch', (time.perf_counter() - t0) * 1e3 / n)
for name, ms in log:
    print(f'{name:8s} {ms:8.2f}')

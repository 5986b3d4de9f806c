import os, sys, torch
sys.path.insert(0, '/root/repo')
from case_rg_b200 import synthetic as syn, generations as FG, _lib as L
V, B, W, T = 30522, 64, 4, 40
sd = syn.make_case_decoder_state(123456, V, 256)
host = syn.make_case_inputs(20211, B, 60, 10, 256, V, 256).pin()
keys = ('mem_q', 'mem_p', 'query', 'passage', 'prior_q', 'prior_p', 'answer_rep', 'source_map')
hd = {k: getattr(host, k) for k in keys}
ug = os.environ.get('UG', '1') == '1'
model = FG.FastCaSE(sd, device='cuda:0', dtype='bf16', max_dec_len=T, beam_width=W, use_graph=ug)
d = host.to('cuda:0')
dd = {k: getattr(d, k) for k in keys}
if os.environ.get('PRE', '1') == '1':
    print('resident', FG.beam(model, dd, None, T, W).shape)
    torch.cuda.synchronize()
for i, out in enumerate(FG.beam_batches(model, (hd for _ in range(3)), None, T, W)):
    print('batch', i, out.shape)
torch.cuda.synchronize()
print('ok')

import sys, torch, time
sys.path.insert(0, '.')
from case_rg_b200 import synthetic as syn, generations as FG
V, B, T, W = 50000, 128, 40, 4
sd = syn.make_gttp_state(1, V, 256, 256)
model = FG.FastGTTP(sd, device='cuda', dtype='bf16', max_dec_len=T, beam_width=W, vocab_impl=1)
keys = ('context', 'background', 'background_map', 'src_output', 'bg_output', 'init_state')
inp = syn.make_gttp_inputs(2, B, 60, 10, 100, V, 256)
host = {k: getattr(inp, k).pin_memory() for k in keys}
dev = {k: v.cuda() for k, v in host.items()}
def run(src, n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    outs = list(FG.beam_batches(model, (src for _ in range(n)), None, T, W))
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for _ in range(2): run(host, 3)
print('pinned host  ms/batch', run(host, 10))
print('device dicts ms/batch', run(dev, 10))
def seq(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        d = {k: v.to('cuda', non_blocking=True) for k, v in host.items()}
        FG.beam(model, d, None, T, W).cpu()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('sequential   ms/batch', seq(10))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(host, 3)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in ev)
mem = [(e.time_range.start - t0, e.time_range.end - t0, e.name) for e in ev if 'emcpy' in e.name and (e.time_range.end - e.time_range.start) > 500]
for m in mem[:12]: print('memcpy %9.1f -> %9.1f us  %s' % m)
ks = [(e.time_range.start - t0, e.time_range.end - t0) for e in ev if 'emcpy' not in e.name and 'emset' not in e.name]
print('kernels span', min(k[0] for k in ks), max(k[1] for k in ks), 'n', len(ks))
# gaps > 200 us between consecutive kernels
ks.sort(); 
for a, b in zip(ks, ks[1:]):
    if b[0] - a[1] > 300: print('gap %.1f us at %.1f' % (b[0] - a[1], a[1]))

"""Kernel-level GPU checks through the C ABI (the end-to-end parity tests live in test_gpu_parity.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.timeout(120)
@pytest.mark.parametrize('R,V,bias', [(8, 1000, False), (64, 30522, False), (256, 30522, False), (300, 5000, True),
                                      (512, 50000, True), (17, 130, True)])
def test_vocab_gemm_tcgen05_vs_torch(R, V, bias):
    """tcgen05 output projection (bf16 x bf16 -> fp32 in TMEM) against torch on the same bf16-rounded operands."""
    from case_rg_b200 import _lib as L
    from case_rg_b200.engine import pack_vocab_tc
    g = torch.Generator().manual_seed(R * 7 + V)
    f = torch.randn(R, 256, generator=g).to(DEV)
    W = (torch.randn(V, 256, generator=g) * 0.1).to(DEV)
    b = torch.randn(V, generator=g).to(DEV) if bias else None
    ldl = -(-V // 8) * 8
    out = torch.full((R, ldl), float('nan'), device=DEV)
    Wp = pack_vocab_tc(W)
    lib = L.load()
    assert Wp.numel() * 2 == lib.case_vocab_tc_packed_weight_bytes(V)
    ws = torch.zeros(lib.case_vocab_tc_workspace_bytes(R), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    L.call('case_vocab_gemm', f.data_ptr(), Wp.data_ptr(), L.ptr(b), out.data_ptr(), R, V, ldl, L.BF16, 1,
           ws.data_ptr(), st)
    torch.cuda.synchronize()
    want = f.bfloat16().float() @ W.bfloat16().float().t()
    if bias:
        want = want + b
    got = out[:, :V]
    assert torch.isfinite(got).all()
    assert rel_err(got, want) < 2e-5, rel_err(got, want)
    # and the SIMT kernel on the same inputs (fp32 activations, bf16 weights)
    out2 = torch.zeros(R, ldl, device=DEV)
    Wb = W.bfloat16().contiguous()
    L.call('case_vocab_gemm', f.data_ptr(), Wb.data_ptr(), L.ptr(b), out2.data_ptr(), R, V, ldl, L.BF16, 0, None, st)
    torch.cuda.synchronize()
    want2 = f @ W.bfloat16().float().t() + (b if bias else 0)
    assert rel_err(out2[:, :V], want2) < 2e-5


@pytest.mark.parametrize('K,N,nseg', [(256, 256, 1), (512, 256, 2), (768, 256, 3), (1280, 768, 3), (1536, 256, 4)])
@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_row_linear_stream_vs_torch(K, N, nseg, dtype):
    """Generic row linear (bulk-copy weight streaming) against torch, incl. concat / per-query segments."""
    import ctypes as C
    from case_rg_b200 import _lib as L
    from case_rg_b200.engine import pack_tiled, unpack_tiled
    R, W = 37, 2
    g = torch.Generator().manual_seed(K + N)
    widths = {1: [K], 2: [256, K - 256], 3: [256, (K - 256) // 2, (K - 256) // 2], 4: [256, 256, 512, 512]}[nseg]
    segs = []
    for i, wd in enumerate(widths):
        per_query = (i == len(widths) - 1 and nseg > 1)
        rows = -(-R // W) if per_query else R
        segs.append((torch.randn(rows, wd, generator=g).to(DEV), W if per_query else 1))
    Wt = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(R, N, generator=g).to(DEV)
    td, cd = (torch.float32, L.F32) if dtype == 'fp32' else (torch.bfloat16, L.BF16)
    Wp = pack_tiled(Wt, td)
    out = torch.zeros(R, N, device=DEV)
    a = L.RowLinArgs()
    for i, (t, div) in enumerate(segs):
        a.seg[i].p, a.seg[i].ld, a.seg[i].width, a.seg[i].div, a.seg[i].gather = t.data_ptr(), t.size(1), t.size(1), div, 0
    a.nseg, a.K, a.Wt, a.bias, a.N, a.act = nseg, K, Wp.data_ptr(), bias.data_ptr(), N, 1
    a.res, a.ldres, a.out, a.ldo, a.R, a.dtype = res.data_ptr(), N, out.data_ptr(), N, R, cd
    L.call('case_row_linear', C.byref(a), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    x = torch.cat([t if d == 1 else t.repeat_interleave(d, 0)[:R] for t, d in segs], 1)
    if dtype == 'bf16':
        x = x.bfloat16().float()          # the tensor-core kernel rounds its A operand to bf16
    want = torch.nn.functional.gelu(x @ unpack_tiled(Wp, N, K).t() + bias) + res
    assert rel_err(out, want) < 2e-5, rel_err(out, want)




@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
@pytest.mark.parametrize('W,S,DV,nsplit,use_prior', [(1, 60, 256, 1, True), (4, 2560, 256, 10, True), (4, 1000, 512, 3, False),
                                                      (8, 333, 256, 2, True), (2, 130, 512, 2, False)])
def test_additive_attention_vs_torch(W, S, DV, nsplit, use_prior, dtype):
    """Fused additive attention (scores, softmax partials, context partials) against torch, including
    all-padding tiles, a PAD-input row and the prior-weighted sum."""
    from case_rg_b200 import _lib as L
    B, H = 3, 256
    g = torch.Generator().manual_seed(W * 100 + S)
    td, cd = (torch.float32, L.F32) if dtype == 'fp32' else (torch.bfloat16, L.BF16)
    qa = torch.randn(B * W, H, generator=g).to(DEV)
    U = torch.randn(B, S, H, generator=g).to(DEV).to(td)
    Mv = torch.randn(B, S, DV, generator=g).to(DEV).to(td)
    v = (torch.randn(H, generator=g) * 0.3).to(DEV)
    mask = torch.rand(B, S, generator=g) > 0.25
    mask[:, 0] = True
    if S > 200:
        mask[1, 64:192] = False            # whole tiles of padding
    mask = mask.to(DEV)
    prior = torch.rand(B, S, generator=g).to(DEV) if use_prior else None
    tok = torch.ones(B * W, 4, dtype=torch.int32, device=DEV)
    tok[0, 2] = 0                          # row 0 consumes a PAD token at t = 2 -> fully masked row
    scores = torch.full((B * W, S), float('nan'), device=DEV)
    stats = torch.zeros(B * W, nsplit, 4, device=DEV)
    ctxp = torch.zeros(B * W, nsplit, DV, device=DEV)
    L.call('case_additive_attn', qa.data_ptr(), U.data_ptr(), Mv.data_ptr(), v.data_ptr(), mask.to(torch.uint8).data_ptr(),
           L.ptr(prior), tok.data_ptr(), 4, 2, B, W, S, DV, nsplit, scores.data_ptr(), stats.data_ptr(), ctxp.data_ptr(),
           0, cd, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    Uf, Mf = U.float(), Mv.float()
    e = (torch.tanh(qa.view(B, W, 1, H) + Uf.view(B, 1, S, H)) @ v).view(B * W, S)
    ok = mask.repeat_interleave(W, 0).clone()
    ok[0] = False
    e = e.masked_fill(~ok, float('-inf'))
    assert torch.equal(torch.isinf(scores), torch.isinf(e))
    fin = ~torch.isinf(e)
    tol = 1e-5 if dtype == 'fp32' else 1e-5
    assert float((scores[fin] - e[fin]).abs().max()) < 1e-4 * max(1.0, float(e[fin].abs().max()))
    # merge the partials the way the finalisers do and compare with a plain softmax attention
    m = stats[..., 0]
    M = m.max(1, keepdim=True).values
    w = torch.where(torch.isinf(m), torch.zeros_like(m), torch.exp(m - M))
    Z = (stats[..., 1] * w).sum(1)
    Q = (stats[..., 2] * w).sum(1)
    ctx = (ctxp * w.unsqueeze(-1)).sum(1) / Z.clamp_min(1e-30).unsqueeze(-1)
    a = torch.softmax(e, 1)
    a = torch.where(torch.isnan(a), torch.zeros_like(a), a)
    want_ctx = torch.bmm(a.view(B, W, S), Mf).view(B * W, DV)
    live = Z > 0
    assert bool((~live)[0]) and bool(live[1:].all())
    assert rel_err(ctx[live], want_ctx[live]) < 2e-4
    pr = prior.repeat_interleave(W, 0) if use_prior else torch.ones_like(a)
    assert rel_err((Q / Z.clamp_min(1e-30))[live], (pr * a).sum(1)[live]) < 2e-4


@pytest.mark.parametrize('compact', [False, True])
@pytest.mark.parametrize('W,S,nsplit,use_prior', [(1, 60, 1, True), (4, 2560, 9, True), (4, 1000, 3, False), (8, 333, 2, True),
                                                  (2, 130, 2, False), (3, 77, 1, True)])
def test_additive_attention_gate_form_vs_torch(W, S, nsplit, use_prior, compact):
    """case_additive_attn_gate: scores, softmax partials and the gate partials sum_s exp(e-m) G[s] against torch
    (masked walk and the compacted valid-key walk), incl. all-padding tiles, an empty query and a PAD-input row."""
    from case_rg_b200 import _lib as L
    B, H = 4, 256
    g = torch.Generator().manual_seed(W * 100 + S + 7)
    qa = torch.randn(B * W, H, generator=g).to(DEV)
    U = torch.randn(B, S, H, generator=g).to(DEV).to(torch.bfloat16)
    G = torch.randn(B, S, 4, generator=g).to(DEV)
    v = (torch.randn(H, generator=g) * 0.3).to(DEV)
    mask = torch.rand(B, S, generator=g) > 0.25
    mask[:, 0] = True
    if S > 200:
        mask[1, 64:192] = False            # whole tiles of padding
    mask[3] = False                        # a query without any valid key
    mask = mask.to(DEV)
    prior = torch.rand(B, S, generator=g).to(DEV) if use_prior else None
    tok = torch.ones(B * W, 4, dtype=torch.int32, device=DEV)
    tok[0, 2] = 0                          # row 0 consumes a PAD token at t = 2 -> fully masked row
    scores = torch.full((B * W, S), float('nan'), device=DEV)
    stats = torch.zeros(B * W, nsplit, 4, device=DEV)
    gpart = torch.zeros(B * W, nsplit, 4, device=DEV)
    cidx = ncount = qorder = nsq = None
    if compact:
        cidx = torch.argsort(~mask, dim=1, stable=True).to(torch.int32)
        ncount = mask.sum(1).to(torch.int32)
        qorder = torch.argsort(ncount, descending=True, stable=True).to(torch.int32)
        scores.masked_fill_(~mask.repeat_interleave(W, 0), float('-inf'))      # the caller's job in this form
        # work-proportional splits: queries use between 1 and nsplit of the slots
        nsq = (ncount.float() / max(1.0, float(ncount.max())) * nsplit).ceil().clamp(1, nsplit).to(torch.int32)
        stats.fill_(float('nan'))
        gpart.fill_(float('nan'))
    L.call('case_additive_attn_gate', qa.data_ptr(), U.data_ptr(), G.data_ptr(), v.data_ptr(),
           mask.to(torch.uint8).data_ptr(), L.ptr(prior), tok.data_ptr(), 4, 2, B, W, S, nsplit, scores.data_ptr(),
           stats.data_ptr(), gpart.data_ptr(), 0, L.ptr(cidx), L.ptr(ncount), L.ptr(qorder), L.ptr(nsq),
           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    e = (torch.tanh(qa.view(B, W, 1, H) + U.float().view(B, 1, S, H)) @ v).view(B * W, S)
    ok = mask.repeat_interleave(W, 0).clone()
    ok[0] = False
    e = e.masked_fill(~ok, float('-inf'))
    assert torch.equal(torch.isinf(scores), torch.isinf(e))
    fin = ~torch.isinf(e)
    assert float((scores[fin] - e[fin]).abs().max()) < 1e-4 * max(1.0, float(e[fin].abs().max()))
    m = stats[..., 0]
    M = m.max(1, keepdim=True).values
    w = torch.where(torch.isinf(m), torch.zeros_like(m), torch.exp(m - M))
    Z = (stats[..., 1] * w).sum(1)
    Q = (stats[..., 2] * w).sum(1)
    gs = (gpart * w.unsqueeze(-1)).sum(1) / Z.clamp_min(1e-30).unsqueeze(-1)
    a = torch.softmax(e, 1)
    a = torch.where(torch.isnan(a), torch.zeros_like(a), a)
    want = torch.bmm(a.view(B, W, S), G).view(B * W, 4)
    live = Z > 0
    assert bool((~live)[0]) and bool(live[1:3 * W].all()) and not bool(live[3 * W:].any())
    assert rel_err(gs[live][:, :3], want[live][:, :3]) < 2e-4
    pr = prior.repeat_interleave(W, 0) if use_prior else torch.ones_like(a)
    assert rel_err((Q / Z.clamp_min(1e-30))[live], (pr * a).sum(1)[live]) < 2e-4


# --------------------------------------------------------------------------- cluster layer kernels
def _chain_case(B, W, T, V=3000, seeds=(51, 52)):
    from case_rg_b200 import synthetic as syn
    sd = syn.make_case_decoder_state(seeds[0], V, 256, peaked=0.3, boost={syn.EOS: 6.0}, gen_gate_bias=2.0)
    inp = syn.make_case_inputs(seeds[1], B, 20, 3, 40, V, 256).to('cuda')
    data = dict(mem_q=inp.mem_q, mem_p=inp.mem_p, query=inp.query, passage=inp.passage, prior_q=inp.prior_q,
                prior_p=inp.prior_p, answer_rep=inp.answer_rep, source_map=inp.source_map)
    return sd, data


def _teacher_forced_states(model, data, prefix):
    """Drive an engine on a fixed token prefix (W = 1): per step (h, q2, logits, dist)."""
    B, n = prefix.shape
    S0 = data['mem_q'].reshape(B, -1, 256).size(1)
    S1 = data['mem_p'].reshape(B, -1, 256).size(1)
    eng = model.engine_for(B, 1, S0, S1, max(n, 2))
    eng.prefill(data['mem_q'], data['mem_p'], data['query'].ne(0), data['passage'].ne(0), data['prior_q'],
                data['prior_p'], data['answer_rep'], data['source_map'])
    eng.state.reset()
    outs = []
    for t in range(n):
        eng.state.tok[:, t] = prefix[:, t].to('cuda', torch.int32)
        dist = eng.step_distribution(t)
        torch.cuda.synchronize()
        outs.append(dict(h=eng.h.clone(), q2=eng.q2.clone(), logits=eng.logits[:, :eng.V].clone(), dist=dist.clone()))
    return outs


@pytest.mark.timeout(300)
@pytest.mark.parametrize('B,T', [(3, 6), (20, 12), (64, 5), (1, 48)])
def test_layer_chain_vs_fp32_and_row_block_kernels(B, T):
    """case_layer_chain (4-CTA clusters, DSMEM exchange) computes what case_layer_front/back compute from
    the same bf16 operands; only summation order differs.  On a teacher-forced prefix (with PAD tokens,
    so masked history keys occur) every step of both bf16 paths is compared with the fp32 engine: the
    cluster path must be as close to fp32 as the row-block path is (bf16 noise level)."""
    from case_rg_b200 import _lib as L
    from case_rg_b200.generations import FastCaSE
    sd, data = _chain_case(B, 1, T)
    g = torch.Generator().manual_seed(77)
    prefix = torch.randint(1000, 3000, (B, T), generator=g)
    prefix[:, 0] = 1
    if T > 3:
        prefix[::2, 2] = 0           # PAD inside the history of every other row
    lib = L.load()
    ref = _teacher_forced_states(FastCaSE(sd, device='cuda', dtype='fp32', use_graph=False), data, prefix)
    outs = {}
    for chain in (0, 1):          # per-engine option word: row-block kernels against the cluster kernels
        model = FastCaSE(sd, device='cuda', dtype='bf16', use_graph=False, opt=0 if chain else L.OPT_NO_CHAIN)
        outs[chain] = _teacher_forced_states(model, data, prefix)
    worst = {}
    for t in range(T):
        for k in ('h', 'q2', 'logits', 'dist'):
            assert torch.isfinite(outs[1][t][k]).all(), (t, k)
            e0, e1 = rel_err(outs[0][t][k], ref[t][k]), rel_err(outs[1][t][k], ref[t][k])
            worst[k] = max(worst.get(k, 0.0), e1)
            assert e1 < (8e-2 if k == 'dist' else 3e-2), (t, k, e0, e1)
            assert e1 < 3.0 * e0 + (2e-2 if k == 'dist' else 5e-3), (t, k, e0, e1)
    print(worst)


@pytest.mark.timeout(120)
@pytest.mark.parametrize('B,W', [(5, 4), (9, 8), (64, 4), (2, 3)])
def test_layer_chain_first_step_beam_rows(B, W):
    """Step 0 of a beam search (no token feedback yet): all W slots of a query, partial last cluster."""
    from case_rg_b200 import _lib as L
    from case_rg_b200.generations import FastCaSE
    sd, data = _chain_case(B, W, 1)
    lib = L.load()
    m32 = FastCaSE(sd, device='cuda', dtype='fp32', use_graph=False)
    m32.fast_search(data, 1, W, L.MODE_BEAM)
    ref = dict(h=m32.last_engine.h.clone(), logits=m32.last_engine.logits[:, :3000].clone())
    err = {}
    for chain in (0, 1):
        model = FastCaSE(sd, device='cuda', dtype='bf16', use_graph=False, opt=0 if chain else L.OPT_NO_CHAIN)
        model.fast_search(data, 1, W, L.MODE_BEAM)
        torch.cuda.synchronize()
        eng = model.last_engine
        err[chain] = dict(h=rel_err(eng.h, ref['h']), logits=rel_err(eng.logits[:, :3000], ref['logits']))
    for k in ('h', 'logits'):
        assert err[1][k] < 3e-2 and err[1][k] < 2.0 * err[0][k] + 3e-3, err


def _common_prefix(a, b):
    n = min(a.size(1), b.size(1))
    neq = (a[:, :n] != b[:, :n]).int()
    first = torch.where(neq.any(1), neq.argmax(1), torch.full((a.size(0),), n))
    return first.float().mean().item(), n


@pytest.mark.timeout(300)
@pytest.mark.parametrize('B,W,T', [(64, 4, 40), (7, 1, 48), (6, 2, 50)])
def test_layer_chain_full_search_agrees(B, W, T):
    """Whole searches (CUDA graph, T up to the shared-memory history limit of 48; T = 50 exercises the
    automatic fallback to the row-block kernels).  Free-running bf16 decodes leave the fp32 trajectory at
    the first near-tie, so the measure is the length of the common prefix with the fp32 answers: the
    cluster path must stay on the fp32 trajectory about as long as the row-block path does."""
    from case_rg_b200 import _lib as L
    from case_rg_b200.generations import FastCaSE
    sd, data = _chain_case(B, W, T, seeds=(53, 54))
    mode = L.MODE_BEAM if W > 1 else L.MODE_PROTO_GREEDY
    lib = L.load()
    ref = FastCaSE(sd, device='cuda', dtype='fp32', use_graph=True).fast_search(data, T, W, mode).cpu()
    pref = {}
    for chain in (0, 1):
        model = FastCaSE(sd, device='cuda', dtype='bf16', use_graph=True, opt=0 if chain else L.OPT_NO_CHAIN)
        pref[chain], n = _common_prefix(model.fast_search(data, T, W, mode).cpu(), ref)
    print(pref, n)
    assert pref[1] >= 0.7 * pref[0] - 1.0, (pref, n)
    if T > 48:
        assert pref[1] == pref[0]       # same kernels ran both times


# --------------------------------------------------------------------------- fused tail
@pytest.mark.timeout(120)
@pytest.mark.parametrize('R,W,V,S0,S1,K', [(8, 1, 1000, 12, 40, 1), (12, 4, 30522, 60, 2560, 4), (6, 2, 50000, 7, 333, 2),
                                            (8, 8, 5003, 16, 100, 8), (3, 3, 2000, 5, 9, 3)])
def test_row_tail_matches_unfused_kernels(R, W, V, S0, S1, K):
    """case_row_tail == case_finalize_rows + case_softmax_mix + 2 x case_copy_scatter + case_topk_rows:
    same gates / fac / ctx, the same distribution up to float summation order, identical top-k indices
    (repeated source ids and exact value ties included), bit-exact copy targets."""
    import ctypes as C
    from case_rg_b200 import _lib as L
    torch.manual_seed(R * 1000 + V)
    dev, f32 = 'cuda', dict(dtype=torch.float32, device='cuda')
    B, H, MS = R // W, 256, L.MAX_SPLIT
    ldv = -(-V // 8) * 8
    ns = (1, 3)
    S = (S0, S1)
    logits = torch.randn(R, ldv, **f32) * 3
    logits[:, 7] = logits[:, 3]                              # exact ties in the base distribution
    hN, h = torch.randn(R, H, **f32), torch.randn(R, H, **f32)
    lg, lb = torch.ones(H, **f32), torch.zeros(H, **f32)
    Wm, bm = torch.randn(3, 3 * H, **f32) * 0.05, torch.randn(3, **f32)
    stats = [torch.rand(R, n, 4, **f32) + 0.1 for n in ns]
    ctxp = [torch.randn(R, n, H, **f32) for n in ns]
    attn = [torch.randn(R, s, **f32) for s in S]
    attn[1][:, ::5] = float('-inf')                          # masked source positions
    prior = [torch.rand(B, s, **f32) for s in S]
    smap = torch.randint(0, V, (B, S0 + S1), device=dev, dtype=torch.int32)
    smap[:, S0:S0 + 4] = smap[:, :1]                         # repeated ids across and inside the memories
    st = torch.cuda.current_stream().cuda_stream

    # unfused reference path (finalize needs h and recomputes hN = LN(h); feed it h with identity LN params)
    hN_ref = torch.empty_like(hN)
    ctx_r = [torch.zeros(R, H, **f32) for _ in range(2)]
    gates_r, fac_r = torch.zeros(R, 4, **f32), torch.zeros(R, 2, MS, **f32)
    L.call('case_finalize_rows', h.data_ptr(), lg.data_ptr(), lb.data_ptr(), stats[0].data_ptr(), ctxp[0].data_ptr(), ns[0],
           stats[1].data_ptr(), ctxp[1].data_ptr(), ns[1], Wm.data_ptr(), bm.data_ptr(), hN_ref.data_ptr(),
           ctx_r[0].data_ptr(), ctx_r[1].data_ptr(), gates_r.data_ptr(), fac_r.data_ptr(), R, st)
    dist_r = torch.zeros(R, ldv, **f32)
    L.call('case_softmax_mix', logits.data_ptr(), ldv, gates_r.data_ptr(), dist_r.data_ptr(), ldv, R, V, 0, st)
    for i, off in enumerate((0, S0)):
        L.call('case_copy_scatter', smap.data_ptr(), S0 + S1, off, prior[i].data_ptr(), attn[i].data_ptr(),
               fac_r[:, i].data_ptr(), 2 * MS, dist_r.data_ptr(), ldv, B, W, S[i], V, st)
    tv_r, ti_r = torch.zeros(R, K, **f32), torch.zeros(R, K, dtype=torch.int32, device=dev)
    L.call('case_topk_rows', dist_r.data_ptr(), ldv, R, V, K, tv_r.data_ptr(), ti_r.data_ptr(), st)

    # fused
    a = L.TailArgs()
    a.R, a.V, a.W, a.K, a.ldl, a.ldd, a.mask_col0, a.nmem, a.do_finalize = R, V, W, K, ldv, ldv, 0, 2, 1
    a.fac_ld, a.map_ld = 2 * MS, S0 + S1
    ctx_f = [torch.zeros(R, H, **f32) for _ in range(2)]
    gates_f, fac_f = torch.zeros(R, 4, **f32), torch.zeros(R, 2, MS, **f32)
    tv_f, ti_f = torch.zeros(R, K, **f32), torch.zeros(R, K, dtype=torch.int32, device=dev)
    dist_f = torch.zeros(R, ldv, **f32)
    for i in range(2):
        a.ns[i], a.fac_off[i], a.map_off[i], a.S[i] = ns[i], i * MS, (0, S0)[i], S[i]
        a.stats[i], a.ctxp[i], a.ctx[i] = stats[i].data_ptr(), ctxp[i].data_ptr(), ctx_f[i].data_ptr()
        a.prior[i], a.attn_un[i] = prior[i].data_ptr(), attn[i].data_ptr()
    a.logits, a.hN, a.Wm, a.bm = logits.data_ptr(), hN_ref.data_ptr(), Wm.data_ptr(), bm.data_ptr()
    a.gates, a.fac, a.map = gates_f.data_ptr(), fac_f.data_ptr(), smap.data_ptr()
    a.top_vals, a.top_idx, a.dist = tv_f.data_ptr(), ti_f.data_ptr(), dist_f.data_ptr()
    L.check(L.load().case_row_tail(C.byref(a), st), 'case_row_tail')
    torch.cuda.synchronize()

    assert rel_err(gates_f, gates_r) < 1e-5 and rel_err(fac_f[:, :, :2], fac_r[:, :, :2]) < 1e-5
    assert rel_err(ctx_f[0], ctx_r[0]) < 1e-5 and rel_err(ctx_f[1], ctx_r[1]) < 1e-5
    assert rel_err(dist_f[:, :V], dist_r[:, :V]) < 1e-5
    assert torch.allclose(dist_f[:, :V].sum(1), dist_r[:, :V].sum(1), rtol=1e-5)
    # scatter targets are exact: entries that no source id points at are untouched multiples of the base
    touched = torch.zeros(B, V, dtype=torch.bool, device=dev)
    touched.scatter_(1, smap.long(), True)
    base = torch.zeros(R, ldv, **f32)
    L.call('case_softmax_mix', logits.data_ptr(), ldv, gates_r.data_ptr(), base.data_ptr(), ldv, R, V, 0, st)
    torch.cuda.synchronize()
    untouched = ~touched.repeat_interleave(W, 0)
    assert rel_err(dist_f[:, :V][untouched], base[:, :V][untouched]) < 1e-6
    # top-k of the fused kernel is exactly the top-k (value desc, index asc) of ITS distribution
    d = dist_f[:, :V].double().cpu()
    key = torch.argsort(torch.argsort(-d, dim=1, stable=True), dim=1)   # rank with ties -> lower index first
    want = torch.argsort(key, dim=1)[:, :K]
    assert torch.equal(ti_f.cpu().long(), want), (ti_f, want)
    assert torch.equal(tv_f.cpu(), torch.gather(dist_f[:, :V].cpu(), 1, want))
    agree = (ti_f == ti_r).float().mean().item()
    assert agree > 0.95, agree


@pytest.mark.timeout(120)
@pytest.mark.parametrize('R,W,V,S0,S1,K,finalize', [(8, 1, 1000, 12, 40, 1, 1), (12, 4, 30522, 60, 2560, 4, 1),
                                                     (6, 2, 50000, 7, 333, 2, 1), (8, 8, 5003, 16, 100, 8, 1),
                                                     (3, 3, 2000, 5, 9, 3, 1), (4, 4, 30522, 60, 10240, 4, 1),
                                                     (6, 2, 4000, 1, 500, 2, 0)])
def test_sparse_tail_matches_dense_tail(R, W, V, S0, S1, K, finalize):
    """case_vocab_base + case_sparse_tail give the top-k of the dense distribution (case_row_tail with dist
    written out): identical indices, values equal up to the summation order of repeated copy targets.
    Includes exact base ties, masked positions, ids repeated across both memories, source ids that hit
    the base top entries, and (finalize = 0) the GTTP form with column 0 masked."""
    import ctypes as C
    from case_rg_b200 import _lib as L
    torch.manual_seed(R * 7919 + V + S1)
    dev, f32 = 'cuda', dict(dtype=torch.float32, device='cuda')
    B, H, MS = R // W, 256, L.MAX_SPLIT
    ldv = -(-V // 8) * 8
    ns = (1, 3)
    S = (S0, S1)
    nmem = 2 if finalize else 1
    logits = torch.randn(R, ldv, **f32) * 3
    logits[:, 7] = logits[:, 3]
    logits[:, 11] = logits[:, 3]                             # exact ties in the base distribution
    top_ids = logits[:, :V].topk(4, dim=1).indices           # some copy targets ARE base top entries
    hN = torch.randn(R, H, **f32)
    Wm, bm = torch.randn(3, 3 * H, **f32) * 0.05, torch.randn(3, **f32)
    stats = [torch.rand(R, n, 4, **f32) + 0.1 for n in ns]
    ctxp = [torch.randn(R, n, H, **f32) for n in ns]
    attn = [torch.randn(R, s, **f32) for s in S]
    attn[1][:, ::5] = float('-inf')
    prior = [torch.rand(B, s, **f32) for s in S]
    smap = torch.randint(0, V, (B, S0 + S1), device=dev, dtype=torch.int32)
    smap[:, S0:S0 + 3] = smap[:, :1]                         # repeated ids across and inside the memories
    smap[:, S0 + 3] = top_ids[::W, 0].int()
    smap[:, S0 + 4] = 0
    st = torch.cuda.current_stream().cuda_stream
    gates, fac = torch.rand(R, 4, **f32), torch.rand(R, 2, MS, **f32)
    fac[:, :, 1] = 0.5

    def args():
        a = L.TailArgs()
        a.R, a.V, a.W, a.K, a.ldl, a.ldd, a.mask_col0, a.nmem, a.do_finalize = R, V, W, K, ldv, ldv, 1 - finalize, nmem, finalize
        a.fac_ld, a.map_ld = 2 * MS, S0 + S1
        keep = dict(ctx=[torch.zeros(R, H, **f32) for _ in range(2)], gates=gates.clone(), fac=fac.clone(),
                    tv=torch.zeros(R, K, **f32), ti=torch.zeros(R, K, dtype=torch.int32, device=dev),
                    dist=torch.zeros(R, ldv, **f32))
        for i in range(nmem):
            j = i if finalize else 1                          # GTTP form: one memory (the big one)
            a.ns[i], a.fac_off[i], a.map_off[i], a.S[i] = ns[j], i * MS, (0, S0)[j], S[j]
            a.stats[i], a.ctxp[i], a.ctx[i] = stats[j].data_ptr(), ctxp[j].data_ptr(), keep['ctx'][i].data_ptr()
            a.prior[i], a.attn_un[i] = (prior[j].data_ptr() if finalize else None), attn[j].data_ptr()
        a.logits, a.hN, a.Wm, a.bm = logits.data_ptr(), hN.data_ptr(), Wm.data_ptr(), bm.data_ptr()
        a.gates, a.fac, a.map = keep['gates'].data_ptr(), keep['fac'].data_ptr(), smap.data_ptr()
        a.top_vals, a.top_idx = keep['tv'].data_ptr(), keep['ti'].data_ptr()
        return a, keep

    lib = L.load()
    a, dense = args()
    a.dist = dense['dist'].data_ptr()
    L.check(lib.case_row_tail(C.byref(a), st), 'case_row_tail')
    k2 = 2 * K
    base_ms, base_e = torch.zeros(R, 4, 2, **f32), torch.zeros(R, 4, k2, **f32)
    base_i = torch.zeros(R, 4, k2, dtype=torch.int32, device=dev)
    L.call('case_vocab_base', logits.data_ptr(), ldv, R, V, 1 - finalize, k2, base_ms.data_ptr(), base_e.data_ptr(),
           base_i.data_ptr(), st)
    a2, sparse = args()
    L.check(lib.case_sparse_tail(C.byref(a2), base_ms.data_ptr(), base_e.data_ptr(), base_i.data_ptr(), k2, None, None, st),
            'case_sparse_tail')
    torch.cuda.synchronize()
    # base statistics against torch
    lg = logits[:, :V].clone()
    if not finalize:
        lg[:, 0] = float('-inf')
    mrow = base_ms[:, :, 0].max(1).values
    srow = (base_ms[:, :, 1] * (base_ms[:, :, 0] - mrow[:, None]).exp()).sum(1)
    assert torch.allclose(mrow, lg.max(1).values)
    assert torch.allclose(srow, (lg - lg.max(1, keepdim=True).values).exp().sum(1), rtol=1e-5)
    assert torch.allclose(sparse['gates'], dense['gates'], rtol=1e-5) and torch.allclose(sparse['fac'], dense['fac'], rtol=1e-5)
    # the dense tile is the reference: exact top-k (value desc, index asc) of it
    d = dense['dist'][:, :V].double().cpu()
    want = torch.argsort(torch.argsort(torch.argsort(-d, dim=1, stable=True), dim=1), dim=1)[:, :K]
    assert torch.equal(dense['ti'].cpu().long(), want)
    got_i, got_v = sparse['ti'].cpu().long(), sparse['tv'].cpu()
    ref_v = torch.gather(dense['dist'][:, :V].cpu(), 1, got_i)
    assert torch.allclose(got_v, ref_v, rtol=1e-5, atol=0), (got_v, ref_v)
    same = (got_i == want)
    if not bool(same.all()):   # only allowed where summation order flipped two values that agree to 2 ulp
        wv = torch.gather(dense['dist'][:, :V].cpu(), 1, want)
        assert torch.allclose(wv[~same], ref_v[~same], rtol=1e-5, atol=0), (got_i, want)
    if finalize:
        # plan mode: the prefill's sorted unique-id list instead of the hash table - same gates, same top-k
        from case_rg_b200.engine import build_copy_plan
        St = S0 + S1
        valid = torch.ones(B, St, dtype=torch.bool, device=dev)
        valid[:, S0::5] = False                              # exactly the positions whose scores are -inf above
        valid[:, S0 + 7] = False                             # and one valid-score position the plan leaves out ...
        cp = dict(n=torch.zeros(B, dtype=torch.int32, device=dev),
                  **{k: torch.zeros(B, St + 1, dtype=torch.int32, device=dev) for k in ('uid', 'first', 'start', 'perm')})
        attn_p = [attn[0], attn[1].clone()]
        attn_p[1][:, 7] = float('-inf')                      # ... whose score the dense reference then must not see either
        build_copy_plan(smap, valid, V, cp['n'], cp['uid'], cp['first'], cp['start'], cp['perm'])
        ad, dense_p = args()
        ad.attn_un[1] = attn_p[1].data_ptr()
        ad.dist = dense_p['dist'].data_ptr()
        L.check(lib.case_row_tail(C.byref(ad), st), 'case_row_tail')
        ap, planned = args()
        ap.attn_un[1] = attn_p[1].data_ptr()
        ap.cp_n, ap.cp_uid, ap.cp_first = cp['n'].data_ptr(), cp['uid'].data_ptr(), cp['first'].data_ptr()
        ap.cp_start, ap.cp_perm, ap.cp_ld = cp['start'].data_ptr(), cp['perm'].data_ptr(), St + 1
        L.check(lib.case_sparse_tail(C.byref(ap), base_ms.data_ptr(), base_e.data_ptr(), base_i.data_ptr(), k2, None, None, st),
                'case_sparse_tail(plan)')
        torch.cuda.synchronize()
        dp = dense_p['dist'][:, :V].double().cpu()
        want_p = torch.argsort(torch.argsort(torch.argsort(-dp, dim=1, stable=True), dim=1), dim=1)[:, :K]
        pi = planned['ti'].cpu().long()
        rvp = torch.gather(dense_p['dist'][:, :V].cpu(), 1, pi)
        assert torch.allclose(planned['tv'].cpu(), rvp, rtol=1e-5, atol=0), (planned['tv'], rvp)
        smp = pi == want_p
        if not bool(smp.all()):
            wv = torch.gather(dense_p['dist'][:, :V].cpu(), 1, want_p)
            assert torch.allclose(wv[~smp], rvp[~smp], rtol=1e-5, atol=0), (pi, want_p)
        # a second launch gives bit-identical values: no atomics, fixed summation order
        ap2, planned2 = args()
        ap2.attn_un[1] = attn_p[1].data_ptr()
        ap2.cp_n, ap2.cp_uid, ap2.cp_first = cp['n'].data_ptr(), cp['uid'].data_ptr(), cp['first'].data_ptr()
        ap2.cp_start, ap2.cp_perm, ap2.cp_ld = cp['start'].data_ptr(), cp['perm'].data_ptr(), St + 1
        L.check(lib.case_sparse_tail(C.byref(ap2), base_ms.data_ptr(), base_e.data_ptr(), base_i.data_ptr(), k2, None, None, st),
                'case_sparse_tail(plan, again)')
        torch.cuda.synchronize()
        assert torch.equal(planned2['tv'], planned['tv']) and torch.equal(planned2['ti'], planned['ti'])
        # gate form: ctxp replaced by the gate partials sum exp(e-m) (W_m,i . mem) - the same gates, factors and top-k
        a3, gform = args()
        gp = [torch.einsum('rnh,kh->rnk', ctxp[i], Wm[:, H * (1 + i):H * (2 + i)]) for i in range(2)]
        gp = [torch.cat([x, torch.zeros(R, x.size(1), 1, **f32)], 2).contiguous() for x in gp]
        for i in range(2):
            a3.ctxp[i], a3.ctx[i] = gp[i].data_ptr(), None
        a3.gate_ctx = 1
        L.check(lib.case_sparse_tail(C.byref(a3), base_ms.data_ptr(), base_e.data_ptr(), base_i.data_ptr(), k2, None, None, st),
                'case_sparse_tail(gate)')
        torch.cuda.synchronize()
        assert torch.allclose(gform['gates'], dense['gates'], rtol=1e-4, atol=1e-6)
        assert torch.allclose(gform['fac'], dense['fac'], rtol=1e-4, atol=1e-7)
        gi = gform['ti'].cpu().long()
        rv = torch.gather(dense['dist'][:, :V].cpu(), 1, gi)
        assert torch.allclose(gform['tv'].cpu(), rv, rtol=2e-4, atol=0)
        sm = gi == want
        if not bool(sm.all()):
            wv = torch.gather(dense['dist'][:, :V].cpu(), 1, want)
            assert torch.allclose(wv[~sm], rv[~sm], rtol=2e-4, atol=0), (gi, want)


@pytest.mark.timeout(120)
@pytest.mark.parametrize('B,S,compact,with_u,nl', [(3, 60, False, True, 4), (5, 700, True, True, 4), (4, 2560, True, True, 4),
                                                   (2, 130, True, False, 2), (2, 257, False, True, 1)])
def test_prefill_project_tcgen05_vs_gemm_and_pack(B, S, compact, with_u, nl):
    """case_prefill_project_tc (one tcgen05 GEMM, K|V tiles and U written from the epilogue, padding keys dropped on the
    load side) against the projected-rows GEMM + case_pack_kv_tiles[_gather] + the Uk.mem GEMM it replaces."""
    import ctypes as C
    from case_rg_b200 import _lib as L
    from case_rg_b200.engine import pack_vocab_tc
    g = torch.Generator().manual_seed(B * 1000 + S)
    H, NH = 256, 8
    mem = torch.randn(B, S, H, generator=g).to(DEV).bfloat16()
    Wkv = (torch.randn(nl * 2 * H, H, generator=g) / 16).to(DEV)
    Wu = (torch.randn(H, H, generator=g) / 16).to(DEV)
    bias = torch.randn(nl * 2 * H, generator=g).to(DEV)
    valid = (torch.rand(B, S, generator=g) < 0.7).to(DEV)
    valid[0] = False                                        # a query without any valid key
    if B > 1:
        valid[1] = True
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    ntile = -(-S // 64)
    mk = lambda: [torch.full((B, NH, ntile, 2, 64, 32), 7.0, dtype=torch.bfloat16, device=DEV) for _ in range(nl)]
    ref, got = mk(), mk()
    kv = (mem.float().reshape(B * S, H) @ Wkv.bfloat16().float().t() + bias).bfloat16().contiguous()
    cidx = ncount = None
    if compact:
        xidx = torch.argsort(~valid, dim=1, stable=True).to(torch.int32).contiguous()
        xcount = valid.sum(1).to(torch.int32).contiguous()
        cidx, ncount = xidx.data_ptr(), xcount.data_ptr()
        outs = (C.c_void_p * 4)(*[t.data_ptr() for t in ref], *[None] * (4 - nl))
        L.call('case_pack_kv_tiles_gather', kv.data_ptr(), kv.size(1), B, S, cidx, ncount, nl, outs, st)
    else:
        outs = (C.c_void_p * 4)(*[t.data_ptr() for t in ref], *[None] * (4 - nl))
        L.call('case_pack_kv_tiles', kv.data_ptr(), L.BF16, kv.size(1), B, S, nl, outs, st)
    U = torch.full((B, S, H), float('nan'), dtype=torch.bfloat16, device=DEV)
    Wp = pack_vocab_tc(torch.cat([Wkv, Wu], 0) if with_u else Wkv)
    outs2 = (C.c_void_p * 4)(*[t.data_ptr() for t in got], *[None] * (4 - nl))
    L.call('case_prefill_project_tc', mem.data_ptr(), Wp.data_ptr(), bias.data_ptr(), B, S, cidx, ncount, nl, outs2,
           U.data_ptr() if with_u else None, st)
    torch.cuda.synchronize()
    for l in range(nl):
        a, b = got[l].float(), ref[l].float()
        assert torch.isfinite(a).all()
        assert torch.equal(a == 0, b == 0) or float(((a == 0) != (b == 0)).float().mean()) < 1e-5   # the same zero rows
        assert torch.allclose(a, b, rtol=1e-2, atol=1e-2), (l, float((a - b).abs().max()))
        assert float((a != b).float().mean()) < 0.02, float((a != b).float().mean())       # 1-ulp roundings only
    if with_u:
        want = (mem.float().reshape(B * S, H) @ Wu.bfloat16().float().t()).view(B, S, H)
        assert torch.isfinite(U.float()).all()                                            # every position written
        assert torch.allclose(U.float(), want, rtol=1e-2, atol=1e-2), float((U.float() - want).abs().max())

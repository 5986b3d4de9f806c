"""SURVEY.md §8f N1 on the device: the pre-decode producers (shared encoder, Interaction, TransformerBlock stacks, prior /
answer representation) - every kernel against plain torch on the same inputs, the whole pipeline against the oracle
(oracle/producers.py, pinned on the unmodified reference) and against the golden made by the unmodified CaSE.forward, and
the answers decoded FROM TOKEN IDS (producers + decoder) against the reference's.

bf16 GEMM operands with fp32 accumulation / statistics / residual streams: north_star's bf16 band (2e-2 relative) is the
tolerance of everything that went through a GEMM; the kernels themselves are checked tighter on exact operands."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from case_rg_b200 import synthetic as syn
from helpers import GOLDEN, H, producers_case

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize('C,dt', [(256, torch.float32), (1280, torch.bfloat16), (256, torch.bfloat16), (1280, torch.float32)])
def test_ln_rows_wide_vs_torch(C, dt):
    from case_rg_b200 import _lib as L
    M = 1037
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(M, C, generator=g) * 2 + 0.5).to(DEV).to(dt)
    add = torch.randn(M, C, generator=g).to(DEV).to(dt)
    w, b = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV), (0.1 * torch.randn(C, generator=g)).to(DEV)
    for use_add in (False, True):
        y16 = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
        y32 = torch.empty(M, C, device=DEV)
        L.call('case_ln_rows_wide', x.data_ptr(), add.data_ptr() if use_add else None, L.BF16 if dt == torch.bfloat16 else L.F32,
               w.data_ptr(), b.data_ptr(), y16.data_ptr(), y32.data_ptr(), M, C, _st())
        want = F.layer_norm(x.float() + (add.float() if use_add else 0), (C,), w, b, 1e-5)
        assert rel(y32, want) < 2e-6
        assert rel(y16, want) < 5e-3


@pytest.mark.parametrize('C,L_,nseq', [(256, 100, 7), (1280, 256, 5), (256, 60, 3), (1280, 77, 4), (256, 512, 2)])
def test_enc_attention_vs_torch(C, L_, nseq):
    """Self-attention with the key padding mask against torch on the same bf16 q / k / v (fp32 math): ragged valid
    lengths, partial last key tile, sequences shorter than one tile."""
    from case_rg_b200 import _lib as L
    nh, hd = 8, C // 8
    g = torch.Generator().manual_seed(C + L_)
    qkv = (torch.randn(nseq * L_, 3 * C, generator=g) * 0.7).to(DEV).bfloat16()
    lens = torch.randint(2, L_ + 1, (nseq,), generator=g)
    lens[0] = L_
    mask = (torch.arange(L_)[None, :] < lens[:, None])
    km = mask.reshape(-1).to(torch.uint8).to(DEV)
    out = torch.full((nseq * L_, C), float('nan'), dtype=torch.bfloat16, device=DEV)
    L.call('case_enc_attention', qkv.data_ptr(), km.data_ptr(), nseq, L_, C, nh, out.data_ptr(), _st())
    torch.cuda.synchronize()
    q, k, v = [t.float().view(nseq, L_, nh, hd).transpose(1, 2) for t in qkv.float().split(C, dim=1)]
    s = (q @ k.transpose(-1, -2)) / (hd ** 0.5)
    s = s.masked_fill(~mask.to(DEV)[:, None, None, :], float('-inf'))
    want = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(nseq * L_, C)
    assert torch.isfinite(out.float()).all()
    # p is rounded to bf16 for the P.V MMA and the output to bf16
    assert rel(out, want) < 1.2e-2, rel(out, want)


@pytest.mark.parametrize('M,N,K,act,res,mask,out32', [(300, 256, 256, 0, None, False, False), (1000, 768, 256, 0, None, False, False),
                                                      (515, 256, 256, 1, 'f32', False, True), (260, 3840, 1280, 0, None, False, False),
                                                      (129, 1280, 1280, 0, 'bf16', False, False), (700, 256, 1280, 2, None, True, True),
                                                      (128, 256, 256, 2, 'f32', True, True), (20000, 256, 256, 0, 'bf16', True, False),
                                                      (200, 256, 64, 0, None, False, True), (333, 512, 192, 1, 'f32', False, False),
                                                      (1, 256, 256, 0, 'f32', True, True), (127, 1280, 1280, 2, None, False, False),
                                                      (40000, 512, 1024, 0, 'bf16', True, False), (257, 768, 1280, 1, 'f32', False, True)])
def test_gemm_rows_tc_vs_torch(M, N, K, act, res, mask, out32):
    """case_gemm_rows_tc (tcgen05 / TMEM, bias + activation + residual + row mask epilogue) against torch fp32 on the same
    bf16-rounded operands: partial last row tile, 1 .. 20 K stages (shorter than,
    equal to and longer than the four-stage ring), 1 .. 15 column chunks, more row tiles than CTAs (persistent loop), both unit shapes (one and
    two row tiles per unit, the latter with an odd tile count and more units than CTAs)."""
    from case_rg_b200 import _lib as L
    from case_rg_b200.producers import _Linear
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5)
    b = torch.randn(N, generator=g) * 0.1
    lin = _Linear(w, b, torch.device(DEV))
    r = None
    if res:
        r = torch.randn(M, N, generator=g).to(DEV)
        r = r.bfloat16() if res == 'bf16' else r
    rm = (torch.rand(M, generator=g) > 0.3).to(torch.uint8).to(DEV) if mask else None
    y = torch.full((M, N), float('nan'), dtype=torch.float32 if out32 else torch.bfloat16, device=DEV)
    L.call('case_gemm_rows_tc', x.data_ptr(), lin.wp.data_ptr(), lin.b.data_ptr(), M, N, K, act, L.ptr(r),
           (L.BF16 if res == 'bf16' else L.F32) if res else 0, L.ptr(rm), y.data_ptr(), L.F32 if out32 else L.BF16, _st())
    torch.cuda.synchronize()
    want = x.float() @ lin.w16.float().t() + lin.b
    want = F.gelu(want) if act == 1 else (torch.relu(want) if act == 2 else want)
    if r is not None:
        want = want + r.float()
    if rm is not None:
        want = want * rm.view(-1, 1).float()
    assert torch.isfinite(y.float()).all()
    assert rel(y, want) < (2e-5 if out32 else 5e-3), rel(y, want)
    if rm is not None and int((rm == 0).sum()) > 0:
        assert float(y[rm == 0].abs().max()) == 0.0


@pytest.mark.parametrize('M,K1,act,res,mask', [(300, 256, 1, 'f32', False), (1000, 1280, 2, None, True), (129, 64, 2, None, False),
                                               (20000, 256, 2, None, True), (40000, 1280, 1, 'f32', True), (1, 128, 1, 'bf16', False)])
def test_ffn_rows_tc_vs_torch(M, K1, act, res, mask):
    """case_ffn_rows_tc (linear1 -> gelu / relu -> linear2 in one launch, hidden tile kept in shared memory) against torch
    fp32 on the same bf16-rounded operands and a bf16-rounded hidden: 1 .. 20 K stages of the first product, partial last row
    tile, more row tiles than CTAs, residual and row mask."""
    from case_rg_b200 import _lib as L
    from case_rg_b200.producers import _Linear
    g = torch.Generator().manual_seed(M + K1)
    x = torch.randn(M, K1, generator=g).to(DEV).bfloat16()
    l1 = _Linear(torch.randn(H, K1, generator=g) / K1 ** 0.5, torch.randn(H, generator=g) * 0.1, torch.device(DEV))
    l2 = _Linear(torch.randn(H, H, generator=g) / H ** 0.5, torch.randn(H, generator=g) * 0.1, torch.device(DEV))
    r = None
    if res:
        r = torch.randn(M, H, generator=g).to(DEV)
        r = r.bfloat16() if res == 'bf16' else r
    rm = (torch.rand(M, generator=g) > 0.3).to(torch.uint8).to(DEV) if mask else None
    y = torch.full((M, H), float('nan'), device=DEV)
    L.call('case_ffn_rows_tc', x.data_ptr(), l1.wp.data_ptr(), l1.b.data_ptr(), K1, act, l2.wp.data_ptr(), l2.b.data_ptr(), M,
           L.ptr(r), (L.BF16 if res == 'bf16' else L.F32) if res else 0, L.ptr(rm), y.data_ptr(), L.F32, _st())
    torch.cuda.synchronize()
    hid = x.float() @ l1.w16.float().t() + l1.b
    hid = (F.gelu(hid) if act == 1 else torch.relu(hid)).bfloat16().float()
    want = hid @ l2.w16.float().t() + l2.b
    if r is not None:
        want = want + r.float()
    if rm is not None:
        want = want * rm.view(-1, 1).float()
    assert torch.isfinite(y).all()
    assert rel(y, want) < 1e-3, rel(y, want)
    if rm is not None and int((rm == 0).sum()) > 0:
        assert float(y[rm == 0].abs().max()) == 0.0


def _pipeline_inputs(B=3, Lq=20, NP=4, Lp=50, V=900, seed=5):
    sd = syn.make_case_producer_state(seed, V, H)
    inp = syn.make_case_inputs(seed + 1, B, Lq, NP, Lp, V, H)
    return sd, inp


@pytest.mark.parametrize('Lq,Lp', [(20, 50), (33, 150), (64, 256), (7, 64)])
def test_interaction_vs_oracle(Lq, Lp):
    """The Interaction kernel on fp32 encoder-like inputs against the oracle's Interaction (Interaction.py:15-76): both
    5H-wide outputs, PAD rows / columns zero, the max over the passages; an all-PAD-but-two passage included; one to four
    passage tiles, a partial last tile, the longest query."""
    from case_rg_b200 import _lib as L
    from oracle.producers import interaction
    B, NP = 3, 4
    g = torch.Generator().manual_seed(9)
    Eq, Ep = torch.randn(B, 1, Lq, H, generator=g), torch.randn(B, NP, Lp, H, generator=g)
    w = torch.randn(1, 3 * H, generator=g) * 0.05
    qm = torch.arange(Lq)[None, None, :] < torch.tensor([Lq, (2 * Lq) // 3, max(1, Lq // 3)])[:, None, None]
    pl = torch.randint(5, Lp + 1, (B, NP), generator=g)
    pl[1, 2] = 2
    pm = torch.arange(Lp)[None, None, :] < pl[:, :, None]
    want_q, want_p = interaction({'x.dual_att_linear.weight': w}, 'x.', Eq, Ep, qm, pm)
    keep = []                                   # device copies must outlive the launch (no temporaries behind data_ptr())

    def d(t):
        keep.append(t.to(DEV).contiguous())
        return keep[-1]
    Gqt = torch.empty(B * NP * Lq, 5 * H, device=DEV)
    Gq = torch.empty(B * Lq, 5 * H, dtype=torch.bfloat16, device=DEV)
    Gp = torch.full((B * NP * Lp, 5 * H), float('nan'), dtype=torch.bfloat16, device=DEV)
    L.call('case_interaction', d(Eq).data_ptr(), d(Ep).data_ptr(), d(qm.reshape(-1).to(torch.uint8)).data_ptr(),
           d(pm.reshape(-1).to(torch.uint8)).data_ptr(), d(w.reshape(-1)).data_ptr(), B, NP, Lq, Lp, Gqt.data_ptr(),
           Gq.data_ptr(), Gp.data_ptr(), _st())
    torch.cuda.synchronize()
    assert torch.isfinite(Gp.float()).all() and torch.isfinite(Gq.float()).all()
    # all five products take bf16 operands (fp32 accumulation) and the outputs are bf16
    assert rel(Gp.view(B, NP, Lp, -1), want_p) < 1.5e-2, rel(Gp.view(B, NP, Lp, -1), want_p)
    assert rel(Gq.view(B, 1, Lq, -1), want_q) < 1.5e-2, rel(Gq.view(B, 1, Lq, -1), want_q)
    assert float(Gp.view(B, NP, Lp, -1)[~pm.to(DEV)].abs().max()) == 0.0


@pytest.mark.parametrize('gemm', ['tc', 'cublas'])
def test_producers_vs_oracle(gemm):
    """The whole pre-decode pipeline from token ids against oracle.producers (fp32): encoder output, passage-selection
    representations and scores, the decoder's memories, prior and answer representation - within the bf16 band; the own
    tcgen05 GEMM and cuBLAS on the same operands give the same numbers."""
    from case_rg_b200.producers import CaseProducers
    from oracle.producers import producers
    sd, inp = _pipeline_inputs()
    want = producers(sd, inp.query, inp.passage)
    got = CaseProducers(sd, device=DEV, gemm=gemm)(inp.query, inp.passage)
    torch.cuda.synchronize()
    for k, tol in (('enc_p', 2e-2), ('enc_q', 2e-2), ('ps_p', 3e-2), ('mem_q', 3e-2), ('mem_p', 3e-2), ('answer_rep', 3e-2),
                   ('prior_p', 5e-2)):
        assert torch.isfinite(got[k]).all(), k
        assert rel(got[k], want[k]) < tol, (gemm, k, rel(got[k], want[k]))
    assert rel(got['rank'], want['passage_score']) < 5e-2
    assert torch.allclose(got['prior_p'].reshape(3, -1).sum(1), torch.ones(3, device=DEV), atol=1e-4)
    assert float(got['prior_p'][~inp.passage.ne(0).to(DEV)].abs().max()) == 0.0


def test_producers_edge_cases_vs_oracle():
    """Ragged edge of the pipeline: a one-token query next to the longest one (Lq = 64), a passage length that is not a
    multiple of any tile - against the oracle; and a passage that is all PAD, where the reference's encoder softmax over
    zero valid keys yields NaN for the whole query (torch semantics): here the padded passage contributes zeros, every
    output stays finite, and the other queries are not affected."""
    from case_rg_b200.producers import CaseProducers
    from oracle.producers import producers
    V, B, Lq, NP, Lp = 700, 2, 64, 3, 77
    sd = syn.make_case_producer_state(15, V, H)
    inp = syn.make_case_inputs(16, B, Lq, NP, Lp, V, H)
    query, passage = inp.query.clone(), inp.passage.clone()
    g = torch.Generator().manual_seed(17)
    query[0, 0] = torch.randint(305, V, (Lq,), generator=g)              # the longest query: no PAD at all
    query[1, 0, 1:] = 0                                                  # a one-token query
    want = producers(sd, query, passage)
    prod = CaseProducers(sd, device=DEV)
    got = prod(query, passage)
    torch.cuda.synchronize()
    for k, tol in (('enc_p', 2e-2), ('enc_q', 2e-2), ('ps_p', 3e-2), ('mem_q', 3e-2), ('mem_p', 3e-2), ('answer_rep', 3e-2),
                   ('prior_p', 5e-2)):
        assert torch.isfinite(want[k]).all() and torch.isfinite(got[k]).all(), k
        assert rel(got[k], want[k]) < tol, (k, rel(got[k], want[k]))
    assert rel(got['rank'], want['passage_score']) < 5e-2
    # an all-PAD passage in query 1
    passage2 = passage.clone()
    passage2[1, 2] = 0
    assert not torch.isfinite(producers(sd, query, passage2)['mem_p'][1]).all()      # the reference's answer: NaN
    got2 = prod(query, passage2)
    torch.cuda.synchronize()
    for k in ('enc_p', 'ps_p', 'mem_q', 'mem_p', 'answer_rep', 'prior_p', 'rank'):
        assert torch.isfinite(got2[k]).all(), k
    assert float(got2['ps_p'][1, 2].abs().max()) == 0.0 and float(got2['prior_p'][1, 2].abs().max()) == 0.0
    for k in ('mem_p', 'mem_q', 'prior_p', 'answer_rep'):
        assert torch.equal(got2[k][0], got[k][0]), k                                  # query 0 is untouched


@pytest.mark.timeout(600)
def test_producers_full_size_c2_vs_oracle():
    """The pipeline at BASELINE configs[1]'s full size (B = 64, 10 x 256 passages, Lq = 60, V = 30522) against the oracle
    evaluated on the same GPU in strict fp32 (TF32 off): the decoder's inputs - memories, prior, answer representation -
    and the passage ranks within the bf16 band; the ranking of the passages of every query agrees wherever the oracle's
    margin between neighbours exceeds the band."""
    from case_rg_b200.producers import CaseProducers
    from oracle.case_decoder import strict_fp32
    from oracle.producers import producers
    strict_fp32()
    V, B, Lq, NP, Lp = syn.BERT_VOCAB, 64, 60, 10, 256
    sd = syn.make_case_producer_state(25, V, H)
    inp = syn.make_case_inputs(26, B, Lq, NP, Lp, V, H)
    q, p = inp.query.to(DEV), inp.passage.to(DEV)
    want = producers({k: v.to(DEV) for k, v in sd.items()}, q, p)
    got = CaseProducers(sd, device=DEV)(q, p)
    torch.cuda.synchronize()
    res = {}
    for k, tol in (('enc_p', 2e-2), ('enc_q', 2e-2), ('ps_p', 3e-2), ('mem_q', 3e-2), ('mem_p', 3e-2), ('answer_rep', 3e-2),
                   ('prior_p', 5e-2)):
        assert torch.isfinite(got[k]).all(), k
        res[k] = rel(got[k], want[k])
        assert res[k] < tol, (k, res[k])
    res['rank'] = rel(got['rank'], want['passage_score'])
    assert res['rank'] < 5e-2
    ws, gs = want['passage_score'].cpu(), got['rank'].cpu()
    order = ws.argsort(1, descending=True)
    margin = (ws.gather(1, order)[:, :-1] - ws.gather(1, order)[:, 1:])
    gsorted = gs.gather(1, order)
    clear = margin > 5e-2 * float(ws.abs().max())
    assert bool(((gsorted[:, :-1] - gsorted[:, 1:]) > 0)[clear].all())
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/parity_producers_c2.json', 'w') as f:
        json.dump(dict(rel_err=res, clear_neighbour_pairs=int(clear.sum()), pairs=int(clear.numel())), f)


def test_decode_from_token_ids_matches_reference_golden():
    """Producers + decoder from token ids (FastCaSE.search_ids) against what the unmodified CaSE.forward(data, 'test')
    returned for the fixture of tests/golden/make_producers_golden.py: rank within the bf16 band, answers through the
    oracle-scored tie criterion of parity_tools (the fixture's decoder weights are peaked)."""
    from case_rg_b200 import generations as FG
    cfg, sd_prod, sd_dec, data, _ = producers_case()
    z = np.load(os.path.join(GOLDEN, 'case_producers.npz'))
    model = FG.FastCaSE(sd_dec, device=DEV, dtype='bf16', producers=sd_prod)
    out = model.search_ids({k: v.to(DEV) for k, v in data.items()}, cfg['T'], 1, mode='module_greedy')
    assert out['answer'].shape == z['answer'].shape
    assert rel(out['rank'], torch.from_numpy(z['rank'])) < 5e-2
    # answers: identical, or - where a row leaves the reference's sequence - the oracle (fp32 producers + decoder, which
    # reproduces the golden answers exactly on the CPU) rates the two tokens within the band of TWO bf16 stages
    import parity_tools as PT
    from oracle.case_decoder import CaseOracle
    from oracle.producers import producers
    o = producers(sd_prod, data['query'], data['passage'])
    inp = syn.CaseInputs(data['query'], data['passage'], data['source_map'], o['mem_q'], o['mem_p'], o['prior_q'], o['prior_p'],
                         o['answer_rep'], data['id'], cfg['V'])
    orc = CaseOracle(sd_dec)
    want, dists, scale = PT.oracle_greedy(lambda: orc.incremental(inp), cfg['B'], cfg['T'])
    assert np.array_equal(want.numpy(), z['answer'])
    res = PT.compare_greedy(out['answer'].cpu(), want, dists, 2 * 2e-2 * scale)
    assert res['miss'] == 0, res
